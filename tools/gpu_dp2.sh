#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check exit $?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/dp_check.log | tail -12
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n', d['n_gpus'], 'rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['roofline']['kernel_ms'])
"; }
for p in 1 0; do
NERF_B200_P2P=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$p bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu 2>gpurun_out/n2_$p.err | pick
done
tail -3 gpurun_out/n2_1.err
