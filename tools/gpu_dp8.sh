#!/bin/bash
mkdir -p gpurun_out
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n', d['n_gpus'], 'rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'render', d['render'].get('msamples_per_sec'), d['roofline']['kernel_ms'])
"; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/dp_check.py > gpurun_out/dp_check8.log 2>&1; echo "dp_check exit $?"; grep "p2p" gpurun_out/dp_check8.log | tail -4
for p in 1 0; do
NERF_B200_P2P=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$p bench.py --gpus 8 --steps 300 --warmup 5 --no-cpu 2>gpurun_out/n8_$p.err | tee gpurun_out/scale8_p2p$p.json | pick
done
