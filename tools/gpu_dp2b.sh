#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -q -m gpu -x 2>&1 | tail -5
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/dp_worker.log | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 30 --warmup 5 --no-extra --no-cpu 2>gpurun_out/n2.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n', d['n_gpus'], 'rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['clocks'], d['roofline']['kernel_ms'])
"
tail -3 gpurun_out/n2.err
