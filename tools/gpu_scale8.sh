#!/bin/bash
# 8-GPU weak-scaling check (NCCL gradient all-reduce inside nerf_step), plus N=1 on the same box for the ratio.
mkdir -p gpurun_out
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n', d['n_gpus'], 'rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['roofline']['kernel_ms'])
"; }
timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu 2>gpurun_out/n1.err | tee gpurun_out/scale_n1.json | pick
for n in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 300 --warmup 5 --no-cpu 2>gpurun_out/n$n.err | tee gpurun_out/scale_n$n.json | pick
done
tail -3 gpurun_out/n8.err
