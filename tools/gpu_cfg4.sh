#!/bin/bash
mkdir -p gpurun_out
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'][:70], '| rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'mlp TF/s', round(d['roofline']['mlp_fwd_dgrad_wgrad_tflops'],1), {k:(round(v['achieved']),v['unit'],round(v['frac'],3)) for k,v in d['roofline']['kernels'].items() if k.startswith('mlp')}, d['roofline']['kernel_ms'], 'render', d['render'])
"; }
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --rays 65536 --samples 128 --hidden 512 2>gpurun_out/cfg4.err | tee gpurun_out/bench_cfg4.json | pick
tail -3 gpurun_out/cfg4.err
