#!/bin/bash
mkdir -p gpurun_out
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'][:60], '| rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'render', d['render'], d['roofline']['kernel_ms'])
"; }
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --samples 192 2>gpurun_out/cfg2.err | tee gpurun_out/bench_cfg2.json | pick
tail -2 gpurun_out/cfg2.err
