#!/bin/bash
# A/B of environment switches on the quick bench: prints the per-kernel times.  usage: gpu_ab.sh "VAR=1" "VAR=2 OTHER=1" ...
mkdir -p gpurun_out
one() { echo "== $*"; env $* timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k:v for k,v in d['roofline']['kernel_ms'].items()}, 'infer', round(d['render']['mlp_fwd_ms'],4))
"; }
one A=1
for v in "$@"; do one $v; done
