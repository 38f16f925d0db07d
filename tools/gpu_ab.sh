#!/bin/bash
# A/B of environment switches on the quick bench: prints the per-kernel times
mkdir -p gpurun_out
one() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k:v for k,v in d['roofline']['kernel_ms'].items()})
"; }
one A=1
one NERF_B200_NO_SIGMA_FOLD=1
one NERF_B200_NO_FC9_MERGE=1
one NERF_B200_NO_SIGMA_FOLD=1 NERF_B200_NO_FC9_MERGE=1
