#!/bin/bash
# A/B of two prebuilt libraries (nerf_rs_b200/build/ab_old.so vs ab_new.so) on ONE box, alternating, quick bench each
one() { cp nerf_rs_b200/build/$1 nerf_rs_b200/libnerf_b200.so; timeout 200 python bench.py --steps ${STEPS:-300} --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'ms/step', round(d['ms_per_step'],4), {k:v for k,v in d['roofline']['kernel_ms'].items() if k.startswith('mlp')}, 'infer', round(d['render']['mlp_fwd_ms'],4))
"; }
for i in 1 2 3; do one ab_old.so; one ab_new.so; done
cp nerf_rs_b200/build/ab_old.so nerf_rs_b200/libnerf_b200.so
