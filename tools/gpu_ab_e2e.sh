#!/bin/bash
# A/B of two prebuilt libraries on the end-to-end (host buffer) path
one() { cp nerf_rs_b200/build/$1 nerf_rs_b200/libnerf_b200.so; timeout 200 python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('$1', 'ms/step', round(d['ms_per_step'],4), 'e2e', round(e['value']), 'eager', round(e['eager_pixels']['value']), 'host idx', round(e['from_host_indices']['value']))
"; }
for i in 1 2 3; do one ab_old.so; one ab_new.so; done
cp nerf_rs_b200/build/ab_new.so nerf_rs_b200/libnerf_b200.so
