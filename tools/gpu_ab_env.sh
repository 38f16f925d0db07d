#!/bin/bash
# same library, same box: quick bench with and without an environment switch.  bash tools/gpu_ab_env.sh VAR [steps]
VAR=$1; STEPS=${2:-30}
one() { timeout 200 python bench.py --steps $STEPS --warmup 5 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(e['value']), 'eager', round(e['eager_pixels']['value']), 'host idx', round(e['from_host_indices']['value']))
"; }
for i in 1 2 3; do one default; export $VAR=1; one $VAR; unset $VAR; done
