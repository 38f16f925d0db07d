#!/bin/bash
# same library, same box: quick bench with and without an environment switch.  bash tools/gpu_ab_env.sh VAR [steps]
VAR=$1; STEPS=${2:-30}
one() { timeout 200 python bench.py --steps $STEPS --warmup 5 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'loss', d['final_loss'])
"; }
for i in 1 2 3; do one default; env $VAR=1 bash -c "$(declare -f one); one $VAR"; done
