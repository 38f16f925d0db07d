#!/bin/bash
# same-box A/B: atomic weight-gradient flush vs deterministic_grads = 1
one() { timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu --no-extra $1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('[$1]', 'ms/step', round(d['ms_per_step'],4), d['roofline']['kernel_ms'])
"; }
for i in 1 2 3; do one ""; one "--deterministic"; done
