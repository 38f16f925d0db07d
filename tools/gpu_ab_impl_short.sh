#!/bin/bash
# like gpu_ab_impl.sh but with the driver's short timed window (20-30 steps at burst clocks)
one() { timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --mlp-impl $1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('impl $1', 'ms/step', round(d['ms_per_step'],4), d['roofline']['kernel_ms'], 'infer', round(d['render']['mlp_fwd_ms'],4), 'clk', d['clocks']['sm_mhz'], 'render Msamples/s', round(d['render']['msamples_per_sec']))
"; }
for i in 1 2 3; do one 3; one 0; sleep 5; done
