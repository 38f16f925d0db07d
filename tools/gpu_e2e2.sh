#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_sampling.py tests/test_cpp_host.py tests/test_metrics.py -q -m gpu -x 2>&1 | tail -4
bash tools/gpu_ab_env.sh NERF_B200_NO_SAMPLER_OVERLAP 30 2>&1 | sed 's/loss.*//'
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('e2e', round(e['value']), 'eager', round(e['eager_pixels']['value']), 'host idx', round(e['from_host_indices']['value']))
"
NERF_B200_NO_SAMPLER_OVERLAP=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('no-overlap: e2e', round(e['value']), 'eager', round(e['eager_pixels']['value']), 'host idx', round(e['from_host_indices']['value']))
"
