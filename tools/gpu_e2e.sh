#!/bin/bash
# end-to-end path check: the e2e / host-mirror tests, then the bench line's e2e keys with and without the early loss read
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_cpp_host.py tests/test_metrics.py tests/test_gpu_mlp.py -q -m gpu -x 2>&1 | tail -8
one() { timeout 600 python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu --no-extra 2>gpurun_out/e2e.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(e['value']), 'eager', round(e['eager_pixels']['value']), 'host idx', round(e['from_host_indices']['value']))
"; tail -2 gpurun_out/e2e.err; }
for i in 1 2; do one early; done; NERF_B200_STEP_SYNC=1 one sync
