#!/bin/bash
# composite / sampler parity subset, then the stand-alone HBM stage timings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_composite.py tests/test_gpu_sampling.py tests/test_golden.py tests/test_gpu_e2e.py -q -m gpu -x > gpurun_out/pytest_stages.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_stages.log
timeout 200 python tools/hbm_stages.py > gpurun_out/hbm_stages.json 2>gpurun_out/hbm_stages.err; echo "stages exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/hbm_stages.json'))
for k,v in d.items(): print(k, v['ms'], 'ms', round(v['achieved']), 'GB/s', round(v['frac'],3))
PY
tail -3 gpurun_out/hbm_stages.err
