#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_golden.py -q -m gpu -x > gpurun_out/pytest_mlp.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_mlp.log
timeout 300 python tools/trace_chain2.py > gpurun_out/trace2.log 2>&1; echo "trace exit $?"; grep -n "===\|epilogue span\|MMA thread\|producer" gpurun_out/trace2.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/bench_quick.log 2>&1; echo "bench exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/bench_quick.log'):
    if l.startswith('{'):
        d=json.loads(l); print('rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['roofline']['kernel_ms'], 'render', d['render'])
PY
[ -n "$PROBE" ] && timeout 120 python tools/hbm_probe.py
