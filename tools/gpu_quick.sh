#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_mlp.py -q -m gpu -x -k "ns512 or ns256" > gpurun_out/pytest_512.log 2>&1; rc=$?; echo "mlp subset exit $rc"; tail -n 4 gpurun_out/pytest_512.log
[ $rc -ne 0 ] && exit 1
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
for i in 1 2; do timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), {k:v for k,v in d['roofline']['kernel_ms'].items()}, 'infer', round(d['render']['mlp_fwd_ms'],4), 'render Msamples/s', round(d['render']['msamples_per_sec']))
"; done
