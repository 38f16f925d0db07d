#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/det_check.py > gpurun_out/det.log 2>&1; echo "det exit $?"; cat gpurun_out/det.log | tail -20
timeout 600 python -m pytest tests/test_gpu_mlp.py -q -m gpu -x > gpurun_out/pytest_mlp.log 2>&1; echo "mlp exit $?"; tail -n 5 gpurun_out/pytest_mlp.log
