#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -q -m gpu -x > gpurun_out/pytest_mlp.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_mlp.log
timeout 300 python tools/trace_chain.py > gpurun_out/trace.log 2>&1; echo "trace exit $?"; tail -n 5 gpurun_out/trace.log
