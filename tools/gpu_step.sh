#!/bin/bash
# bring-up of a kernel change: MLP/e2e GPU tests on the fresh build, then (if they pass) the same-box A/B against build/ab_old.so
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -q -m gpu -x --timeout 120 > gpurun_out/pytest_step.log 2>&1; rc=$?
echo "pytest mlp exit $rc"; tail -n 25 gpurun_out/pytest_step.log
[ $rc -ne 0 ] && exit 1
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_metrics.py -q -m gpu -x --timeout 120 > gpurun_out/pytest_step2.log 2>&1; rc=$?
echo "pytest e2e exit $rc"; tail -n 15 gpurun_out/pytest_step2.log
[ $rc -ne 0 ] && exit 1
bash tools/gpu_ab_impl.sh 2>&1 | tee gpurun_out/ab.log
