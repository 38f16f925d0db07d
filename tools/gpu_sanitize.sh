#!/bin/bash
# compute-sanitizer over a small subset of the parity tests (ONE tool per gpurun call: B200_PROFILING.md).
#   bash tools/gpu_sanitize.sh memcheck|initcheck|synccheck|racecheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL='(tcgen05_predict_matches_oracle and (ns64 or ns256 or ns512) and not big) or (test_step_gradients_loss_and_adam and (ns256-0 or ns64-0 or ns512))'
# the plain run first: the sanitizer only runs on a command that has just exited 0 without it
timeout 600 python -m pytest tests/test_gpu_mlp.py -q -m gpu -x -k "$SEL" tests/test_gpu_composite.py tests/test_gpu_sampling.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -1 gpurun_out/sanitize_plain.log
timeout 1500 compute-sanitizer --tool $TOOL --target-processes all --error-exitcode 7 --log-file gpurun_out/sanitize_$TOOL.log \
    python -m pytest tests/test_gpu_mlp.py -q -m gpu -x -k "$SEL" tests/test_gpu_composite.py tests/test_gpu_sampling.py > gpurun_out/sanitize_${TOOL}_pytest.log 2>&1
echo "compute-sanitizer $TOOL exit $?"
tail -3 gpurun_out/sanitize_${TOOL}_pytest.log
grep -c "=========" gpurun_out/sanitize_$TOOL.log
grep -E "ERROR SUMMARY|Invalid|Uninitialized|hazard|Barrier error" gpurun_out/sanitize_$TOOL.log | sort | uniq -c | head -20
