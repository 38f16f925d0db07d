#!/bin/bash
# First-contact run on a B200: every stage under its own timeout so a hang cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run sampling python -m pytest tests/test_gpu_sampling.py -x -q -m gpu
run composite python -m pytest tests/test_gpu_composite.py -x -q -m gpu
run mlp_simt python -m pytest tests/test_gpu_mlp.py -x -q -m gpu -k "simt_fp32 or asserts"
run diag python tools/diag_mlp.py
run mlp_tc python -m pytest tests/test_gpu_mlp.py -q -m gpu -k "tcgen05 or step_gradients" 
run e2e python -m pytest tests/test_gpu_e2e.py -q -m gpu
run smoke python __graft_entry__.py --smoke
