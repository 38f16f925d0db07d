#!/bin/bash
# cycle counters of the TS-mode chain kernel: run tools/tc3_stats.py on the -DNERF_TC3_STATS build, then restore the product library
mkdir -p gpurun_out
cp nerf_rs_b200/libnerf_b200.so /tmp/prod.so
cp nerf_rs_b200/build/ab_stats.so nerf_rs_b200/libnerf_b200.so
PYTHONPATH=. timeout 300 python tools/tc3_stats.py "$@" 2>&1 | tee gpurun_out/tc3_stats.log
cp /tmp/prod.so nerf_rs_b200/libnerf_b200.so
