#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 6; do
NERF_B200_DBG=$d timeout 300 python tools/trace_chain2.py > gpurun_out/trace2_dbg$d.log 2>&1; echo "dbg $d trace exit $?"; grep -n "===\|epilogue span\|MMA thread" gpurun_out/trace2_dbg$d.log | head -3; sed -n 120,126p gpurun_out/trace2_dbg$d.log
done
