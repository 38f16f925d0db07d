#!/bin/bash
# one `ncu --set full --import-source` capture of a chain kernel (regex in $1, default the backward chain) for the source page
mkdir -p gpurun_out
K=${1:-"k_chain3<1"}
timeout 300 python tools/prof_target.py > gpurun_out/prof_target.log 2>&1 || { echo "plain run failed"; tail gpurun_out/prof_target.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain3 -s ${2:-3} -c 1 -f -o gpurun_out/prof_src python tools/prof_target.py > gpurun_out/ncu_src.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_src.log; ls -la gpurun_out/prof_src.ncu-rep
