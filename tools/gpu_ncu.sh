#!/bin/bash
# ncu evidence for profiles/: the plain run first (must exit 0), then the launch list, then ONE --set full capture
#   bash tools/gpu_ncu.sh <tag>          (default tag r02)
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 python tools/prof_target.py > gpurun_out/prof_target.log 2>&1 || { echo "plain run failed"; tail gpurun_out/prof_target.log; exit 1; }
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_short.log 2>&1 || { echo "bench failed"; tail gpurun_out/bench_short.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_wgrad|k_composite|k_sample|k_adam" -s 10 -c 24 -f -o gpurun_out/prof_${TAG} python tools/prof_target.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_${TAG}.ncu-rep
