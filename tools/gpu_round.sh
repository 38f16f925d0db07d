#!/bin/bash
# Round check on a B200: full GPU test suite, bench, then the ncu launch list + one full capture.
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
run pytest_gpu 900 python -m pytest tests -q -m gpu -x
run smoke 300 python __graft_entry__.py --smoke
TAILN=4 run bench 600 python bench.py --steps 30 --warmup 5
if [ "$1" == "ncu" ]; then
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_short.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_wgrad" -s 12 -c 4 -o gpurun_out/prof_mlp python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
