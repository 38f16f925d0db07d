#!/bin/bash
# Round check on a 2 x B200 box: GPU test suite (the data-parallel test included), smoke, and the bench exactly as the driver
# launches it at N = 1 and N = 2 (both arms).
mkdir -p gpurun_out
run() { name=$1; to=$2; shift; shift; echo "=== $name"; timeout $to "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
run pytest_gpu 1200 python -m pytest tests -q -m gpu -x
run smoke 300 python __graft_entry__.py --smoke
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TAILN=3 run bench_n1 600 python bench.py --gpus 1 --steps 30 --warmup 5
TAILN=3 run bench_n2 600 $TR --master-port 29561 bench.py --gpus 2 --steps 30 --warmup 5
TAILN=3 run ref_n1 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
TAILN=3 run ref_n2 600 $TR --master-port 29562 bench.py --impl reference --gpus 2 --steps 3 --warmup 1
