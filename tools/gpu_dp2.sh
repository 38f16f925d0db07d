#!/bin/bash
# 2-GPU check: golden GPU tests on one GPU, then the data-parallel bench at N=2 (NCCL all-reduce).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_golden.py -q -m gpu -x > gpurun_out/pytest_golden.log 2>&1; echo "golden exit $?"; tail -n 6 gpurun_out/pytest_golden.log
timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu > gpurun_out/bench_n1.log 2>&1; echo "bench n1 exit $?"; tail -n 3 gpurun_out/bench_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit $?"; tail -n 5 gpurun_out/bench_n2.log
