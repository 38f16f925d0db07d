#!/bin/bash
# the bench exactly as the driver launches it at N GPUs (both arms): bash tools/gpu_bench_n.sh N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29581 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
timeout 300 $TR --master-port 29582 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/ref_n$N.log 2> gpurun_out/ref_n$N.err; echo "ref exit $?"
grep "^{" gpurun_out/bench_n$N.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['clocks'])
    print('kernel_ms', d['roofline']['kernel_ms'])
    print('sustained', d['sustained'])
    for k,v in d['configs'].items(): print(k, round(v['rays_per_sec']), round(v['ms_per_step'],4), v['kernel_ms'])
    print('render', d['render'])
"
grep "^{" gpurun_out/ref_n$N.log | cut -c1-300
tail -3 gpurun_out/bench_n$N.err
