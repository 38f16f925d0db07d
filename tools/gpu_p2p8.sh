#!/bin/bash
# N-GPU A/B of the gradient exchange on ONE box: two-phase peer-memory kernel (NERF_B200_P2P=1, default), the one-shot
# peer-memory kernel (=2), NCCL all-reduce + Adam (=0), and the single-GPU step for the efficiency denominator.
#   gpurun --gpus 8 -- 'bash tools/gpu_p2p8.sh 8'
N=${1:-8}
mkdir -p gpurun_out
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'n', d['n_gpus'], 'rays/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k:v for k,v in d['roofline']['kernel_ms'].items() if 'adam' in k or 'allreduce' in k})
"; }
timeout 200 python bench.py --gpus 1 --steps 300 --warmup 5 --no-cpu --no-extra 2>gpurun_out/p2p_n1.err | pick single
for p in 1 2 0; do
NERF_B200_P2P=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$p bench.py --gpus $N --steps 300 --warmup 5 --no-cpu --no-extra 2>gpurun_out/p2p_n${N}_$p.err | pick "P2P=$p"
done
