"""Data-parallel semantics on CPU (gloo, world_size 2): summing per-rank gradients and scaling by
1/N -- what nerf_step does around its NCCL all-reduce -- equals the single-process gradient of the
concatenated batch, because the loss is a per-rank mean over R*4 elements (src/model.rs:298)."""
import os
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model_torch as M
from tests import gpu_util as G


def _grads(mcfg, params, pts, t, dirs, gold, r, s):
    tr = M.Trainer(mcfg, params, lr=5e-4)
    out, _ = tr.predict(torch.from_numpy(pts), torch.from_numpy(t), r, s, torch.from_numpy(dirs), literal=False)
    loss = tr.step(out, torch.from_numpy(gold))
    return tr.grads_flat(), loss


def _worker(rank, world, path, r, s):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = "29613"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    mcfg = M.ModelConfig(hidden=32)
    params = M.init_params(mcfg, 0)
    pts, t, dirs, gold = G.make_points(r * world, s, 7)
    sl = slice(rank * r, (rank + 1) * r)
    g, loss = _grads(mcfg, params, pts.reshape(world * r, -1)[sl].reshape(-1).copy(), t.reshape(world * r, -1)[sl].reshape(-1).copy(),
                     dirs[sl].copy(), gold.reshape(world * r, 4)[sl].reshape(-1).copy(), r, s)
    dist.all_reduce(g, op=dist.ReduceOp.SUM)          # ncclAllReduce(sum) in the product
    g = g * (1.0 / world)                             # grad_scale folded into K-adam
    lt = torch.tensor([loss])
    dist.all_reduce(lt)
    if rank == 0:
        np.savez(path, g=g.numpy(), loss=lt.item() / world)
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_equals_global_batch_gradient():
    world, r, s = 2, 8, 16
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "dp.npz")
        mp.spawn(_worker, args=(world, path, r, s), nprocs=world, join=True)
        z = np.load(path)
    mcfg = M.ModelConfig(hidden=32)
    params = M.init_params(mcfg, 0)
    pts, t, dirs, gold = G.make_points(r * world, s, 7)
    g, loss = _grads(mcfg, params, pts, t, dirs, gold, r * world, s)
    assert np.allclose(z["g"], g.numpy(), rtol=1e-4, atol=1e-7)
    assert abs(float(z["loss"]) - loss) < 1e-6
