"""Data-parallel check, run under torchrun (one process per GPU):
   replicas stay bit-identical, and the peer-memory fused all-reduce+Adam matches the NCCL all-reduce + Adam path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import nerf_rs_b200 as nb  # noqa: E402
from oracle import model_torch as M  # noqa: E402
from tests import gpu_util as G  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def run(p2p, steps=20):
    os.environ["NERF_B200_P2P"] = "1" if p2p else "0"
    cfg = nb.default_config(image_w=64, image_h=64, num_rays=512, num_samples=32, hidden=128)
    m = nb.NeRF(cfg, device=local)
    m.set_weights(M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy())
    rng = np.random.default_rng(0)
    m.set_images(rng.random((4, 64 * 64, 4)).astype(np.float32))
    m.set_view_angles(nb.get_view_angles(6)[:4])
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(nb.NeRF.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    m.comm_init_rank(bytes(uid.cpu().numpy().tobytes()), rank, world)
    for it in range(steps):
        m.train_iter(100 + it)
    m.sync()
    w, g = m.get_weights(), m.get_grads()
    names = m.profile_read()
    m.profile(True)
    for it in range(5):
        m.train_iter(900 + it)
    names = m.profile_read()
    m.profile(False)
    m.close()
    return w, g, sorted(names)


for p2p in (True, False):
    w, g, names = run(p2p)
    tw = torch.from_numpy(w).cuda()
    allw = [torch.empty_like(tw) for _ in range(world)]
    dist.all_gather(allw, tw)
    same = all(torch.equal(allw[0], x) for x in allw)
    if rank == 0:
        print(f"p2p={p2p}: replicas bit-identical across {world} ranks: {same}; kernels: {[n for n in names if 'adam' in n or 'allreduce' in n]}")
    assert same
    if p2p:
        w_p2p, g_p2p = w, g
    else:
        dw = np.abs(w - w_p2p).max() / np.abs(w).max()
        dg = np.linalg.norm(g - g_p2p) / np.linalg.norm(g)
        if rank == 0:
            print(f"p2p vs nccl after 20 steps: max |dw|/|w| = {dw:.2e}, summed-gradient rel diff at the last step = {dg:.2e}")
        assert dw < 5e-3
dist.barrier()
dist.destroy_process_group()
