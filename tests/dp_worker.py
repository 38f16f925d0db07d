"""Data-parallel check, run under torchrun with one process per GPU (tests/test_gpu_dp.py launches it; tools/gpu_dp2.sh too).

1. One seed, N ranks: every rank draws its OWN rays -- pixel picks, view picks and depth jitter come from the Philox streams at the
   rank's global ray offset (rank * R), bit-identical to the oracle's streams -- so N x R rays form one batch of N*R rays.
2. The all-reduced gradient (nerf_get_grads: the cross-rank SUM) / N equals the oracle gradient of the CONCATENATED batch
   (the loss is a mean over a rank's R*4 elements, src/model.rs:298).
3. Replicas stay bit-identical over training iterations, with the peer-memory fused all-reduce + Adam and with NCCL + Adam,
   and the two exchange paths agree.
Prints DP_CHECK_OK on rank 0 when everything holds."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_rs_b200 as nb  # noqa: E402
from oracle import model_torch as M  # noqa: E402
from oracle import ray_c  # noqa: E402
from tests import gpu_util as G  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W_IMG, R, S, PICKS, HIDDEN = 64, 512, 32, 4, 128


def make(p2p):
    os.environ["NERF_B200_P2P"] = "1" if p2p else "0"
    # (deterministic weight gradients: without them the unordered atomics of 20 Adam steps amplify into a 1e-3..5e-3 difference
    #  between ANY two runs, which says nothing about the exchange paths being compared below)
    cfg = nb.default_config(image_w=W_IMG, image_h=W_IMG, num_rays=R, num_samples=S, hidden=HIDDEN, deterministic_grads=1)
    m = nb.NeRF(cfg, device=local)
    mcfg = G.model_cfg(cfg)
    params_t = M.init_params(mcfg, 0)
    m.set_weights(M.flatten_params(params_t).numpy())
    rng = np.random.default_rng(0)
    imgs = rng.random((4, W_IMG * W_IMG, 4)).astype(np.float32)
    m.set_images(imgs)
    angles = nb.get_view_angles(6)[:4]
    m.set_view_angles(angles)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(nb.NeRF.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    m.comm_init_rank(bytes(uid.cpu().numpy().tobytes()), rank, world)
    return m, cfg, mcfg, params_t, imgs, angles


def gather(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [x.cpu().numpy() for x in out]


# ---- 1 + 2: disjoint Philox batches, global-batch gradient
m, cfg, mcfg, params_t, imgs, angles = make(True)
seed = 0x5EED0001
b = m.get_batch(None, None, PICKS, None, True, seed)
base = rank * R
y = np.minimum((ray_c.philox_uniform(seed, 0, base, R) * np.float32(W_IMG)).astype(np.int64), W_IMG - 1)
x = np.minimum((ray_c.philox_uniform(seed, 1, base, R) * np.float32(W_IMG)).astype(np.int64), W_IMG - 1)
assert np.array_equal(b["indices"], np.stack([y, x], 1)), "pixel picks are not the Philox stream at the rank's global ray offset"
vi = np.minimum((ray_c.philox_uniform(seed, 2, base // (R // PICKS), PICKS) * np.float32(4)).astype(np.int64), 3)
u = ray_c.philox_uniform(seed, 3, base * S, R * S).reshape(R, S)
_, pts_o, t_o, gold_o = ray_c.get_multiview_batch(imgs, angles, b["indices"], vi, S, np.sort(u, axis=1), W_IMG, W_IMG)
assert b["t"].tobytes() == t_o.tobytes() and b["points"].tobytes() == pts_o.tobytes(), "view picks / jitter are not the rank's Philox streams"
assert b["gold"].tobytes() == gold_o.tobytes()
all_idx = gather(b["indices"])
for i in range(world):
    for j in range(i + 1, world):
        assert not np.array_equal(all_idx[i], all_idx[j]), f"ranks {i} and {j} drew the same pixels"
out, _ = m.predict(train=True)
nb.Trainer(m, 5e-4).step(out, b["gold"].reshape(-1))
g_sum = m.get_grads()                                   # cross-rank sum
cat = lambda k: np.concatenate(gather(b[k]))
pts_all, t_all, dirs_all, gold_all = cat("points"), cat("t"), cat("dirs"), cat("gold")
if rank == 0:
    tr = M.Trainer(M.replace(mcfg, emulate_bf16=True, emulate_bf16_grads=True), params_t, lr=5e-4)
    o, _ = tr.predict(torch.from_numpy(pts_all.reshape(-1)), torch.from_numpy(t_all.reshape(-1)), world * R, S, torch.from_numpy(dirs_all), literal=False)
    tr.step(o, torch.from_numpy(gold_all.reshape(-1)))
    want = tr.grads_flat().numpy()
    got = g_sum / world
    off = 0
    errs = []
    for i_dim, o_dim in mcfg.layer_dims():
        n = i_dim * o_dim + o_dim
        errs.append(float(np.linalg.norm(got[off:off + n] - want[off:off + n]) / np.linalg.norm(want[off:off + n])))
        off += n
    print(f"summed gradient / {world} vs the oracle gradient of the concatenated batch of {world * R} rays, per layer:", [f"{e:.2e}" for e in errs])
    assert max(errs) < 4e-2
m.close()


# ---- 3: replicas bit-identical; peer-memory exchange vs NCCL
def run(p2p, steps=20):
    m = make(p2p)[0]
    for it in range(steps):
        m.train_iter(100 + it)
    m.sync()
    w, g = m.get_weights(), m.get_grads()
    m.close()
    return w, g


res = {}
for p2p in (True, False):
    w, g = run(p2p)
    allw = gather(w)
    same = all(np.array_equal(allw[0], a) for a in allw)
    if rank == 0:
        print(f"p2p={p2p}: replicas bit-identical across {world} ranks after 20 steps: {same}")
    assert same
    res[p2p] = (w, g)
dw = np.abs(res[True][0] - res[False][0]).max() / np.abs(res[False][0]).max()
if rank == 0:
    print(f"peer-memory vs NCCL exchange after 20 steps: max |dw| / max |w| = {dw:.2e}")
assert dw < 1e-3   # (the two paths differ only in the order the ranks' gradients are summed in: nothing at 2 ranks)
dist.barrier()
if rank == 0:
    print("DP_CHECK_OK")
dist.destroy_process_group()
