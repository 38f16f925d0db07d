"""Timeline of CTA 0 (leader of cluster 0) of the CTA-pair chain kernel k_chain2 (clock64 stamps)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import nerf_rs_b200 as nb  # noqa: E402
from tests import gpu_util as G  # noqa: E402


def main(program=1, rays=4096, samples=64, show_ops=90, show_steps=60):
    cfg = nb.default_config(image_w=100, image_h=100, num_rays=rays, num_samples=samples, hidden=256)
    m = nb.NeRF(cfg)
    pts, t, dirs, gold = G.make_points(rays, samples, 1)
    out, _ = m.predict(pts, t, dirs.reshape(-1), train=True)
    nb.Trainer(m).step(out, gold)
    m.predict(pts, t, dirs.reshape(-1), train=True)
    buf = np.zeros((3, 2048, 4), dtype=np.uint64)
    rc = m.lib.nerf_debug_trace(m.h, program, buf.ctypes.data)
    assert rc == 0, rc
    mma, epi2, prod = (buf[i].astype(np.int64) for i in range(3))
    epi, epx = epi2[0::2], epi2[1::2]
    n_mma = int((mma[:, 3] > 0).sum())
    n_epi = int((epi[:, 3] > 0).sum())
    t0 = int(min(mma[0, 0], epi[0, 0]))
    print(f"=== program {program}: {n_mma} MMA ops, {n_epi} epilogue steps traced")
    print("MMA op: start | wait(epi_done/full) | issue+commit | period")
    for i in range(min(show_ops, n_mma)):
        s0, s1, s2, s3 = (int(x) - t0 for x in mma[i])
        nxt = int(mma[i + 1, 0]) - t0 if i + 1 < n_mma else s3
        print(f"  op{i:4d} {s0:8d} +{s1 - s0:6d} +{s3 - s2:6d} | {nxt - s0:6d}")
    print("epilogue step: start | pre(wait_read/prologue) | wait acc_full | tmem->panels | enc+signal | save | period")
    for i in range(min(show_steps, n_epi)):
        s0, s1, s2, s3 = (int(x) - t0 for x in epi[i])
        x3, x4, xa, xb = (int(v) - t0 for v in epx[i])
        nxt = int(epi[i + 1, 0]) - t0 if i + 1 < n_epi else s3
        print(f"  st{i:4d} {s0:8d} +{s1 - s0:6d} +{s2 - s1:6d} +{x3 - s2:6d} +{x4 - x3:6d} (fence.proxy {xa - x3:5d} tcfence+syncwarp {xb - xa:5d} arrive {x4 - xb:5d}) +{s3 - x4:6d} | {nxt - s0:6d}")
    total = int(epi[n_epi - 1, 3] - epi[0, 0])
    print(f"epilogue span {total} cycles over {n_epi} steps = {total / n_epi:.0f} per step; "
          f"acc wait total {int((epi[:n_epi, 2] - epi[:n_epi, 1]).sum())}, work total {int((epi[:n_epi, 3] - epi[:n_epi, 2]).sum())}")
    w = mma[:n_mma]
    print(f"MMA thread: wait {int((w[:, 1] - w[:, 0]).sum())}, issue {int((w[:, 3] - w[:, 2]).sum())}, span {int(w[-1, 3] - w[0, 0])}")
    np_ = int((prod[:, 1] > 0).sum())
    print(f"producer: mean wait for an empty stage {float((prod[:np_, 1] - prod[:np_, 0]).mean()):.0f} over {np_} ops")


if __name__ == "__main__":
    for prog in (1, 0, 2):
        main(prog)
