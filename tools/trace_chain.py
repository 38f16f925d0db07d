"""Timeline of CTA 0 of the fused MLP chain kernel (clock64 stamps): where the cycles of a tile go."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import nerf_rs_b200 as nb  # noqa: E402
from tests import gpu_util as G  # noqa: E402
from tests import tc_plan_util as U  # noqa: E402

KINDS = ["PRO_F", "RELU", "LIN", "SIGMA", "RGBA", "PRO_B", "DMASK", "DCOPY"]


def main(program=1, rays=4096, samples=64):
    cfg = nb.default_config(image_w=100, image_h=100, num_rays=rays, num_samples=samples, hidden=256)
    m = nb.NeRF(cfg)
    pts, t, dirs, gold = G.make_points(rays, samples, 1)
    out, _ = m.predict(pts, t, dirs.reshape(-1), train=True)
    nb.Trainer(m).step(out, gold)           # leaves d_sigma / d_rgba / rgba populated for program 2
    m.predict(pts, t, dirs.reshape(-1), train=True)
    buf = np.zeros((3, 2048, 2), dtype=np.uint64)
    rc = m.lib.nerf_debug_trace(m.h, program, buf.ctypes.data)
    assert rc == 0, rc
    plan = U.get_plan(cfg, program)
    ops, jobs = plan["ops"], plan["jobs"]
    n_ops, n_jobs = len(ops), len(jobs)
    mma = buf[0].astype(np.int64)
    epi = buf[1].astype(np.int64)
    prod = buf[2].astype(np.int64)
    t0 = min(int(mma[0, 0]), int(epi[0, 0]))
    tiles = min(3, 2048 // (2 * n_ops))
    print(f"program {program}: {n_ops} ops, {n_jobs} jobs per tile")
    for tile in range(tiles):
        print(f"--- tile {tile} (cycles relative to kernel start of CTA 0)")
        print("MMA issuer: op  slot acc n  k | start  wait_bars  wait_full  issue")
        for i in range(n_ops):
            e = (tile * n_ops + i) * 2
            s0, s1, s2, s3 = mma[e, 0] - t0, mma[e, 1] - t0, mma[e + 1, 0] - t0, mma[e + 1, 1] - t0
            o = ops[i]
            print(f"  op{i:3d} s{o['a_slot']} a{o['acc']} n{o['n']:3d} k{o['kcount']} {'F' if o['flags'] & 1 else ' '}{'C' if o['flags'] & 2 else ' '} | {s0:8d} +{s1 - s0:6d} +{s2 - s1:6d} +{s3 - s2:6d}")
        print("epilogue warp 0: job kind acc | start  wait_acc  work  fence+arrive")
        for j in range(n_jobs):
            e = (tile * n_jobs + j) * 2
            s0, s1, s2, s3 = epi[e, 0] - t0, epi[e, 1] - t0, epi[e + 1, 0] - t0, epi[e + 1, 1] - t0
            jb = jobs[j]
            print(f"  job{j:3d} {KINDS[jb['kind']]:6s} a{jb['acc'] if jb['acc'] != 255 else '-'}c{jb['acc_col']:3d} n{jb['ncols']:3d} | {s0:8d} +{s1 - s0:6d} +{s2 - s1:6d} +{s3 - s2:6d}")
    e_last = (tiles * n_jobs - 1) * 2 + 1
    print("cycles per tile (epilogue end to end):", (epi[e_last, 1] - epi[0, 0]) / tiles)
    pw = prod[:tiles * n_ops]
    print("producer: mean wait for an empty stage:", float((pw[:, 1] - pw[:, 0]).mean()))


if __name__ == "__main__":
    for prog in (1, 0, 2):
        main(prog)
