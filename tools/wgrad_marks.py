#!/usr/bin/env python
"""Per-CTA timeline of k_wgrad at the bench shape: which unit / CTA finishes last, and how long the ring takes to fill."""
import ctypes
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_rs_b200 as nb
from oracle import model_torch as M

cfg = nb.default_config()
m = nb.NeRF(cfg)
m.set_weights(M.flatten_params(M.init_params(M.ModelConfig(hidden=256), 0)).numpy())
rng = np.random.default_rng(1)
ang = nb.get_view_angles(6)
m.set_images(rng.random((len(ang), 800 * 800, 4), dtype=np.float32))
m.set_view_angles(ang)
for it in range(20):
    m.train_iter(it)
m.sync()
out = np.zeros((256, 8), dtype=np.uint64)
n = m.lib.nerf_debug_wgrad_marks(m.h, out.ctypes.data_as(ctypes.c_void_p), 256)
assert n > 0, n
o = out[:n].astype(np.int64)
t0 = o[:, 0].min()
dur = (o[:, 3] - o[:, 0]) / 1e3
print("kernel span us", (o[:, 3].max() - t0) / 1e3, "start skew us", (o[:, 0].max() - t0) / 1e3)
print("unit panels ctas tiles/cta  first_stage_us  mma_done_us(min/max)  end_us(min/max)")
for u in sorted(set(o[:, 4])):
    r = o[o[:, 4] == u]
    tiles = r[:, 6] - r[:, 5]
    print(int(u), int(r[0, 7]), len(r), tiles.min(), tiles.max(), round(float(((r[:, 1] - r[:, 0]) / 1e3).mean()), 1),
          round(float((r[:, 2] - t0).min() / 1e3), 1), round(float((r[:, 2] - t0).max() / 1e3), 1),
          round(float((r[:, 3] - t0).min() / 1e3), 1), round(float((r[:, 3] - t0).max() / 1e3), 1))
us_per_iter = (o[:, 2] - o[:, 1]) / 1e3 / np.maximum(1, 2 * (o[:, 6] - o[:, 5]))
print("us per half tile by panels:", {int(p): round(float(us_per_iter[o[:, 7] == p].mean()), 3) for p in sorted(set(o[:, 7]))})
