#!/usr/bin/env python
"""Per-CTA timeline of k_wgrad at the bench shape: how evenly the CTAs finish, ring fill and accumulator-flush tails."""
import ctypes
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_rs_b200 as nb
import bench

hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = nb.default_config(hidden=hidden)
m = nb.NeRF(cfg)
m.set_weights(bench.synthetic_weights(m.cfg))
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
print('idle CTAs', int((o[:, 5] == 0).sum()))
o = o[o[:, 5] > 0]
t0 = o[:, 0].min()
end = (o[:, 3] - t0) / 1e3
mma = (o[:, 2] - t0) / 1e3
print("kernel span us", round(float(end.max()), 1), "start skew us", (o[:, 0].max() - t0) / 1e3)
print("CTA end us: min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % (end.min(), np.percentile(end, 10), np.median(end), np.percentile(end, 90), end.max()))
print("last MMA done us: min %.1f median %.1f max %.1f ; flush tail us mean %.1f" % (mma.min(), np.median(mma), mma.max(), float((end - mma).mean())))
print("ring fill us mean %.1f" % float(((o[:, 1] - o[:, 0]) / 1e3).mean()))
print("segments per CTA:", {int(k): int((o[:, 5] == k).sum()) for k in sorted(set(o[:, 5]))})
print("GB loaded %.3f -> %.0f GB/s over the span" % (o[:, 7].sum() / 1e9, o[:, 7].sum() / 1e9 / (end.max() * 1e-6)))
late = np.argsort(end)[-5:]
print("slowest CTAs (cta, first unit, segs, iters, end us):", [(int(i), int(o[i, 4]), int(o[i, 5]), int(o[i, 6]), round(float(end[i]), 1)) for i in late])
