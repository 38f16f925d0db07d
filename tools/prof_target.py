#!/usr/bin/env python
"""Profiling target for ncu: a few training iterations at the bench shape (BASELINE configs[1]), one inference pass, then the
HBM-bound stage kernels alone at render scale. Small on purpose: ncu replays every kernel ~40 times."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_rs_b200 as nb  # noqa: E402
import bench  # noqa: E402  (synthetic weights / images, no oracle code)

m = nb.NeRF(nb.default_config())
m.set_weights(bench.synthetic_weights(m.cfg))
rng = np.random.default_rng(1)
ang = nb.get_view_angles(6)
m.set_images(rng.random((len(ang), 800 * 800, 4), dtype=np.float32))
m.set_view_angles(ang)
for it in range(4):
    m.train_iter(1 + it)
m.get_batch(None, None, 64, None, True, 99, want=())
m.predict(train=False, want_sigma=False)
for st in ("sample", "composite_fwd", "composite_bwd"):
    m.bench_stage(st, 131072, 192, 1)
m.sync()
print("ok", m.last_loss())
