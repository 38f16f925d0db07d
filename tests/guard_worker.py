"""Runs the hot path with NERF_B200_GUARD=1 (guard bands around every device buffer, NaN fill: csrc/guard.h) and reports whether
any band was overwritten and whether anything uninitialised reached an output. Launched by tests/test_gpu_guards.py in its own
process (the mode is read once, at the library's first allocation)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
assert os.environ.get("NERF_B200_GUARD") == "1"
import nerf_rs_b200 as nb  # noqa: E402

lib = nb.load()
rng = np.random.default_rng(0)
for hidden, rays, samples, impl in ((256, 300, 48, 0), (512, 96, 64, 0), (64, 70, 50, 0), (100, 84, 64, 0), (256, 300, 48, 3)):
    kw = dict(xyz_freqs=0, dir_freqs=-1, skip_layer=0, use_rgb_head=0) if hidden == 100 else {}
    cfg = nb.default_config(image_w=40, image_h=40, num_rays=rays, num_samples=samples, hidden=hidden, mlp_impl=impl, **kw)
    m = nb.NeRF(cfg)
    m.set_images(rng.random((3, 1600, 4)).astype(np.float32))
    m.set_view_angles(nb.get_view_angles(6)[:3])
    b = m.get_batch(None, None, 1, None, True, 5)
    out, sig = m.predict(train=True)
    loss = nb.Trainer(m).step(out, b["gold"].reshape(-1))
    for it in range(3):
        m.train_iter(10 + it)
    m.sync()
    o2, s2 = m.predict(b["points"].reshape(-1), b["t"].reshape(-1), b["dirs"].reshape(-1) if cfg.dir_freqs >= 0 else None, train=False)
    frame = m.render_sharded(0.3, 0.2, randomize=True, seed=1, packed=True)
    vals = [out, sig, o2, s2, m.get_grads(), m.get_weights(), np.array([loss, m.last_loss()])]
    assert all(np.isfinite(v).all() for v in vals), f"hidden {hidden}: a NaN (uninitialised read?) reached an output"
    n = ctypes.c_int32()
    bad = lib.nerf_debug_check_guards(ctypes.byref(n))
    print(f"hidden {hidden} impl {impl}: {n.value} guarded allocations, {bad} with overwritten guard bands")
    assert bad == 0 and n.value > 10
    m.close()
print("GUARDS_OK")
