"""Determinism probe: same seeds, two fresh contexts -> per-step relative loss difference."""
import sys
import numpy as np
sys.path.insert(0, ".")
import nerf_rs_b200 as nb
from oracle import model_torch as M
from tests import gpu_util as G
from tests.test_gpu_e2e import _sphere_images

def run(hidden, rays, samples, steps=60):
    cfg = nb.default_config(image_w=64, image_h=64, num_rays=rays, num_samples=samples, hidden=hidden)
    out = []
    for rep in range(2):
        m = nb.NeRF(cfg)
        m.set_weights(M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy())
        m.set_images(_sphere_images(4, 64, 64))
        m.set_view_angles(nb.get_view_angles(6)[:4])
        ls = []
        for it in range(steps):
            m.train_iter(1000 + it)
            ls.append(m.last_loss())
        out.append(np.array(ls))
    d = np.abs(out[0] - out[1]) / np.abs(out[0])
    print(f"hidden {hidden} rays {rays} S {samples}: first step with rel diff > 1e-6: {int(np.argmax(d > 1e-6)) if (d > 1e-6).any() else None}, max rel diff {d.max():.3e} at step {int(d.argmax())}")
    print("   ", np.array2string(d[:12], precision=2), "...", np.array2string(d[-6:], precision=2))

for rep in range(3):
    run(128, 256, 32)
run(256, 1024, 64)
run(64, 256, 32)
