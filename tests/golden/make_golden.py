#!/usr/bin/env python
"""Regenerates the committed golden fixtures in this directory.

    python tests/golden/make_golden.py

What is a *reference* value and what is an *oracle* value is kept apart:

* ``reference_kat.json`` -- the only value-pinning vectors the reference's own tests hold
  (src/ray_sampling.rs:443-449 `point_rotates_to_90`, :70-77 `testRotatePitch`), copied as
  decimal literals from the Rust source. They are NOT produced by this script; it only re-checks
  that the oracle still reproduces them before writing anything else.
* ``ray_golden.npz`` / ``model_golden.npz`` -- outputs of the CPU oracle (oracle/ray_oracle.c,
  oracle/model_torch.py) on small seeded inputs. The reference (Rust + tch, Device::Mps) cannot
  be built or imported in this image, so these freeze the *restated* arithmetic: they catch
  drift of the oracle and give the GPU tests a fixture that does not depend on the host's libm
  or torch build. Inputs are stored with the outputs, so nothing is re-derived at test time.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import model_torch as M  # noqa: E402
from oracle import ray_c, ray_np  # noqa: E402


def check_reference_kat():
    kat = json.load(open(os.path.join(HERE, "reference_kat.json")))
    for case in kat["rotate_yaw"]:
        got = ray_c.rotate_yaw(np.array(case["v"], dtype=np.float32), np.float32(eval(case["angle"], {"pi": np.float32(np.pi)})))
        assert got.tobytes() == np.array(case["want"], dtype=np.float32).tobytes(), (case, got)
    for case in kat["rotate_pitch_round_trip"]:
        a = np.float32(eval(case["angle"], {"pi": np.float32(np.pi)}))
        v = np.array(case["v"], dtype=np.float32)
        got = ray_c.rotate_pitch(ray_c.rotate_pitch(v, a), -a)
        assert np.array_equal(got, v), (case, got)


def make_ray():
    rng = np.random.default_rng(20261018)
    w, h, r, s, picks = 128, 128, 84, 64, 12      # the reference's shipped sizes (model.rs:7-8, ray_sampling.rs:7-8)
    angles = ray_c.get_view_angles(6)             # image_loading.rs:67-80 -> 84 pairs
    n_img = 12
    imgs = rng.random((n_img, h * w, 4)).astype(np.float32)
    idx = np.stack([rng.integers(0, h, r), rng.integers(0, w, r)], 1).astype(np.int64)
    vi = rng.integers(0, n_img, picks).astype(np.int64)
    u = rng.random((r, s)).astype(np.float32)     # unsorted: the sampler sorts (ray_sampling.rs:125)
    _, pts, t, gold = ray_c.get_multiview_batch(imgs, angles, idx, vi, s, u, w, h)
    _, pts_np, t_np, gold_np, _ = ray_np.get_multiview_batch(imgs, angles, idx, vi, s, u, w, h)
    assert pts.tobytes() == pts_np.tobytes() and t.tobytes() == t_np.tobytes() and gold.tobytes() == gold_np.tobytes()
    bsz = r // picks
    dirs = np.concatenate([ray_c.ray_dirs(idx[i * bsz:(i + 1) * bsz], float(angles[vi[i]][0]), float(angles[vi[i]][1]), w, h)
                           for i in range(picks)])
    # deterministic branch u = i/S (ray_sampling.rs:112)
    pts_det, t_det = ray_np.sample_rays(idx[:8], s, angles[5][0], angles[5][1], None, w, h)
    # gold is a pure gather of imgs (dataset.rs:111-114): store the sources' flat indices instead of the images
    np.savez_compressed(os.path.join(HERE, "ray_golden.npz"), w=w, h=h, angles=angles, img_seed=20261018, idx=idx, vi=vi, u=u,
                        points=pts, t=t, dirs=dirs.astype(np.float32), gold=gold, points_det=pts_det, t_det=t_det,
                        det_pose=np.array([angles[5][0], angles[5][1]], dtype=np.float32))


def make_model():
    torch.manual_seed(0)
    out = {}
    # "as_shipped" = the reference's shipped network with the INTENDED [ray, sample] transmittance layout
    # (what the CUDA path computes); the literal graph with the .view scramble of
    # model.rs:241 is frozen beside it as as_shipped_literal_* (oracle-only).
    for name, mcfg, r, s in (("as_shipped", M.replace(M.ModelConfig.as_shipped(), bug_compat_T_view=False), 12, 16),
                             ("ns64", M.ModelConfig(hidden=64), 8, 16)):
        params = M.init_params(mcfg, 7)
        rng = np.random.default_rng(11)
        idx = np.stack([rng.integers(0, 100, r), rng.integers(0, 100, r)], 1).astype(np.int64)
        u = np.sort(rng.random((r, s)).astype(np.float32), axis=1)
        yaw, pitch = np.float32(0.7), np.float32(0.4)
        pts, t = ray_np.sample_rays(idx, s, yaw, pitch, u, 100, 100)
        dirs = ray_np.ray_dirs(idx, yaw, pitch, 100, 100).astype(np.float32)
        gold = rng.random(r * 4).astype(np.float32)
        tr = M.Trainer(mcfg, params, lr=5e-4)
        flat0 = M.flatten_params(params).numpy().copy()
        pix, sig = tr.predict(torch.from_numpy(pts.reshape(-1)), torch.from_numpy(t.reshape(-1)), r, s,
                              torch.from_numpy(dirs) if mcfg.cd else None, literal=False)
        loss = tr.step(pix, torch.from_numpy(gold))
        out.update({f"{name}_params": flat0, f"{name}_points": pts.reshape(-1), f"{name}_t": t.reshape(-1), f"{name}_dirs": dirs,
                    f"{name}_gold": gold, f"{name}_pixels": pix.detach().numpy(), f"{name}_sigma": sig.detach().numpy(),
                    f"{name}_loss": np.float32(loss), f"{name}_grads": tr.grads_flat().numpy(),
                    f"{name}_params_after": tr.params_flat().numpy(), f"{name}_rs": np.array([r, s])})
    lit = M.ModelConfig.as_shipped()
    tr = M.Trainer(lit, M.unflatten_params(lit, torch.from_numpy(out["as_shipped_params"])), lr=5e-4)
    pix, _ = tr.predict(torch.from_numpy(out["as_shipped_points"]), torch.from_numpy(out["as_shipped_t"]), 12, 16, None, literal=True)
    out["as_shipped_literal_pixels"] = pix.detach().numpy()
    out["as_shipped_literal_loss"] = np.float32(tr.step(pix, torch.from_numpy(out["as_shipped_gold"])))
    # compositing alone (model.rs:234-249) with explicit deltas
    rng = np.random.default_rng(3)
    sg = rng.random((6, 16)).astype(np.float32) * 4
    col = rng.random((6, 16, 4)).astype(np.float32)
    dl = rng.random((6, 16)).astype(np.float32) * 0.1
    comp = M.compositing_literal(torch.from_numpy(sg), torch.from_numpy(col), torch.from_numpy(dl), bug_compat_T_view=False)
    comp_bug = M.compositing_literal(torch.from_numpy(sg), torch.from_numpy(col), torch.from_numpy(dl), bug_compat_T_view=True)
    out.update(comp_sigma=sg, comp_colors=col, comp_delta=dl, comp_out=comp.numpy(), comp_out_bug=comp_bug.numpy())
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)


if __name__ == "__main__":
    ray_c.build()
    check_reference_kat()
    make_ray()
    make_model()
    print("golden fixtures written to", HERE)
