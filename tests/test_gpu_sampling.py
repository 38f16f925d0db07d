"""K-sample on the GPU vs the CPU oracle: ray indices and sample positions bit-exact
(north-star parity gate), through the C ABI (nerf_get_batch)."""
import numpy as np
import pytest

import nerf_rs_b200 as nb
from oracle import ray_c, ray_np

pytestmark = pytest.mark.gpu


def _model(w, h, r, s, **kw):
    cfg = nb.default_config(image_w=w, image_h=h, num_rays=r, num_samples=s, hidden=64, mlp_impl=1, **kw)
    return nb.NeRF(cfg)


@pytest.mark.parametrize("w,h,r,s,v", [(128, 128, 84, 64, 84), (100, 100, 1024, 64, 64), (800, 800, 4096, 192, 64), (16, 12, 12, 8, 4)])
def test_get_batch_bit_exact(w, h, r, s, v):
    rng = np.random.default_rng(5)
    m = _model(w, h, r, s)
    angles = nb.get_view_angles(6)
    assert angles.tobytes() == ray_c.get_view_angles(6).tobytes()   # image_loading.rs:67-80
    nv = min(v, 84)
    imgs = rng.random((nv, h * w, 4)).astype(np.float32)
    m.set_images(imgs)
    m.set_view_angles(angles)
    idx = np.stack([rng.integers(0, h, r), rng.integers(0, w, r)], 1).astype(np.int64)
    picks = v
    vi = rng.integers(0, nv, picks).astype(np.int64)
    u = rng.random((r, s)).astype(np.float32)          # UNSORTED: the device must sort like :125
    b = m.get_batch(idx, vi, picks, u, True, 0)
    _, pc, tc, gc = ray_c.get_multiview_batch(imgs, angles, idx, vi, s, u, w, h)
    assert b["t"].tobytes() == tc.tobytes()
    assert b["points"].tobytes() == pc.tobytes()
    assert b["gold"].tobytes() == gc.tobytes()
    assert np.array_equal(b["indices"], idx)
    bsz = r // picks
    dirs = np.concatenate([ray_c.ray_dirs(idx[i * bsz:(i + 1) * bsz], float(angles[vi[i]][0]), float(angles[vi[i]][1]), w, h)
                           for i in range(picks)])
    assert b["dirs"].tobytes() == dirs.tobytes()
    # randomize = false -> t = 2 i / S (ray_sampling.rs:112)
    b2 = m.get_batch(idx, vi, picks, None, False, 0, want=("points", "t"))
    _, pc2, tc2, _ = ray_c.get_multiview_batch(imgs, angles, idx, vi, s, None, w, h)
    assert b2["t"].tobytes() == tc2.tobytes() and b2["points"].tobytes() == pc2.tobytes()


def test_uneven_split_is_an_error_not_a_panic():
    m = _model(16, 16, 12, 8)
    m.set_view_angles(nb.get_view_angles(2))
    with pytest.raises(nb.NerfError):   # dataset.rs:73-81 assert
        m.get_batch(None, None, 5, None, True, 0)


def test_philox_picks_and_jitter_match_oracle():
    w, h, r, s = 100, 100, 256, 64
    m = _model(w, h, r, s)
    angles = nb.get_view_angles(6)
    m.set_view_angles(angles)
    seed = 0x1234ABCD5678
    b = m.get_batch(None, None, 4, None, True, seed)
    y = np.minimum((ray_c.philox_uniform(seed, 0, 0, r) * np.float32(h)).astype(np.int64), h - 1)
    x = np.minimum((ray_c.philox_uniform(seed, 1, 0, r) * np.float32(w)).astype(np.int64), w - 1)
    assert np.array_equal(b["indices"], np.stack([y, x], 1))
    vi = np.minimum((ray_c.philox_uniform(seed, 2, 0, 4) * np.float32(84)).astype(np.int64), 83)
    u = ray_c.philox_uniform(seed, 3, 0, r * s).reshape(r, s)
    _, pc, tc, _ = ray_c.get_multiview_batch(np.zeros((84, 1, 4), np.float32), angles, b["indices"], vi, s, u, w, h) if False else (None, None, None, None)
    pts, ts = [], []
    bsz = r // 4
    for i in range(4):
        p, t = ray_c.sample_rays(b["indices"][i * bsz:(i + 1) * bsz], s, float(angles[vi[i]][0]), float(angles[vi[i]][1]), u[i * bsz:(i + 1) * bsz], w, h)
        pts.append(p); ts.append(t)
    assert b["t"].tobytes() == np.concatenate(ts).tobytes()
    assert b["points"].tobytes() == np.concatenate(pts).tobytes()


def test_stratified_mode():
    w, h, r, s = 100, 100, 64, 64
    m = _model(w, h, r, s, depth_mode=1)
    angles = nb.get_view_angles(6)
    m.set_view_angles(angles)
    rng = np.random.default_rng(1)
    idx = np.stack([rng.integers(0, h, r), rng.integers(0, w, r)], 1).astype(np.int64)
    u = rng.random((r, s)).astype(np.float32)
    b = m.get_batch(idx, np.array([3]), 1, u, True, 0)
    p, t = ray_np.sample_rays(idx, s, angles[3][0], angles[3][1], u, w, h, mode="stratified")
    assert b["t"].tobytes() == t.tobytes() and b["points"].tobytes() == p.tobytes()
    assert np.all(np.diff(b["t"], axis=1) >= 0)


def test_host_pose_matches_oracle_matrices():
    import ctypes
    lib = nb.load()
    yaw = np.zeros((3, 4), np.float32); pit = np.zeros((3, 3), np.float32); off = ctypes.c_float()
    for a, b in [(0.0, 0.0), (np.pi / 2, -np.pi / 2), (1.234, 2.5)]:
        lib.nerf_debug_host_pose(a, b, yaw.ctypes.data, pit.ctypes.data, ctypes.byref(off))
        assert yaw.tobytes() == ray_np.yaw_matrix(np.float32(a)).tobytes()
        assert pit.tobytes() == ray_c.pitch_matrix(float(np.float32(b))).tobytes()
    assert np.float32(off.value) == np.float32(0.028867517)
