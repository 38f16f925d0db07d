"""Logging projections (src/logging.rs, src/display.rs:96-110): closed forms for the CPU restatement and
GPU-vs-oracle parity (counts, occupancy maps and last-writer-wins maps exact; f64 density sums to 1e-12)."""
import numpy as np
import pytest

import nerf_rs_b200 as nb
from oracle import metrics_np as MN


def test_oracle_closed_forms():
    assert MN.prediction_array_as_u32([1.0, 0.5, 0.0, 1.0]) == (255 << 16) | (127 << 8)            # (c*255) as u8 truncates
    assert MN.prediction_array_as_u32([2.0, -1.0, float("nan"), 1.0]) == 255 << 16                 # saturating cast, NaN -> 0
    sx, sy = MN.log_screen_coords([[3, 1], [3, 2]], 8, 8)
    assert sx[3] == 2 and sy[1] == 1 and sy[2] == 1 and sx.sum() == 2                              # `[x, y]` binds element 0 to x
    t = MN.log_query_distances([[0.0, 0.0019, 0.002, 1.9999]])
    assert t[0] == 2 and t[1] == 1 and t[999] == 1 and t.sum() == 4
    yx, zx, yz = MN.log_query_points_as_maps([[[0.0, 0.0, 0.0]], [[5.0, 5.0, 5.0]]])
    assert yx[50, 50] == 0xFFFFFF and zx[25, 50] == 0xFFFFFF and yz[50, 25] == 0xFFFFFF
    assert yx.reshape(-1)[9999] == 0xFFFFFF                                                        # .min(9999) for the far point
    pts = [[[0.0, 0.0, 0.0], [0.001, 0.001, 0.001]]]
    dyx, _, _ = MN.log_density_maps(pts, [[0.25, 0.75]])
    assert dyx[50, 50] == MN.prediction_array_as_u32([0.75] * 3 + [1.0])                           # the later sample wins
    bx, by, bz = MN.log_densities(pts, [[0.25, 0.75]])
    assert bx[500] == by[500] == bz[500] == 1.0
    bb = MN.draw_predictions([[1, 2], [1, 2]], [[0.1, 0.2, 0.3, 0.4], [1.0, 1.0, 1.0, 0.0]], 4, 3)
    assert bb[1, 2] == 0xFFFFFF and np.count_nonzero(bb) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("hidden,impl", [(64, 1), (128, 0)])
def test_device_metrics_match_the_restatement(hidden, impl):
    from oracle import model_torch as M
    from tests import gpu_util as G
    rng = np.random.default_rng(3)
    w = h = 100
    r, s, v = 256, 32, 4
    cfg = nb.default_config(image_w=w, image_h=h, num_rays=r, num_samples=s, hidden=hidden, mlp_impl=impl)
    m = nb.NeRF(cfg)
    m.set_weights(M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy())
    m.set_images(rng.random((v, w * h, 4), dtype=np.float32))
    m.set_view_angles(nb.get_view_angles(6)[:v])
    idx = np.stack([rng.integers(0, h, r), rng.integers(0, w, r)], 1).astype(np.int64)
    idx[1] = idx[0]                                    # a duplicated pixel: the later ray must win the back buffer
    vi = rng.integers(0, v, v).astype(np.int64)
    u = rng.random((r, s)).astype(np.float32)
    b = m.get_batch(idx, vi, v, u, True, 0)
    if impl == 0:   # fused sampling: without a host read-back of the points they never reach HBM and the kernel rebuilds them
        m.get_batch(idx, vi, v, u, True, 0, want=("t",))
    out, sig = m.predict(train=False)
    got = m.log_metrics()
    sx, sy = MN.log_screen_coords(idx, w, h)
    assert np.array_equal(got["screen_x"], sx) and np.array_equal(got["screen_y"], sy)
    assert np.array_equal(got["t"], MN.log_query_distances(b["t"]))
    for k, want in zip(("world_yx", "world_zx", "world_yz"), MN.log_query_points_as_maps(b["points"])):
        assert np.array_equal(got[k], want), k
    for k, want in zip(("density_yx", "density_zx", "density_yz"), MN.log_density_maps(b["points"], sig)):
        assert np.array_equal(got[k], want), k
    for k, want in zip(("density_x", "density_y", "density_z"), MN.log_densities(b["points"], sig)):
        assert np.allclose(got[k], want, rtol=1e-12, atol=1e-12), k
    assert np.array_equal(got["prediction"], MN.draw_predictions(idx, out, w, h))


@pytest.mark.gpu
def test_metrics_need_a_batch_and_a_prediction():
    m = nb.NeRF(nb.default_config(image_w=64, image_h=64, num_rays=64, num_samples=16, hidden=64))
    with pytest.raises(nb.NerfError):
        m.log_metrics()
    m.set_view_angles(nb.get_view_angles(6)[:4])
    m.get_batch(None, None, 4, None, True, 1, want=())
    with pytest.raises(nb.NerfError):
        m.log_metrics()                                # densities requested before predict
    assert m.log_metrics(densities=False, prediction=False)["t"].sum() == 64 * 16
    rng = np.random.default_rng(0)
    m.set_images(rng.random((4, 64 * 64, 4), dtype=np.float32))
    m.train_iter(3)                                    # fused iteration: the batch, its densities and pixels stay resident
    got = m.log_metrics()
    assert got["t"].sum() == 64 * 16 and np.count_nonzero(got["prediction"]) > 0 and np.isfinite(got["density_x"]).all()
