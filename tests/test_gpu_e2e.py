"""End-to-end surface: device-resident training iterations, render, checkpoint round trip."""
import os
import tempfile

import numpy as np
import pytest
import torch

import nerf_rs_b200 as nb
from nerf_rs_b200 import _lib
from oracle import model_torch as M
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def _sphere_images(v, w, h):
    """Synthetic 'sphere silhouette' targets (the reference's own synthetic target, dataset.rs:50-57)."""
    yy, xx = np.mgrid[0:h, 0:w]
    d = np.hypot((xx - w / 2) / (w / 2), (yy - h / 2) / (h / 2))
    img = np.zeros((h * w, 4), np.float32)
    inside = (d < 0.5).reshape(-1)
    img[inside] = [0.8, 0.4, 0.2, 1.0]
    return np.repeat(img[None], v, 0)


def test_train_iter_reduces_loss_and_is_deterministic():
    cfg = nb.default_config(image_w=64, image_h=64, num_rays=256, num_samples=32, hidden=128)
    losses = []
    for rep in range(2):
        m = nb.NeRF(cfg)
        mcfg = G.model_cfg(cfg)
        m.set_weights(M.flatten_params(M.init_params(mcfg, 0)).numpy())
        m.set_images(_sphere_images(4, 64, 64))
        m.set_view_angles(nb.get_view_angles(6)[:4])
        ls = []
        for it in range(60):
            m.train_iter(1000 + it)
            ls.append(m.last_loss())
        losses.append(ls)
        assert m.launch_count > 0
    assert np.isfinite(losses[0]).all()
    assert np.mean(losses[0][-10:]) < 0.7 * np.mean(losses[0][:5])
    # Same seeds -> same curve up to the order of the fp32 weight-gradient atomics: the 1e-7 differences that leaves in
    # the weights occasionally flip a bf16 rounding of an activation, which shows up as isolated ~1e-3 relative spikes
    # in single losses (tools/det_check.py: 3e-4 .. 1.6e-3 over 60 steps) while the forward itself is bit-reproducible
    # (tests/test_gpu_mlp.py::test_step_is_reproducible).
    assert np.allclose(losses[0], losses[1], rtol=3e-2)   # (20-50x the observed spikes; a race shows up as a curve that diverges)
    assert np.allclose(losses[0][:4], losses[1][:4], rtol=1e-5)


def test_micro_batched_step_equals_single_launch():
    kw = dict(image_w=100, image_h=100, num_rays=64, num_samples=32, hidden=128)
    pts, t, dirs, gold = G.make_points(64, 32, 2)
    grads = []
    for chunk in (0, 16):
        cfg = nb.default_config(max_rays_per_launch=chunk, **kw)
        m = nb.NeRF(cfg)
        m.set_weights(M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy())
        out, _ = m.predict(pts, t, dirs.reshape(-1), train=True)
        nb.Trainer(m).step(out, gold)
        grads.append(m.get_grads())
    assert np.allclose(grads[0], grads[1], rtol=1e-3, atol=1e-6 * np.abs(grads[0]).max())


def test_render_matches_predict_and_packs_0rgb():
    cfg = nb.default_config(image_w=32, image_h=24, num_rays=96, num_samples=32, hidden=64)
    m = nb.NeRF(cfg)
    mcfg = G.model_cfg(cfg)
    params_t = M.init_params(mcfg, 0)
    m.set_weights(M.flatten_params(params_t).numpy())
    yaw, pitch = 0.3, 0.2
    rgba, packed = m.render(yaw, pitch, 2, 20, randomize=False, packed=True)
    assert rgba.shape == (18, 32, 4) and packed.shape == (18, 32)
    # oracle: every pixel of rows [2,20), deterministic depths (display.rs:58-62, ray_sampling.rs:112)
    from oracle import ray_np
    idx = np.array([[y, x] for y in range(2, 20) for x in range(32)], dtype=np.int64)
    pts, t = ray_np.sample_rays(idx, 32, np.float32(yaw), np.float32(pitch), None, 32, 24)
    dirs = ray_np.ray_dirs(idx, np.float32(yaw), np.float32(pitch), 32, 24)
    want, _ = M.predict(M.replace(mcfg, emulate_bf16=True), params_t, torch.from_numpy(pts.reshape(-1)), torch.from_numpy(t.reshape(-1)),
                        idx.shape[0], 32, torch.from_numpy(dirs), literal=False)
    assert G.rel_err(rgba.reshape(-1, 4), want.detach().numpy()) < 5e-3
    c = np.clip(rgba[..., :3] * np.float32(255.0), 0, 255).astype(np.uint32)     # (c*255) as u8 (display.rs:46-52)
    assert np.array_equal(packed, (c[..., 0] << 16) | (c[..., 1] << 8) | c[..., 2])


def test_checkpoint_round_trip():
    cfg = nb.default_config(image_w=100, image_h=100, num_rays=16, num_samples=32, hidden=64)
    pts, t, dirs, gold = G.make_points(16, 32, 4)
    m = nb.NeRF(cfg)
    for _ in range(3):
        out, _ = m.predict(pts, t, dirs.reshape(-1), train=True)
        nb.Trainer(m).step(out, gold)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ckpt.npz")
        m.save(path)
        m2 = nb.NeRF(cfg)
        m2.load(path)
    for mm in (m, m2):
        out, _ = mm.predict(pts, t, dirs.reshape(-1), train=True)
        nb.Trainer(mm).step(out, gold)
    assert np.allclose(m.get_weights(), m2.get_weights(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("s", [64, 192])
def test_fused_sampling_in_the_mlp_prologue_is_bit_identical(s):
    """North star: sample positions are generated inside the MLP kernel's prologue and never written to HBM.
    The fused path must give exactly the pixels / densities of the path that reads the sampler's points."""
    cfg = nb.default_config(image_w=100, image_h=100, num_rays=96, num_samples=s, hidden=256)
    m = nb.NeRF(cfg)
    mcfg = G.model_cfg(cfg)
    m.set_weights(M.flatten_params(M.init_params(mcfg, 2)).numpy())
    rng = np.random.default_rng(3)
    m.set_images(rng.random((6, 100 * 100, 4)).astype(np.float32))
    m.set_view_angles(nb.get_view_angles(6)[10:16])
    idx = np.stack([rng.integers(0, 100, 96), rng.integers(0, 100, 96)], 1).astype(np.int64)
    vi = rng.integers(0, 6, 6).astype(np.int64)
    u = rng.random((96, s)).astype(np.float32)
    b = m.get_batch(idx, vi, 6, u, True, 0, want=("points", "t", "dirs"))     # points requested -> written to HBM, MLP reads them
    out_a, sig_a = m.predict(train=False)
    m.get_batch(idx, vi, 6, u, True, 0, want=())                              # nothing requested -> fused: no points in HBM
    out_b, sig_b = m.predict(train=False)
    assert np.array_equal(sig_a, sig_b) and np.array_equal(out_a, out_b)
    out_c, sig_c = m.predict(b["points"].reshape(-1), b["t"].reshape(-1), b["dirs"].reshape(-1), train=False)   # literal signature
    assert np.array_equal(sig_a, sig_c) and np.array_equal(out_a, out_c)


def test_render_sharded_single_rank_equals_render():
    """Without a communicator the sharded render is one band: identical to nerf_render of the whole frame."""
    cfg = nb.default_config(image_w=48, image_h=40, num_rays=512, num_samples=24, hidden=64)
    m = nb.NeRF(cfg)
    full, pk = m.render(0.4, 0.3, packed=True)
    sh, pks = m.render_sharded(0.4, 0.3, packed=True)
    assert np.array_equal(full, sh) and np.array_equal(pk, pks)
    part = m.render(0.4, 0.3, 7, 19)                 # an unaligned row band equals the same rows of the full frame
    assert np.array_equal(part, full[7:19])


def test_micro_batched_train_iter_equals_single_launch():
    """nerf_train_iter on a batch larger than the saved-activation budget runs forward -> compositing backward -> dgrad ->
    wgrad per micro-batch (no full forward first): same loss and the same gradient as the single-launch iteration."""
    res = []
    for chunk in (0, 64):
        cfg = nb.default_config(image_w=64, image_h=64, num_rays=256, num_samples=32, hidden=128, max_rays_per_launch=chunk)
        m = nb.NeRF(cfg)
        m.set_weights(M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy())
        m.set_images(_sphere_images(4, 64, 64))
        m.set_view_angles(nb.get_view_angles(6)[:4])
        m.train_iter(7)
        res.append((m.last_loss(), m.get_grads()))
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[0][0])
    assert np.allclose(res[0][1], res[1][1], rtol=1e-3, atol=1e-6 * np.abs(res[0][1]).max())


def test_overlapped_host_copy_equals_plain_copy():
    """nerf_predict_points streams the points in 8 chunks on a second stream while the forward kernel already runs (its prologue
    polls the copy engine's counter): same bits as copy-then-run, also when rays and tiles do not divide into the chunks."""
    import os
    cfg = nb.default_config(image_w=64, image_h=64, num_rays=100, num_samples=50, hidden=128)   # 5000 samples: 39.06 tiles, 12.5 rays per chunk
    pts, t, dirs, gold = G.make_points(100, 50, 9)
    res = []
    for plain in (False, True):
        if plain:
            os.environ["NERF_B200_NO_H2D_OVERLAP"] = "1"
        try:
            m = nb.NeRF(cfg)
            m.set_weights(M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy())
            outs = [m.predict(pts * np.float32(1 + 0.01 * k), t, dirs.reshape(-1), train=True) for k in range(3)]   # back-to-back calls reuse the counter
            loss = nb.Trainer(m).step(outs[-1][0], gold)
            res.append((outs, loss, m.get_grads()))
        finally:
            os.environ.pop("NERF_B200_NO_H2D_OVERLAP", None)
    for (oa, sa), (ob, sb) in zip(res[0][0], res[1][0]):
        assert np.array_equal(oa, ob) and np.array_equal(sa, sb)
    assert res[0][1] == res[1][1]
    assert np.allclose(res[0][2], res[1][2], rtol=1e-3, atol=1e-6 * np.abs(res[1][2]).max())   # (wgrad atomics are unordered)


def test_device_resident_prediction_equals_eager_predict():
    """predict(lazy=True) keeps the prediction on the device like the reference's Tensor (main.rs:58 -> :72): compositing is left
    to the step (or to the first read), Trainer.step consumes the handle, and pixels, densities, loss, gradients and updated
    weights are the eager path's bit for bit (deterministic weight gradients, so the comparison can be exact)."""
    cfg = nb.default_config(image_w=64, image_h=64, num_rays=100, num_samples=50, hidden=128, deterministic_grads=1)
    pts, t, dirs, gold = G.make_points(100, 50, 9)
    w0 = M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy()

    def run(lazy, read_before_step):
        m = nb.NeRF(cfg)
        m.set_weights(w0)
        tr = nb.Trainer(m)
        if lazy:
            pred, none = m.predict(pts, t, dirs.reshape(-1), train=True, lazy=True)
            assert none is None and isinstance(pred, nb.Prediction) and pred.shape == (100, 4)
            px0 = pred.numpy() if read_before_step else None
            loss = tr.step(pred, gold)
            px, sig = np.asarray(pred), pred.densities()      # after the step: the compositing backward rewrote the same pixels
            if px0 is not None:
                assert np.allclose(px0, px, rtol=0, atol=1e-6)
            m.predict(pts, t, dirs.reshape(-1), train=True, lazy=True)
            with pytest.raises(nb.NerfError):
                pred.numpy()                                    # stale handle
            with pytest.raises(nb.NerfError):
                tr.step(pred, gold)
        else:
            px, sig = m.predict(pts, t, dirs.reshape(-1), train=True)
            loss = tr.step(px, gold)
        return px, sig, loss, m.get_grads(), m.get_weights()

    eager = run(False, False)
    for read_first in (False, True):
        lazy = run(True, read_first)
        assert np.allclose(lazy[0], eager[0], rtol=0, atol=1e-6) and np.array_equal(lazy[1], eager[1])
        assert lazy[2] == eager[2]
        assert np.array_equal(lazy[3], eager[3]) and np.array_equal(lazy[4], eager[4])
    # resident batch: a deferred prediction is also what log_metrics' density / prediction maps read
    m = nb.NeRF(nb.default_config(image_w=64, image_h=64, num_rays=128, num_samples=32, hidden=64))
    m.set_images(_sphere_images(4, 64, 64))
    m.set_view_angles(nb.get_view_angles(6)[:4])
    m.get_batch(None, None, 4, None, True, 3, want=())
    pred, _ = m.predict(train=True, lazy=True)
    logs = m.log_metrics(True, True)
    px = pred.numpy()
    m2 = nb.NeRF(nb.default_config(image_w=64, image_h=64, num_rays=128, num_samples=32, hidden=64))
    m2.set_weights(m.get_weights())
    m2.set_images(_sphere_images(4, 64, 64))
    m2.set_view_angles(nb.get_view_angles(6)[:4])
    m2.get_batch(None, None, 4, None, True, 3, want=())
    px2, _ = m2.predict(train=True)
    logs2 = m2.log_metrics(True, True)
    assert np.array_equal(px, px2) and np.array_equal(logs["prediction"], logs2["prediction"]) and np.array_equal(logs["density_yx"], logs2["density_yx"])


@pytest.mark.parametrize("chunk", [0, 128])
def test_overlapped_paths_equal_the_single_stream_run(chunk):
    """tests/pipeline_worker.py (chunk > 0: micro-batched steps, which re-run the forward after the compositing backward and
    must therefore NOT release the batch inputs early): a mixed sequence of device-resident / eager predictions, host-index batches and fused iterations
    with pinned buffers rewritten between calls gives the same bits whether copies, sampler and the loss read overlap the
    previous step (as shipped) or everything runs on one stream and every step is waited for."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for serial in (False, True):
        env = dict(os.environ, PIPELINE_WORKER_CHUNK=str(chunk))
        if serial:
            env.update(NERF_B200_STEP_SYNC="1", NERF_B200_NO_SAMPLER_OVERLAP="1", NERF_B200_NO_H2D_OVERLAP="1")
        p = subprocess.run([sys.executable, os.path.join(root, "tests", "pipeline_worker.py")], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
        line = [l for l in p.stdout.splitlines() if l.startswith("DIGEST")][-1]
        print(("serial   " if serial else "overlap  ") + line)
        outs.append(line)
    assert outs[0] == outs[1]


def test_full_size_properties():
    """BASELINE configs[1] at full size (800x800 views, 4096 rays x 64 samples, W=256): size-independent properties instead of
    a CPU comparison -- (1) the fused-sampling forward equals the forward on the sampler's own points, bit for bit;
    (2) a micro-batched step gives the single-launch gradient; (3) compositing is linear in the colours; (4) the same seeds
    give the same first losses."""
    rng = np.random.default_rng(5)
    cfg = nb.default_config()
    assert (cfg.num_rays, cfg.num_samples, cfg.hidden, cfg.image_w) == (4096, 64, 256, 800)
    w0 = M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy()
    angles = nb.get_view_angles(6)
    imgs = rng.random((8, 800 * 800, 4), dtype=np.float32)

    def fresh(**kw):
        m = nb.NeRF(nb.default_config(**kw))
        m.set_weights(w0)
        m.set_images(imgs)
        m.set_view_angles(angles[:8])
        return m

    m = fresh()
    b = m.get_batch(None, None, 8, None, True, 11)                      # Philox picks + jitter; points read back
    out_pts, sig_pts = m.predict(b["points"].reshape(-1), b["t"].reshape(-1), b["dirs"].reshape(-1), train=False)
    m.get_batch(b["indices"], None, 8, None, True, 11, want=())         # same batch, points never written: fused prologue
    # (view picks are Philox-drawn again from the same seed; the explicit pixel indices reproduce the rays)
    out_fused, sig_fused = m.predict(train=False)
    assert np.array_equal(out_pts, out_fused) and np.array_equal(sig_pts, sig_fused)

    grads, losses = [], []
    for chunk in (0, 1024):
        mm = fresh(max_rays_per_launch=chunk)
        ls = []
        for it in range(3):
            mm.train_iter(100 + it)
            ls.append(mm.last_loss())
        losses.append(ls)
        grads.append(mm.get_grads())
    assert np.allclose(losses[0][0], losses[1][0], rtol=1e-6)            # same batch, same weights: same first loss
    assert np.allclose(losses[0], losses[1], rtol=2e-3)
    assert np.all(np.isfinite(grads[0])) and np.abs(grads[0]).max() > 0

    sig = rng.random((4096, 64), dtype=np.float32) * 3
    dl = np.full((4096, 64), 2.0 / 64, np.float32)
    ca, cb = rng.random((2, 4096, 64, 4), dtype=np.float32)
    lhs = nb.compositing(m, sig, ca + cb, dl)
    rhs = nb.compositing(m, sig, ca, dl) + nb.compositing(m, sig, cb, dl)
    assert np.allclose(lhs, rhs, rtol=1e-5, atol=1e-6)
    assert np.all(nb.compositing(m, sig, np.ones_like(ca), dl)[:, 0] <= 1.0 + 1e-6)   # weights sum to at most 1 (opacity)
