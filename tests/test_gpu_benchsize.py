"""Parity at the sizes bench.py measures: the CUDA path through the C ABI against the CPU oracle at
BASELINE configs[1] (4096 rays x 64 samples, 8x256) -- predict, loss, per-layer gradients, Adam --, configs[2]'s forward
(4096 x 192), a hidden-512 training step at 512 rays x 128 samples, the 1 000-step loss curve on the 8x256 network, and the
positional encoder's SFU sin/cos against libm at every octave. Tolerances are the north star's (1e-2 relative for MLP outputs
and pixels, 2 % for loss curves) and are written next to each assert; achieved errors are printed (pytest -s / -rP)."""
import numpy as np
import pytest
import torch

import nerf_rs_b200 as nb
from nerf_rs_b200 import _lib
from oracle import model_torch as M
from tests import gpu_util as G
from tests.test_gpu_mlp import _layer_slices, _train_curves

pytestmark = pytest.mark.gpu


def _model(rays, samples, hidden, seed=0, **kw):
    cfg = nb.default_config(image_w=100, image_h=100, num_rays=rays, num_samples=samples, hidden=hidden, **kw)
    m = nb.NeRF(cfg)
    mcfg = G.model_cfg(cfg)
    params_t = M.init_params(mcfg, seed)
    m.set_weights(M.flatten_params(params_t).numpy())
    pts, t, dirs, gold = G.make_points(rays, samples, seed + 1)
    return m, cfg, mcfg, params_t, pts, t, dirs, gold


def _oracle_step(ocfg, params_t, pts, t, dirs, gold, r, s):
    tr = M.Trainer(ocfg, params_t, lr=5e-4)
    out, sig = tr.predict(torch.from_numpy(pts), torch.from_numpy(t), r, s, torch.from_numpy(dirs), literal=False)
    loss = tr.step(out, torch.from_numpy(gold))
    return out.detach().numpy(), sig.detach().numpy(), loss, tr.grads_flat().numpy()


def _grad_errors(mcfg, g, ref):
    return [float(np.linalg.norm(g[a:b] - ref[a:b]) / np.linalg.norm(ref[a:b])) for a, b in _layer_slices(mcfg)]


def test_cfg1_full_size_predict_loss_gradients_adam():
    """BASELINE configs[1]: 4096 rays x 64 samples = 2048 tiles of the 8x256 network, the exact bench.py shape."""
    r, s = 4096, 64
    m, cfg, mcfg, params_t, pts, t, dirs, gold = _model(r, s, 256)
    w0 = m.get_weights()
    out, sig = m.predict(pts, t, dirs.reshape(-1), train=True)
    loss = nb.Trainer(m, 5e-4).step(out, gold)
    g = m.get_grads()
    assert np.isfinite(out).all() and np.isfinite(sig).all() and np.isfinite(g).all()

    e_out, e_sig, e_loss, e_g = _oracle_step(M.replace(mcfg, emulate_bf16=True, emulate_bf16_grads=True), params_t, pts, t, dirs, gold, r, s)
    f_out, f_sig, f_loss, f_g = _oracle_step(mcfg, params_t, pts, t, dirs, gold, r, s)
    errs = dict(pix_vs_bf16_oracle=G.rel_err(out, e_out), sig_vs_bf16_oracle=G.rel_err(sig, e_sig),
                pix_vs_fp32_oracle=G.rel_err(out, f_out), sig_vs_fp32_oracle=G.rel_err(sig, f_sig),
                loss_vs_fp32_oracle=abs(loss - f_loss) / f_loss)
    print("cfg1 full size:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["pix_vs_bf16_oracle"] < 4e-3 and errs["sig_vs_bf16_oracle"] < 4e-3      # same rounding points: tight
    assert errs["pix_vs_fp32_oracle"] < 1e-2 and errs["sig_vs_fp32_oracle"] < 1e-2      # north-star gate
    assert errs["loss_vs_fp32_oracle"] < 1e-2
    ge, gf = _grad_errors(mcfg, g, e_g), _grad_errors(mcfg, g, f_g)
    print("cfg1 per-layer gradient error (relative L2) vs the bf16-gradient-emulating oracle:", [f"{x:.2e}" for x in ge])
    print("cfg1 per-layer gradient error (relative L2) vs the fp32 oracle:                  ", [f"{x:.2e}" for x in gf])
    assert max(ge) < 5e-3       # measured 4e-4 (fc1) .. 2e-6: at this size the rounding noise averages out
    assert max(gf) < 2e-2       # against pure fp32 autograd (bf16 operands of the gradient GEMMs are the difference): measured 5e-3
    # Adam (model.rs:306-309, 322) on the kernel's own gradient is exact
    p1, _, _ = M.adam_reference(torch.from_numpy(w0), torch.from_numpy(g), torch.zeros(g.size), torch.zeros(g.size), 1)
    assert np.allclose(m.get_weights(), p1.numpy(), rtol=1e-6, atol=1e-7)


def test_cfg2_forward_4096x192():
    """BASELINE configs[2]'s shape (4096 rays x 192 samples = 6144 tiles): pixels and densities against both oracles."""
    r, s = 4096, 192
    m, cfg, mcfg, params_t, pts, t, dirs, gold = _model(r, s, 256)
    out, sig = m.predict(pts, t, dirs.reshape(-1), train=False)
    e_out, e_sig = G.oracle_predict(M.replace(mcfg, emulate_bf16=True), params_t, pts, t, dirs, r, s)
    f_out, f_sig = G.oracle_predict(mcfg, params_t, pts, t, dirs, r, s)
    errs = (G.rel_err(out, e_out.detach().numpy()), G.rel_err(sig, e_sig.detach().numpy()),
            G.rel_err(out, f_out.detach().numpy()), G.rel_err(sig, f_sig.detach().numpy()))
    print("cfg2 forward: pixels / sigma vs bf16 oracle, vs fp32 oracle:", [f"{x:.2e}" for x in errs])
    assert errs[0] < 4e-3 and errs[1] < 4e-3
    assert errs[2] < 1e-2 and errs[3] < 1e-2


def test_hidden512_step_512x128():
    """BASELINE configs[4]'s width at 512 rays x 128 samples (512 tiles): loss and per-layer gradients."""
    r, s = 512, 128
    m, cfg, mcfg, params_t, pts, t, dirs, gold = _model(r, s, 512)
    out, _ = m.predict(pts, t, dirs.reshape(-1), train=True)
    loss = nb.Trainer(m, 5e-4).step(out, gold)
    g = m.get_grads()
    e_out, _, e_loss, e_g = _oracle_step(M.replace(mcfg, emulate_bf16=True, emulate_bf16_grads=True), params_t, pts, t, dirs, gold, r, s)
    f_out, _, f_loss, f_g = _oracle_step(mcfg, params_t, pts, t, dirs, gold, r, s)
    ge, gf = _grad_errors(mcfg, g, e_g), _grad_errors(mcfg, g, f_g)
    print("hidden 512: pixels vs fp32 oracle", f"{G.rel_err(out, f_out):.2e}", "gradients vs bf16 oracle", [f"{x:.2e}" for x in ge],
          "vs fp32 oracle", [f"{x:.2e}" for x in gf])
    assert G.rel_err(out, e_out) < 4e-3 and G.rel_err(out, f_out) < 1e-2
    assert abs(loss - f_loss) < 1e-2 * f_loss
    assert max(ge) < 5e-3 and max(gf) < 2e-2      # measured 1e-3 / 7e-3


def test_loss_curve_1k_steps_hidden256_within_2_percent():
    """North-star gate on the network every BASELINE config uses (8x256, L = 10/4): 1 000 steps of 256 fresh rays x 64 samples,
    CUDA path vs the fp32 torch oracle from the same initial weights; curves compared after a 50-step moving average."""
    cfg = nb.default_config(image_w=100, image_h=100, mlp_impl=_lib.MLP_TCGEN05, hidden=256, num_rays=256, num_samples=64)
    ref, got = _train_curves(cfg, False, 1000, 100)
    k = np.ones(50) / 50
    rs, gs = np.convolve(ref, k, "valid"), np.convolve(got, k, "valid")
    print("1k-step curve at hidden 256: first / last smoothed loss", float(gs[0]), float(gs[-1]),
          "max |cuda - oracle| / max(oracle)", float(np.abs(gs - rs).max() / rs.max()))
    assert gs[-1] < 0.5 * gs[0]                            # it actually trains
    assert np.abs(gs - rs).max() <= 0.02 * rs.max()         # within 2 % of the curve's scale everywhere
    assert np.all(np.abs(gs - rs) <= 0.02 * rs + 2e-5)      # and within 2 % pointwise (+ bf16 noise floor)


def test_positional_encoder_sfu_sincos_vs_libm():
    """The fused prologue evaluates every feature as sin.approx / cos.approx of the exactly scaled argument 2^k x (|arg| up to
    ~1 000 at octave 9). Read the encoded panel the training forward saves and compare each octave with float64 libm: the
    deviation must stay within bf16 rounding (half an ulp at 1.0 = 2^-9) plus the SFU's absolute error."""
    r, s = 64, 64
    m, cfg, mcfg, params_t, pts, t, dirs, gold = _model(r, s, 256)
    m.predict(pts, t, dirs.reshape(-1), train=True)
    n_tiles = r * s // 128
    got = np.concatenate([G.decode_panel(m.debug_read_panel(0, tile, 0)) for tile in range(n_tiles)])[:, :63]   # activation slot 0 = X
    x = pts.reshape(-1, 3).astype(np.float64)
    assert np.array_equal(got[:, :3], G.bf16_round(pts.reshape(-1, 3)))
    worst = []
    for k in range(10):
        arg = x * 2.0 ** k
        want = np.concatenate([np.sin(arg), np.cos(arg)], axis=1)
        err = np.abs(got[:, 3 + 6 * k:9 + 6 * k] - want).max()
        flips = float(np.mean(got[:, 3 + 6 * k:9 + 6 * k] != G.bf16_round(want.astype(np.float32))))
        worst.append((k, float(np.abs(arg).max()), float(err), flips))
    print("posenc octave, max |arg|, max abs error vs libm, fraction of bf16 values that differ from bf16(libm):")
    for w in worst:
        print("   k=%d  |arg|<=%.0f  err=%.2e  flips=%.4f" % w)
    for k, amax, err, flips in worst:
        assert err <= 2.0 ** -9 + 4e-4, (k, err)     # bf16 half-ulp below 1.0 + SFU absolute error (~1e-7 * |arg|)
        assert flips < 0.12, (k, flips)              # a flip = the SFU error crossed a bf16 rounding boundary
