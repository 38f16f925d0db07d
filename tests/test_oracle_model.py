"""Closed-form checks for the tensor-side oracle: the reference pins
none of these values itself, so the restatement is checked against analytic results."""
import math

import numpy as np
import torch

from oracle import model_torch as M
from oracle import ray_np


def test_param_counts():
    assert M.ModelConfig().num_params() == 530181            # BASELINE.md section 3, W=256
    assert M.ModelConfig(hidden=512).num_params() == 2043397  # W=512
    assert M.ModelConfig.as_shipped().num_params() == 76455   # model.rs as shipped
    dims = M.ModelConfig().layer_dims()
    assert dims[0] == (63, 256) and dims[5] == (319, 256) and dims[7] == (256, 257)
    assert dims[8] == (283, 128) and dims[9] == (128, 4)


def test_flatten_round_trip():
    cfg = M.ModelConfig(hidden=64)
    p = M.init_params(cfg, 0)
    flat = M.flatten_params(p)
    assert flat.numel() == cfg.num_params()
    q = M.unflatten_params(cfg, flat)
    for (a, b), (c, d) in zip(p, q):
        assert torch.equal(a, c) and torch.equal(b, d)


def test_deltas():
    t = torch.tensor([[0.0, 0.5, 1.5], [0.25, 0.5, 0.75]])
    d = M.deltas_from_t(t)
    assert torch.equal(d, torch.tensor([[0.5, 1.0, 0.5], [0.25, 0.25, 1.25]]))


def test_compositing_closed_forms():
    torch.manual_seed(0)
    r, s = 7, 64
    t = torch.sort(torch.rand(r, s) * 2, dim=1).values
    delta = M.deltas_from_t(t)
    col = torch.ones(r, s, 4)
    # sigma == 0 -> out == 0
    out = M.compositing_literal(torch.zeros(r, s), col, delta)
    assert torch.equal(out, torch.zeros(r, 4))
    # sigma == k, col == 1 -> 1 - exp(-k (T_FAR - t0)): delta telescopes
    k = 1.7
    out = M.compositing_literal(torch.full((r, s), k), col, delta)
    want = 1 - torch.exp(-k * (M.T_FAR - t[:, 0]))
    assert torch.allclose(out, want[:, None].expand(r, 4), atol=2e-6)
    # sum of weights = 1 - exp(-sum sigma delta)
    sig = torch.rand(r, s) * 3
    _, w = M.compositing(sig, col, delta)
    assert torch.allclose(w.sum(1), 1 - torch.exp(-(sig * delta).sum(1)), atol=2e-6)


def test_literal_graph_equals_scan_form():
    torch.manual_seed(1)
    r, s = 12, 64
    sig = torch.randn(r, s)  # raw sigma may be negative (no activation, model.rs:168-171)
    t = torch.sort(torch.rand(r, s) * 2, dim=1).values
    delta = M.deltas_from_t(t)
    col = torch.rand(r, s, 4)
    a = M.compositing_literal(sig, col, delta)
    b, _ = M.compositing(sig, col, delta)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


def test_T_view_scramble_is_what_survey_says():
    # T_ref[r,i] = T_true[(r*S+i) mod R, (r*S+i) div R]; 3x2 example
    r, s = 3, 2
    true = torch.arange(6.0).view(r, s)                   # [[0,1],[2,3],[4,5]]
    stacked = true.t().contiguous()                       # what stack(dim 0) holds: [S,R]
    assert torch.equal(stacked.view(r, s), torch.tensor([[0.0, 2.0], [4.0, 1.0], [3.0, 5.0]]))
    torch.manual_seed(2)
    R, S = 6, 8
    sig, delta, col = torch.rand(R, S), torch.rand(R, S), torch.rand(R, S, 4)
    good = M.compositing_literal(sig, col, delta, False)
    bad = M.compositing_literal(sig, col, delta, True)
    assert not torch.allclose(good, bad)
    a = (sig * delta).neg().exp()
    T = torch.cumprod(torch.cat([torch.ones(R, 1), a[:, :-1]], 1), 1)
    flat = torch.arange(R * S)
    Tref = T[flat % R, flat // R].view(R, S)
    want = ((Tref * (1 - a)).unsqueeze(2) * col).sum(1)
    assert torch.allclose(bad, want, atol=1e-6)


def test_predict_shapes_and_asserts_both_configs():
    for cfg, r, s in ((M.ModelConfig.as_shipped(), 84, 64), (M.ModelConfig(hidden=64), 8, 16)):
        p = M.init_params(cfg, 0)
        b = r * s
        pts = torch.rand(b * 3) * 2 - 1
        t = torch.sort(torch.rand(r, s) * 2, 1).values.reshape(-1)
        dirs = torch.nn.functional.normalize(torch.randn(r, 3), dim=1) if cfg.cd else None
        out, sig = M.predict(cfg, p, pts, t, r, s, dirs)
        assert out.shape == (r, 4) and sig.shape == (r, s)
        try:
            M.predict(cfg, p, pts[:-3], t, r, s, dirs)  # model.rs:162 assert_eq!
            raise RuntimeError("should have asserted")
        except AssertionError:
            pass
    # as shipped, alpha channel = sum of weights (colour 4 is 1)
    cfg = M.ModelConfig.as_shipped()
    cfg = M.replace(cfg, bug_compat_T_view=False)
    p = M.init_params(cfg, 0)
    r, s = 4, 64
    pts = torch.rand(r * s * 3)
    t = torch.sort(torch.rand(r, s) * 2, 1).values
    out, sig = M.predict(cfg, p, pts, t.reshape(-1), r, s)
    assert torch.allclose(out[:, 3], 1 - torch.exp(-(sig * M.deltas_from_t(t)).sum(1)), atol=1e-5)


def test_mse_and_adam_first_step():
    x, y = torch.rand(5, 4), torch.rand(5, 4)
    assert torch.allclose(M.mse_loss(x, y), torch.nn.functional.mse_loss(x, y))
    cfg = M.ModelConfig(hidden=64)
    p = M.init_params(cfg, 0)
    tr = M.Trainer(cfg, p, lr=5e-4)
    r, s = 8, 16
    pts = torch.rand(r * s * 3) * 2 - 1
    t = torch.sort(torch.rand(r, s) * 2, 1).values.reshape(-1)
    dirs = torch.nn.functional.normalize(torch.randn(r, 3), dim=1)
    before = tr.params_flat().clone()
    out, _ = tr.predict(pts, t, r, s, dirs)
    loss = tr.step(out, torch.rand(r * 4))
    assert loss > 0
    g = tr.grads_flat()
    moved = tr.params_flat() - before
    nz = g.abs() > 1e-5
    # Adam step 1 moves each parameter by ~ lr * sign(g)  (App. A.5)
    assert torch.allclose(moved[nz], -5e-4 * torch.sign(g[nz]), rtol=2e-2, atol=1e-7)
    # adam_reference matches torch.optim.Adam over several steps
    torch.manual_seed(3)
    pr = torch.randn(100, requires_grad=True)
    opt = torch.optim.Adam([pr], lr=5e-4)
    p2, m, v = pr.detach().clone(), torch.zeros(100), torch.zeros(100)
    for step in range(1, 6):
        gstep = torch.randn(100)
        pr.grad = gstep.clone()
        opt.step()
        p2, m, v = M.adam_reference(p2, gstep, m, v, step)
        assert torch.allclose(pr.detach(), p2, rtol=1e-6, atol=1e-8)


def test_posenc_layout_and_twins():
    x = np.array([[0.1, -0.7, 1.3]], dtype=np.float32)
    e = ray_np.posenc(x, 10)
    assert e.shape == (1, 63)
    assert np.array_equal(e[0, :3], x[0])
    assert np.allclose(e[0, 3:6], np.sin(x[0])) and np.allclose(e[0, 6:9], np.cos(x[0]))
    assert np.allclose(e[0, 9:12], np.sin(2 * x[0]), atol=1e-6)
    assert np.allclose(e[0, 57:60], np.sin(512 * x[0].astype(np.float64)), atol=1e-4)
    et = M.posenc(torch.from_numpy(x), 10).numpy()
    assert np.allclose(e, et, atol=1e-6)
    assert ray_np.posenc(x, 0).shape == (1, 3)
    assert ray_np.posenc(x, 4).shape == (1, 27)


def test_bf16_emulation_is_close_to_fp32():
    cfg = M.ModelConfig(hidden=64)
    p = M.init_params(cfg, 0)
    x = M.posenc(torch.rand(256, 3) * 2 - 1, 10)
    d = M.posenc(torch.nn.functional.normalize(torch.randn(256, 3), dim=1), 4)
    s0, c0, _ = M.mlp_forward(cfg, p, x, d)
    s1, c1, _ = M.mlp_forward(M.replace(cfg, emulate_bf16=True), p, x, d)
    assert torch.allclose(c0, c1, atol=1e-2) and torch.allclose(s0, s1, atol=2e-2)
    assert not torch.equal(c0, c1)
