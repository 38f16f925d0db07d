"""CPU checks of the host-built per-tile programs the fused tcgen05 MLP kernels execute.

1. Numeric: the tables (MMA ops + packed weight chunks + epilogue jobs + weight-gradient units),
   emulated in float32 numpy, reproduce the torch oracle's MLP forward (src/model.rs:97-131),
   its pre-activation gradients and its parameter gradients.
2. (Protocol simulations of the kernels that run these tables: tests/test_lane_plan.py for the SS-mode
   CTA-pair kernel, tests/test_ts_plan.py for the TS-mode kernel, tests/test_wgrad_protocol.py.)
No GPU needed: the tables come from host-only debug entry points of libnerf_b200.so.
"""
import numpy as np
import pytest
import torch

import nerf_rs_b200 as nb
from oracle import model_torch as M
from oracle import ray_np
from tests import tc_plan_util as U

CONFIGS = {
    "ns256": dict(hidden=256),
    "ns128": dict(hidden=128),
    "ns64": dict(hidden=64),
    "w100_as_shipped": dict(hidden=100, xyz_freqs=0, dir_freqs=-1, skip_layer=0, use_rgb_head=0),
    "ns256_noskip_nodir": dict(hidden=256, skip_layer=0, dir_freqs=-1),
    "ns200": dict(hidden=200, xyz_freqs=6, dir_freqs=2, skip_layer=3),
    "ns150": dict(hidden=150),                                   # 129..192: padded to 4 panels
}


def _mcfg(over):
    d = dict(hidden=256, xyz_freqs=10, dir_freqs=4, skip_layer=5, use_rgb_head=1)
    d.update(over)
    return M.ModelConfig(hidden=d["hidden"], xyz_freqs=d["xyz_freqs"], dir_freqs=d["dir_freqs"], skip_layer=d["skip_layer"],
                         use_rgb_head=bool(d["use_rgb_head"]))


WIDE = {"ns512": dict(hidden=512), "ns480_noskip": dict(hidden=480, skip_layer=0), "ns300": dict(hidden=300)}   # (257..448: padded to 8 panels)   # pair kernel only (8 hidden panels)


@pytest.mark.parametrize("name", list(CONFIGS) + list(WIDE))
def test_program_numerics_match_oracle(name):
    over = CONFIGS[name] if name in CONFIGS else WIDE[name]
    cfg = nb.default_config(**over)
    mcfg = _mcfg(over)
    fwd = U.get_plan(cfg, 0)
    infer = U.get_plan(cfg, 1)
    bwd = U.get_plan(cfg, 2)
    assert fwd["n_params"] == mcfg.num_params()
    assert fwd["wpack_bytes"] == infer["wpack_bytes"] and np.array_equal(fwd["ops"], infer["ops"])
    params_t = M.init_params(mcfg, 0)
    params = M.flatten_params(params_t).numpy().copy()
    rng = np.random.default_rng(0)
    pts = (rng.random((128, 3)).astype(np.float32) * 2 - 1)
    dirs = rng.standard_normal((128, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    x_enc = ray_np.posenc(pts, mcfg.xyz_freqs)
    d_enc = ray_np.posenc(dirs, mcfg.dir_freqs) if mcfg.cd else np.zeros((128, 0), np.float32)

    # ---- forward
    out = U.emulate_chain(fwd, params, x_enc, d_enc)
    pt = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in params_t]
    xt = torch.from_numpy(x_enc)
    dt = torch.from_numpy(d_enc) if mcfg.cd else None
    sigma, rgba, _ = M.mlp_forward(mcfg, pt, xt, dt)
    assert np.allclose(out["sigma"], sigma.detach().numpy(), rtol=1e-4, atol=1e-5)
    if mcfg.use_rgb_head:
        assert np.allclose(out["rgba"], rgba.detach().numpy(), rtol=1e-4, atol=1e-5)
    else:
        assert out["rgba"] is None

    # ---- backward: inject upstream gradients, compare parameter gradients with autograd
    d_sigma = rng.standard_normal(128).astype(np.float32)
    d_rgba = rng.standard_normal((128, 4)).astype(np.float32)
    loss = (sigma * torch.from_numpy(d_sigma)).sum()
    if mcfg.use_rgb_head:
        loss = loss + (rgba * torch.from_numpy(d_rgba)).sum()
    loss.backward()
    want = torch.cat([torch.cat([w.grad.reshape(-1) if w.grad is not None else torch.zeros(w.numel()),
                                 b.grad.reshape(-1) if b.grad is not None else torch.zeros(b.numel())]) for w, b in pt]).numpy()
    bo = U.emulate_chain(bwd, params, x_enc, d_enc, d_sigma=d_sigma, d_rgba=d_rgba,
                         rgba=out["rgba"] if out["rgba"] is not None else np.zeros((128, 4), np.float32), masks=out["masks"])
    got = U.emulate_wgrad(bwd, out["saved"], bo["saved"], fwd["n_params"])
    scale = np.abs(want).max()
    assert np.allclose(got, want, rtol=2e-3, atol=2e-4 * scale), float(np.abs(got - want).max() / scale)


def test_chunk_stream_is_contiguous_and_bounded():
    cfg = nb.default_config()
    for program in (0, 2):
        p = U.get_plan(cfg, program)
        off = 0
        for op, pc in zip(p["ops"], p["chunks"]):
            assert op["w_off"] == off == pc["dst_off"]
            assert op["n"] % 16 == 0 and 16 <= op["n"] <= 256 and op["n"] * 128 <= 32768
            assert 1 <= op["kcount"] <= 4
            off += int(op["n"]) * 128
        assert off == p["wpack_bytes"]
    # north-star forward: 1.06 MB of bf16 weights per tile pass
    assert 1.0e6 < U.get_plan(cfg, 0)["wpack_bytes"] < 1.2e6


def test_unsupported_geometry_is_rejected():
    for hidden in (576, 1024):     # more than 8 panels (widths up to 512 are padded to 1, 2, 4 or 8 panels)
        cfg = nb.default_config(hidden=hidden)
        with pytest.raises(nb.NerfError):
            U.get_plan(cfg, 0)


@pytest.mark.parametrize("hidden,n_tiles,n_ctas", [(256, 2048, 148), (256, 42, 148), (512, 2048, 148), (128, 1, 148), (256, 2048, 16), (100, 7, 132)])
def test_wgrad_partition_covers_every_tile_once_and_is_balanced(hidden, n_tiles, n_ctas):
    """Host logic of the weight-gradient work split (tc_wgrad_partition): every (unit, tile) is assigned to exactly one CTA
    segment, no CTA has more than three segments, and at bench size the fitted cost per CTA is balanced to a few percent."""
    import ctypes
    from nerf_rs_b200 import _lib
    kw = dict(hidden=hidden)
    if hidden == 100:
        kw.update(xyz_freqs=0, dir_freqs=-1, skip_layer=0, use_rgb_head=0)
    cfg = nb.default_config(**kw)
    lib = _lib.load()
    out = np.zeros((n_ctas, 10), np.int32)
    cost = np.zeros(128, np.int32)
    nu = ctypes.c_int32(128)
    rc = lib.nerf_debug_wgrad_partition(ctypes.byref(cfg), n_ctas, n_tiles, out.ctypes.data_as(ctypes.c_void_p),
                                         cost.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nu))
    assert rc == 0
    n_units = nu.value
    seen = np.zeros((n_units, n_tiles), np.int32)
    load = np.zeros(n_ctas)
    for c in range(n_ctas):
        assert 0 <= out[c, 0] <= 3
        for k in range(out[c, 0]):
            u, t0, t1 = out[c, 1 + 3 * k: 4 + 3 * k]
            assert 0 <= u < n_units and 0 <= t0 < t1 <= n_tiles
            seen[u, t0:t1] += 1
            load[c] += (t1 - t0) * 2 * (516 + 87 * cost[u]) + 20000
    assert (seen == 1).all()
    if n_tiles >= 2048 and n_ctas == 148:
        busy = load[load > 0]
        assert len(busy) == n_ctas and busy.max() <= 1.06 * busy.mean()
