"""CPU checks of the TS-mode chain programs (mlp_tc3.cu; hidden <= 256) -- the tables the kernel receives as parameters.

1. Numeric: emulated in float32 numpy (MMA ops on activation panels / slot-E panels, packed weight chunks, half-GEMM steps
   with the first half held back until the layer's second half is done, ReLU masks derived from the stored activations)
   they reproduce the torch oracle's MLP forward (src/model.rs:97-131) and, through the weight-gradient units, its
   parameter gradients.
2. Protocol: a random-interleaving simulation of the weight producer, the two lanes' MMA issuers, the in-order tensor pipe,
   the epilogue and the saver warps finds no deadlock and no hazard (stale / premature activation or slot-E read, ring stage
   or accumulator overwritten in use, slot-E panel of the next tile written before its last reader was issued).
3. Structure: what the kernel relies on (<= 5 ops per step, N in {32, 64, 128}, contiguous chunk stream, one slot-E write
   per panel and tile, placed after the panel's last reader).
No GPU needed: the tables come from a host-only debug entry point of libnerf_b200.so.
"""
import numpy as np
import pytest
import torch

import nerf_rs_b200 as nb
from oracle import model_torch as M
from oracle import ray_np
from tests import tc_plan_util as U
from tests import ts_plan_util as T
from tests.test_tc_plan import CONFIGS, _mcfg


@pytest.mark.parametrize("name", list(CONFIGS))
def test_ts_program_numerics_match_oracle(name):
    over = CONFIGS[name]
    cfg = nb.default_config(**over)
    mcfg = _mcfg(over)
    base = U.get_plan(cfg, 0)                      # biases, weight-gradient units, slot counts (shared with the SS programs)
    fwd, infer, bwd = (T.get_ts_plan(cfg, p) for p in range(3))
    assert np.array_equal(fwd["ops"], infer["ops"]) and fwd["wpack_bytes"] == infer["wpack_bytes"]
    params_t = M.init_params(mcfg, 0)
    params = M.flatten_params(params_t).numpy().copy()
    bias = U.padded_bias(base, params)
    rng = np.random.default_rng(0)
    pts = (rng.random((128, 3)).astype(np.float32) * 2 - 1)
    dirs = rng.standard_normal((128, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    x_enc = ray_np.posenc(pts, mcfg.xyz_freqs)
    d_enc = ray_np.posenc(dirs, mcfg.dir_freqs) if mcfg.cd else np.zeros((128, 0), np.float32)

    out = T.emulate_ts_chain(fwd, bias, params, x_enc, d_enc)
    out_i = T.emulate_ts_chain(infer, bias, params, x_enc, d_enc)
    pt = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in params_t]
    sigma, rgba, _ = M.mlp_forward(mcfg, pt, torch.from_numpy(x_enc), torch.from_numpy(d_enc) if mcfg.cd else None)
    assert np.allclose(out["sigma"], sigma.detach().numpy(), rtol=1e-4, atol=1e-5)
    assert np.array_equal(out["sigma"], out_i["sigma"]) and not out_i["saved"]
    if mcfg.use_rgb_head:
        assert np.allclose(out["rgba"], rgba.detach().numpy(), rtol=1e-4, atol=1e-5)
    else:
        assert out["rgba"] is None

    d_sigma = rng.standard_normal(128).astype(np.float32)
    d_rgba = rng.standard_normal((128, 4)).astype(np.float32)
    loss = (sigma * torch.from_numpy(d_sigma)).sum()
    if mcfg.use_rgb_head:
        loss = loss + (rgba * torch.from_numpy(d_rgba)).sum()
    loss.backward()
    want = torch.cat([torch.cat([w.grad.reshape(-1) if w.grad is not None else torch.zeros(w.numel()),
                                 b.grad.reshape(-1) if b.grad is not None else torch.zeros(b.numel())]) for w, b in pt]).numpy()
    bo = T.emulate_ts_chain(bwd, bias, params, x_enc, d_enc, d_sigma=d_sigma, d_rgba=d_rgba,
                            rgba=out["rgba"] if out["rgba"] is not None else np.zeros((128, 4), np.float32), masks=out["masks"])
    # the same saved-panel slots feed the weight-gradient units as in the SS programs (a missing slot is a KeyError there)
    got = U.emulate_wgrad(U.get_plan(cfg, 2), out["saved"], bo["saved"], base["n_params"])
    scale = np.abs(want).max()
    assert np.allclose(got, want, rtol=2e-3, atol=2e-4 * scale), float(np.abs(got - want).max() / scale)


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("program", [0, 1, 2])
def test_ts_protocol_simulation(name, program):
    cfg = nb.default_config(**CONFIGS[name])
    ts = T.get_ts_plan(cfg, program)
    rng = np.random.default_rng(program * 100 + len(name))
    for n_pairs in (1, 2, 3, 5):       # one lane only; both lanes; lane 1 runs out first
        T.simulate_ts_protocol(ts, n_pairs, rng, train=program != 1)


def test_ts_protocol_simulation_catches_broken_programs():
    """The simulator is only worth something if it fails on wrong schedules."""
    cfg = nb.default_config()
    rng = np.random.default_rng(5)
    ts = T.get_ts_plan(cfg, 0)
    # (1) the next tile's encoded positions written one GEMM too early: fc6 (the skip layer) has not read the current ones
    bad = dict(ts, steps=ts["steps"].copy())
    i = int(np.nonzero(bad["steps"]["pre_enc"] == T.PRE_A)[0][0])
    bad["steps"]["pre_enc"][i] = T.PRE_NONE
    bad["steps"]["pre_enc"][i - 4] = T.PRE_A
    with pytest.raises(AssertionError):
        for _ in range(20):
            T.simulate_ts_protocol(bad, 4, rng, train=True)
    # (2..4) protocol bugs switched on inside the simulator
    for bug in ("early_final_signal", "no_act_saved_wait", "empty_count_1"):
        with pytest.raises(AssertionError):
            for _ in range(30):
                T.simulate_ts_protocol(ts, 4, rng, train=True, broken=bug)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_ts_program_structure(name):
    cfg = nb.default_config(**CONFIGS[name])
    for program in range(3):
        ts = T.get_ts_plan(cfg, program)
        ops, steps, chunks = ts["ops"], ts["steps"], ts["chunks"]
        assert len(ops) <= 80 and len(steps) <= 24 and len(ops) == len(chunks)
        off = 0
        for op, pc in zip(ops, chunks):       # the packed stream is contiguous, chunk = [n rows][64 K] bf16
            assert op["w_off"] == off == pc["dst_off"] and op["n"] in (32, 64, 128)
            off += int(op["n"]) * 128
        assert off == ts["wpack_bytes"]
        assert steps[0]["op_begin"] == steps[0]["op_end"] == 0
        for st in steps[1:]:
            assert 1 <= st["op_end"] - st["op_begin"] <= 5
        # each slot-E panel of the next tile is written exactly once, after the panel's last reader
        for what, src in ((T.PRE_A, T.A_SMEM), (T.PRE_B, T.A_SMEM_B)):
            readers = [i for i, st in enumerate(steps) if i and any(ops[o]["a_src"] == src for o in range(st["op_begin"], st["op_end"]))]
            writers = [i for i, st in enumerate(steps) if st["pre_enc"] == what]
            if readers:
                assert len(writers) == 1 and writers[0] > max(readers), (name, program, what, readers, writers)
            else:
                assert not writers
