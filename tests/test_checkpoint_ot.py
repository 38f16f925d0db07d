"""tch `.ot` VarStore checkpoints (src/model.rs:211-217). The container format is pinned on a file written by
libtorch's own OutputArchive (tests/golden/tch_varstore.ot, generator beside it); the variable-name scheme is restated from
tch (not vendored): see nerf_rs_b200/checkpoint.py."""
import os

import numpy as np
import pytest

import nerf_rs_b200 as nb
from nerf_rs_b200 import checkpoint as C

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tch_varstore.ot")
DIMS = [(3, 6)] + [(6, 6)] * 6 + [(6, 7), (6, 3), (3, 4)]     # the fixture's ten layers (in, out)


def _fixture_expected():
    layers, nxt = [], 0.0
    for din, dout in DIMS:
        b = np.arange(dout, dtype=np.float32) * np.float32(0.5) + np.float32(nxt)
        nxt += 100.0
        w = (np.arange(dout * din, dtype=np.float32) * np.float32(0.25) + np.float32(nxt)).reshape(dout, din)
        nxt += 100.0
        layers.append((w, b))
    return layers


def test_reads_a_file_written_by_libtorch_outputarchive():
    t = C.read_ot(GOLD)
    assert list(t)[:4] == ["bias", "weight", "bias__2", "weight__3"] and len(t) == 20
    assert [n for pair in C.var_names(10) for n in pair] == list(t)
    got = C.layers_from_vars(t, DIMS)
    for (w, b), (we, be) in zip(got, _fixture_expected()):
        assert np.array_equal(w, we) and np.array_equal(b, be)
    flat = C.flat_from_layers(got)
    assert flat.size == sum(i * o + o for i, o in DIMS) and flat[0] == 100.0      # fc1 weight first, then its bias


def test_round_trip_with_adam_state_and_foreign_pair_order(tmp_path):
    rng = np.random.default_rng(0)
    flat = rng.standard_normal(sum(i * o + o for i, o in DIMS)).astype(np.float32)
    p = str(tmp_path / "ck.ot")
    C.save_varstore(p, flat, DIMS, adam=(flat * 2, flat * flat, 41))
    f2, adam = C.load_varstore(p, DIMS)
    assert np.array_equal(f2, flat) and adam[2] == 41 and np.array_equal(adam[0], flat * 2) and np.array_equal(adam[1], flat * flat)
    import torch
    names = [n for n, _ in torch.jit.load(p).named_parameters()]
    assert names[:20] == [n for pair in C.var_names(10) for n in pair]            # what VarStore::load looks up
    # a tch whose nn::linear registers the weight first: weight, bias, weight__2, bias__3, ...
    named, k = [], 0
    for w, b in C.layers_from_flat(flat, DIMS):
        named += [("weight" if k == 0 else f"weight__{k}", w), ("bias" if k == 0 else f"bias__{k + 1}", b)]
        k += 2
    q = str(tmp_path / "wf.ot")
    C.write_ot(q, named)
    f3, adam3 = C.load_varstore(q, DIMS)
    assert np.array_equal(f3, flat) and adam3 is None


def test_shape_mismatch_is_an_error(tmp_path):
    with pytest.raises(ValueError):
        C.layers_from_vars(C.read_ot(GOLD), [(3, 8)] + DIMS[1:])
    with pytest.raises(ValueError):
        C.layers_from_vars(C.read_ot(GOLD), DIMS[:9])


def test_layer_dims_follow_the_config():
    assert C.layer_dims(nb.as_shipped_config()) == [(3, 100)] + [(100, 100)] * 6 + [(100, 101), (100, 50), (50, 4)]   # model.rs:48-55, 89-90
    d = C.layer_dims(nb.default_config())
    assert d[0] == (63, 256) and d[5] == (319, 256) and d[7] == (256, 257) and d[8] == (283, 128) and d[9] == (128, 4)
    assert sum(i * o + o for i, o in d) == 530181


@pytest.mark.gpu
def test_model_save_load_ot(tmp_path):
    from oracle import model_torch as M
    from tests import gpu_util as G
    cfg = nb.default_config(image_w=64, image_h=64, num_rays=64, num_samples=16, hidden=64)
    a = nb.NeRF(cfg)
    a.set_weights(M.flatten_params(M.init_params(G.model_cfg(cfg), 0)).numpy())
    pts, t, dirs, gold = G.make_points(64, 16, 5)
    out, _ = a.predict(pts, t, dirs.reshape(-1), train=True)
    nb.Trainer(a).step(out, gold)                       # one Adam step so the optimiser state is not trivial
    p = str(tmp_path / "checkpoint-0-1.ot")
    a.save(p)
    b = nb.NeRF(cfg)
    b.load(p)
    assert np.array_equal(a.get_weights(), b.get_weights())
    ma, va, sa = a.get_adam_state()
    mb, vb, sb = b.get_adam_state()
    assert sa == sb == 1 and np.array_equal(ma, mb) and np.array_equal(va, vb)
    oa, _ = a.predict(pts, t, dirs.reshape(-1), train=False)
    ob, _ = b.predict(pts, t, dirs.reshape(-1), train=False)
    assert np.array_equal(oa, ob)
    c = nb.NeRF(nb.default_config(image_w=64, image_h=64, num_rays=64, num_samples=16, hidden=128))
    with pytest.raises(nb.NerfError):
        c.load(p)                                       # wrong widths: refused, not silently reshaped
