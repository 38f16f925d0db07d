"""K-composite (forward, backward + fused MSE) and K-adam vs the torch oracle."""
import numpy as np
import pytest
import torch

import nerf_rs_b200 as nb
from oracle import model_torch as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    return nb.NeRF(nb.default_config(image_w=16, image_h=16, num_rays=8, num_samples=8, hidden=64, mlp_impl=1))


@pytest.mark.parametrize("r,s", [(84, 64), (1024, 64), (257, 192), (5, 1), (3, 33), (16, 256)])
def test_compositing_matches_literal_reference_graph(model, r, s):
    torch.manual_seed(r * 1000 + s)
    sig = torch.randn(r, s)                       # raw sigma, may be negative (model.rs:168-171)
    t = torch.sort(torch.rand(r, s) * 2, dim=1).values
    delta = M.deltas_from_t(t)
    col = torch.rand(r, s, 4)
    want = M.compositing_literal(sig, col, delta)  # the reference's S-op graph, correct [ray,sample] layout
    got = nb.compositing(model, sig.numpy(), col.numpy(), delta.numpy())
    assert np.allclose(got, want.numpy(), rtol=1e-5, atol=1e-5)
    # as shipped: colours (sigma, sigma, sigma, 1) (model.rs:192-204)
    col2 = torch.stack([sig, sig, sig, torch.ones_like(sig)], 0).permute(1, 2, 0)
    want2 = M.compositing_literal(sig, col2, delta)
    got2 = nb.compositing(model, sig.numpy(), None, delta.numpy())
    assert np.allclose(got2, want2.numpy(), rtol=1e-5, atol=1e-5)


def test_compositing_closed_forms(model):
    r, s = 33, 64
    t = torch.sort(torch.rand(r, s) * 2, dim=1).values
    delta = M.deltas_from_t(t).numpy()
    ones = np.ones((r, s, 4), np.float32)
    assert np.array_equal(nb.compositing(model, np.zeros((r, s), np.float32), ones, delta), np.zeros((r, 4), np.float32))
    k = 1.7
    got = nb.compositing(model, np.full((r, s), k, np.float32), ones, delta)
    want = 1 - np.exp(-k * (2.0 - t[:, 0].numpy()))
    assert np.allclose(got, want[:, None], atol=3e-6)
    with pytest.raises(nb.NerfError):
        nb.compositing(model, np.zeros((4, 300), np.float32), None, np.zeros((4, 300), np.float32))
