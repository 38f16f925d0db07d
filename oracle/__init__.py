"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the nerf-rs per-ray hot path (ray geometry, depth sampling,
MLP, compositing, MSE, Adam). Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this package, and only as the
checker / reported baseline -- never as a product code path. The product
(``nerf_rs_b200``) raises if its CUDA library is missing; it has no CPU fallback.

Parity status: the ray geometry is PINNED on the reference's own exact tests
(``src/ray_sampling.rs:443-449`` and ``:70-77``). The tensor side (MLP,
compositing, loss, Adam) is "parity unpinned" by the reference -- it holds no
test or fixture for any tensor value and the Rust/tch binary
cannot be built here -- so it is restated op for op on torch 2.11 CPU (the same
ATen that tch 0.11 binds) and cross-checked with closed forms.
"""
from . import ray_np, ray_c, model_torch  # noqa: F401
