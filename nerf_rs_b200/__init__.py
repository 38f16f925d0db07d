"""nerf_rs_b200 -- B200-native drop-in for the per-ray training/render hot path of cadddr/nerf-rs.

The product is ``libnerf_b200.so`` (hand-written sm_100a kernels behind the C ABI in
``include/nerf_b200.h``); this package is the thin host-side mirror of the reference's call
surface used by the tests and ``bench.py``. Importing it never touches the oracle and there is
no CPU fallback: without the built library or a B200 the constructors raise.
"""
from ._lib import NerfConfig, NerfError, load, LIB_PATH  # noqa: F401
from .api import (NeRF, Trainer, Prediction, compositing, get_multiview_batch, get_view_angles, default_config,  # noqa: F401
                  as_shipped_config, load_image_as_array, load_image_rgba8, get_image_paths)
