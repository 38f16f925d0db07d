"""Shared helpers for the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np
import torch

import nerf_rs_b200 as nb
from oracle import model_torch as M
from oracle import ray_c, ray_np


def have_gpu():
    try:
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def model_cfg(cfg, **over):
    return M.ModelConfig(hidden=cfg.hidden, xyz_freqs=cfg.xyz_freqs, dir_freqs=cfg.dir_freqs, skip_layer=cfg.skip_layer,
                         use_rgb_head=bool(cfg.use_rgb_head), sigma_relu=bool(cfg.sigma_relu), **over)


def make_points(r, s, seed=0):
    """Seeded (points[B*3], t[B], dirs[R,3], gold[R*4]) via the oracle sampler (real ray geometry)."""
    rng = np.random.default_rng(seed)
    idx = np.stack([rng.integers(0, 100, r), rng.integers(0, 100, r)], 1).astype(np.int64)
    u = np.sort(rng.random((r, s)).astype(np.float32), axis=1)
    yaw, pitch = np.float32(0.7), np.float32(0.4)
    pts, t = ray_np.sample_rays(idx, s, yaw, pitch, u, 100, 100)
    dirs = ray_np.ray_dirs(idx, yaw, pitch, 100, 100)
    gold = rng.random(r * 4).astype(np.float32)
    return pts.reshape(-1).copy(), t.reshape(-1).copy(), dirs.astype(np.float32).copy(), gold


def oracle_predict(mcfg, params_t, pts, t, dirs, r, s):
    out, sig = M.predict(mcfg, params_t, torch.from_numpy(pts), torch.from_numpy(t), r, s,
                         torch.from_numpy(dirs) if mcfg.cd else None, literal=False)
    return out, sig


def decode_panel(u16):
    """[8192] uint16 saved panel image -> float32 [128, 64]. Saved layout (chain-kernel epilogue -> wgrad operand):
    [half tile (64 rows)][16-byte chunk (8)][row (64)][8 bf16]."""
    r = np.arange(128)[:, None]
    c = np.arange(64)[None, :]
    off = (r >> 6) * 4096 + (c >> 3) * 512 + (r & 63) * 8 + (c & 7)
    v = u16[off].astype(np.uint32) << 16
    return v.view(np.float32)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def bf16_round(a):
    """float32 -> nearest-even bfloat16 -> float32 (what cvt.rn.bf16x2.f32 does)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()
