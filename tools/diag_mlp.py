"""Layer-by-layer diagnosis of the fused tcgen05 MLP on a GPU box: compares the saved activation /
gradient panels of tile 0 with the torch oracle (bf16-emulating). Prints a table; never raises."""
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, ".")
import nerf_rs_b200 as nb  # noqa: E402
from nerf_rs_b200 import _lib  # noqa: E402
from oracle import model_torch as M  # noqa: E402
from oracle import ray_np  # noqa: E402
from tests import gpu_util as G  # noqa: E402


def main(hidden=256, r=8, s=32):
    cfg = nb.default_config(image_w=100, image_h=100, num_rays=r, num_samples=s, hidden=hidden)
    m = nb.NeRF(cfg)
    mcfg = M.replace(G.model_cfg(cfg), emulate_bf16=True)
    params_t = M.init_params(mcfg, 0)
    m.set_weights(M.flatten_params(params_t).numpy())
    pts, t, dirs, gold = G.make_points(r, s, 1)
    out, sig = m.predict(pts, t, dirs.reshape(-1), train=True)
    n = min(128, r * s)
    x = torch.from_numpy(ray_np.posenc(pts.reshape(-1, 3)[:n], cfg.xyz_freqs))
    d = torch.from_numpy(np.repeat(ray_np.posenc(dirs, cfg.dir_freqs), s, axis=0)[:n])
    rb = lambda z: z.to(torch.bfloat16).to(torch.float32)
    np_ = (hidden + 63) // 64
    acts = {}
    h = rb(x)
    acts["X"] = h
    for l in range(1, 8):
        w, b = params_t[l - 1]
        h = rb(torch.relu(torch.nn.functional.linear(h, rb(w), b)))
        acts[f"H{l}"] = h
        if mcfg.skip_layer and l == mcfg.skip_layer:
            h = torch.cat([rb(x), h], -1)
    w, b = params_t[7]
    df = torch.nn.functional.linear(h, rb(w), b)
    acts["feat"] = rb(df[:, 1:])
    slots = {"X": (0, 1), "D": (1, 1)}
    for l in range(1, 8):
        slots[f"H{l}"] = (2 + (l - 1) * np_, np_)
    slots["feat"] = (2 + 7 * np_, np_)
    acts["D"] = rb(d)
    print(f"--- forward panels, tile 0 (hidden={hidden})")
    for name in ["X", "H1", "H2", "H3", "H4", "H5", "H6", "H7", "feat", "D"]:
        s0, cnt = slots[name]
        got = np.concatenate([G.decode_panel(m.debug_read_panel(0, 0, s0 + p)) for p in range(cnt)], axis=1)[:n]
        want = acts[name].numpy()
        wpad = np.zeros_like(got)
        wpad[:, :want.shape[1]] = want
        err = np.abs(got - wpad)
        print(f"{name:5s} max|err|={err.max():.4e}  mean|err|={err.mean():.3e}  max|ref|={np.abs(wpad).max():.3e}  "
              f"bad_rows={int((err.max(1) > 0.05 * max(1e-6, np.abs(wpad).max())).sum())} bad_cols={int((err.max(0) > 0.05 * max(1e-6, np.abs(wpad).max())).sum())}")
    print("sigma[:4]", sig.reshape(-1)[:4], "oracle", df[:4, 0].numpy())


if __name__ == "__main__":
    for hidden in (64, 128, 256):
        try:
            main(hidden)
        except Exception:
            traceback.print_exc()
