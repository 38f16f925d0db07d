#!/usr/bin/env python
"""HBM-bound stage kernels alone at render-scale sizes (working set > the 126 MB L2): achieved GB/s of the algorithmic
bytes against MEASURED_PEAKS.json's copy bandwidth. Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_rs_b200 as nb  # noqa: E402

BYTES = {"sample": lambda r, s: 16 * r * s + 8 * r + 28 * r,            # points 12 + t 4 per sample; pixel, ray record, dir per ray
         "composite_fwd": lambda r, s: 24 * r * s + 16 * r,
         "composite_bwd": lambda r, s: 44 * r * s + 52 * r,               # 24 read + 20 written per sample; gold, pixels, loss per ray
         "adam": lambda r, s: 28 * r * s}


def run(model, sizes=((262144, 64), (131072, 192)), iters=10, stages=("sample", "composite_fwd", "composite_bwd", "adam"), tag=""):
    peak = 6550.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    out = {}
    for name in stages:
        for r, s in sizes:
            if name == "adam" and s != sizes[0][1]:
                continue
            ms = model.bench_stage(name, r, s, iters)
            b = BYTES[name](r, s)
            out[f"{name}{tag}_{r}x{s}"] = {"ms": round(ms, 4), "bytes": b, "achieved": b / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": b / (ms * 1e-3) / 1e9 / peak}
    return out


if __name__ == "__main__":
    m = nb.NeRF(nb.default_config(num_rays=256, num_samples=64, hidden=64))
    out = run(m)
    m2 = nb.NeRF(nb.default_config(num_rays=256, num_samples=64, hidden=64, depth_mode=1))   # stratified depths: no sort
    out.update(run(m2, stages=("sample",), tag="_stratified"))
    print(json.dumps(out, indent=1))
