"""Native callers of the C ABI (no Python, no torch in the process that drives the GPU):

* tests/native/nerf_driver.c -- plain C, the call sequence of src/main.rs:44-72 (set_images -> get_batch(host indices, jitter)
  -> predict_points -> step) for three iterations; on a GPU its dump (points, distances, gold, pixels, densities, losses, final
  weights) is compared with the CPU oracle. This is what a Rust `extern "C"` binding executes (ffi/rust/src/lib.rs).
* nerf_rs_b200/csrc/host/nerf_b200.hpp -- the C++ mirror of the reference's call surface: compiles, links, and without a GPU
  fails loudly with NERF_ERR_NO_DEVICE (there is no CPU fallback)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "nerf_rs_b200")
LINK = ["-L", LIBDIR, "-lnerf_b200", f"-Wl,-rpath,{LIBDIR}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]

SRC = r'''
#include "nerf_rs_b200/csrc/host/nerf_b200.hpp"
#include <cmath>
#include <cstdio>
int main() {
    auto angles = nerf::get_view_angles(6);
    if (angles.size() != 84) return 2;
    try { nerf::load_image_as_array("/nonexistent/image-0.png"); return 4; } catch (const nerf::Error &) {}   // image_loading.rs:6
    try {
        nerf::NeRF model;                      // NeRF::new(): needs a B200
        std::mt19937_64 rng(0);
        model.set_view_angles(angles);
        auto batch = nerf::get_multiview_batch(model, rng);
        nerf::Trainer trainer(model);
        auto log = model.log_metrics(false, false);          // logging.rs:13-107 on the device
        // the prediction as a device handle (main.rs:58 -> :72): step consumes it, the pixels come on demand
        auto &[indices, qp, dist, gold] = batch;
        std::vector<float> dirs((size_t)model.config().num_rays * 3, 0.f);
        for (size_t i = 2; i < dirs.size(); i += 3) dirs[i] = -1.f;
        auto pred = model.predict_device(qp, dist, &dirs);
        auto px_before = pred.pixels();
        float loss = trainer.step(pred, gold);
        auto px_after = pred.pixels();                          // (the compositing backward rewrote the same pixels)
        if (!(loss > 0.f) || px_before.size() != gold.size()) return 5;
        for (size_t i = 0; i < px_before.size(); ++i) if (std::fabs(px_before[i] - px_after[i]) > 1e-5f) return 6;
        auto eager = model.predict(qp, dist, &dirs);
        try { pred.pixels(); return 7; } catch (const nerf::Error &e) { if (e.status != NERF_ERR_STATE) return 8; }   // stale handle
        (void)indices; (void)log; (void)eager;
    } catch (const nerf::Error &e) {
        std::printf("status %d\n", e.status);
        return e.status == NERF_ERR_NO_DEVICE ? 0 : 3;
    }
    std::printf("ran on the device\n");
    return 0;
}
'''


def _build_driver(d):
    exe = os.path.join(d, "nerf_driver")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "native", "nerf_driver.c"), "-o", exe] + LINK)
    return exe


def test_cpp_mirror_and_c_driver_compile_link_and_fail_loudly_without_gpu():
    import torch
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        open(src, "w").write(SRC)
        subprocess.check_call(["g++", "-std=c++17", "-I", ROOT, src, "-o", exe] + LINK)
        drv = _build_driver(d)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        if torch.cuda.is_available():
            assert "ran on the device" in r.stdout
            return
        assert "status -6" in r.stdout                       # NERF_ERR_NO_DEVICE: no CPU fallback
        inp = os.path.join(d, "in.bin")
        np.array([16, 16, 8, 8, 64, 2, 1, 1], np.int32).tofile(inp)
        r = subprocess.run([drv, inp, os.path.join(d, "out.bin")], capture_output=True, text=True)
        assert r.returncode == 10 + 6 and "nerf_create" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_native_c_driver_matches_oracle_over_three_steps():
    import torch
    from oracle import model_torch as M
    from oracle import ray_c
    w = h = 100
    R, S, hidden, n_views, n_picks, n_steps = 256, 64, 256, 4, 4, 3
    mcfg = M.ModelConfig(hidden=hidden)
    params_t = M.init_params(mcfg, 0)
    weights = M.flatten_params(params_t).numpy()
    rng = np.random.default_rng(11)
    images = rng.random((n_views, w * h, 4)).astype(np.float32)
    angles = ray_c.get_view_angles(6)[:n_views].astype(np.float32)
    steps = []
    for _ in range(n_steps):
        idx = np.stack([rng.integers(0, h, R), rng.integers(0, w, R)], 1).astype(np.int64)
        vi = rng.integers(0, n_views, n_picks).astype(np.int64)
        jit = np.sort(rng.random((R, S)).astype(np.float32), axis=1)
        steps.append((idx, vi, jit))
    with tempfile.TemporaryDirectory() as d:
        exe = _build_driver(d)
        inp, outp = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(inp, "wb") as f:
            np.array([w, h, R, S, hidden, n_views, n_picks, n_steps], np.int32).tofile(f)
            weights.astype(np.float32).tofile(f); images.tofile(f); angles.tofile(f)
            for idx, vi, jit in steps:
                idx.tofile(f); vi.tofile(f); jit.tofile(f)
        r = subprocess.run([exe, inp, outp], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = open(outp, "rb").read()
    n_params = int(np.frombuffer(raw[:8], np.int64)[0])
    assert n_params == weights.size
    off = 8
    B = R * S

    def take(n):
        nonlocal off
        a = np.frombuffer(raw, np.float32, n, off)
        off += 4 * n
        return a

    tr = M.Trainer(M.replace(mcfg, emulate_bf16=True, emulate_bf16_grads=True), params_t, lr=5e-4)
    for it, (idx, vi, jit) in enumerate(steps):
        pts, t, gold, pix, sig, loss = take(3 * B), take(B), take(4 * R), take(4 * R), take(B), float(take(1)[0])
        _, pts_o, t_o, gold_o = ray_c.get_multiview_batch(images, angles, idx, vi, S, jit, w, h)
        assert pts.tobytes() == pts_o.tobytes() and t.tobytes() == t_o.tobytes() and gold.tobytes() == gold_o.tobytes()   # bit-exact
        dirs = np.concatenate([ray_c.ray_dirs(idx[i * (R // n_picks):(i + 1) * (R // n_picks)], float(angles[vi[i]][0]),
                                              float(angles[vi[i]][1]), w, h) for i in range(n_picks)])
        o, s_o = tr.predict(torch.from_numpy(pts_o.reshape(-1)), torch.from_numpy(t_o.reshape(-1)), R, S, torch.from_numpy(dirs), literal=False)
        want_loss = tr.step(o, torch.from_numpy(gold_o.reshape(-1)))
        o, s_o = o.detach().numpy().reshape(-1), s_o.detach().numpy().reshape(-1)
        e_pix, e_sig = np.abs(pix - o).max() / np.abs(o).max(), np.abs(sig - s_o).max() / np.abs(s_o).max()
        print(f"native step {it}: pixels {e_pix:.2e}, sigma {e_sig:.2e} relative to the oracle; loss {loss:.6f} vs {want_loss:.6f}")
        assert e_pix < 1e-2 and e_sig < 1e-2                       # north-star tolerance (weights drift apart by bf16 noise over the steps)
        assert abs(loss - want_loss) < 1e-2 * want_loss
    w_final = take(n_params)
    assert off == len(raw)
    w_oracle = tr.params_flat().numpy()
    # three Adam steps at lr 5e-4 moved the weights, and both paths moved them the same way. (Adam's early updates are ~ +-lr whatever
    # the gradient's size, so parameters whose gradient is bf16 noise can step in opposite directions: compare the update
    # vectors, not single elements.)
    dw, dw_o = w_final - weights, w_oracle - weights
    rel = float(np.linalg.norm(dw - dw_o) / np.linalg.norm(dw_o))
    print(f"native driver: |dw| = {np.linalg.norm(dw_o):.3e}, relative difference of the 3-step update vs the oracle {rel:.2e}")
    assert np.abs(dw_o).max() > 1e-4 and rel < 0.1
