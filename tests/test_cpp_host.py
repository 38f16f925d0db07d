"""The C++ host mirror of the reference call surface compiles against the C ABI (g++, no GPU run)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "nerf_rs_b200/csrc/host/nerf_b200.hpp"
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
        (void)batch; (void)trainer; (void)log;
    } catch (const nerf::Error &e) {
        std::printf("status %d\n", e.status);
        return e.status == NERF_ERR_NO_DEVICE ? 0 : 3;
    }
    return 0;
}
'''


def test_cpp_mirror_compiles_links_and_fails_loudly_without_gpu():
    import torch
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        exe = os.path.join(d, "t")
        open(src, "w").write(SRC)
        libdir = os.path.join(ROOT, "nerf_rs_b200")
        subprocess.check_call(["g++", "-std=c++17", "-I", ROOT, src, "-o", exe, "-L", libdir, "-lnerf_b200", f"-Wl,-rpath,{libdir}",
                               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
        if torch.cuda.is_available():
            return
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "status -6" in r.stdout
