"""Build libnerf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnerf_b200.so")
SOURCES = ["context.cu", "sampling.cu", "composite.cu", "adam.cu", "mlp_simt.cu", "mlp_tc.cu", "mlp_tc2.cu", "mlp_tc3.cu", "mlp_tc_plan.cpp", "comm.cpp", "host_io.cpp", "metrics.cu", "guard.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-Xptxas", "-v",
]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    # (files only: a directory's mtime changes whenever the tree is copied, which would rebuild on every fresh box)
    return any(os.path.isfile(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src + ".o")
        objs.append(obj)
        cmd = ["nvcc"] + NVCC_FLAGS + os.environ.get("NERF_B200_NVCC_EXTRA", "").split() + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src}\n{out}\n")
        elif verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
        with open(os.path.join(HERE, "build", src + ".log"), "w") as f:
            f.write(out)
    if failed:
        raise RuntimeError("nvcc build failed")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs + ["-lcudart", "-ldl", "-lz"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
