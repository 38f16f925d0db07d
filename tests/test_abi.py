"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls without a GPU), and the product refuses to run without one."""
import ctypes
import os
import re

import pytest

import nerf_rs_b200 as nb
from nerf_rs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    lib = nb.load()
    declared = _declared()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.nerf_abi_version() == 2


def test_config_struct_matches_header():
    cfg = nb.default_config()
    assert cfg.struct_size == ctypes.sizeof(nb.NerfConfig) == 19 * 4 and cfg.deterministic_grads == 0
    assert (cfg.image_w, cfg.num_rays, cfg.num_samples, cfg.hidden, cfg.xyz_freqs, cfg.dir_freqs, cfg.skip_layer) == (800, 4096, 64, 256, 10, 4, 5)
    assert abs(cfg.learning_rate - 5e-4) < 1e-9          # cli.rs:64-65
    s = nb.as_shipped_config()
    assert (s.image_w, s.num_rays, s.num_samples, s.hidden, s.xyz_freqs, s.dir_freqs, s.use_rgb_head) == (128, 84, 64, 100, 0, -1, 0)


def test_error_strings_and_null_handling():
    lib = nb.load()
    assert lib.nerf_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.nerf_strerror(_lib.ERR_NO_DEVICE)
    assert lib.nerf_default_config(None) == _lib.ERR_INVALID_ARG
    assert lib.nerf_destroy(None) == 0
    assert lib.nerf_num_params(None) == 0


def test_no_gpu_means_no_product():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nb.NerfError) as e:
        nb.NeRF(nb.default_config())
    assert e.value.status == _lib.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    import subprocess
    import sys
    code = "import sys; import nerf_rs_b200; nerf_rs_b200.load(); print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    assert out == "False"
    for fn in os.listdir(os.path.join(ROOT, "nerf_rs_b200")):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "nerf_rs_b200", fn)).read().replace("no CPU fallback", "").lower().replace("the oracle", "") or fn == "__init__.py"


def test_view_angle_grid_matches_oracle():
    from oracle import ray_c
    for n in (1, 2, 6):
        assert nb.get_view_angles(n).tobytes() == ray_c.get_view_angles(n).tobytes()


def test_bench_touches_the_oracle_only_in_its_cpu_leg():
    """bench.py may execute oracle/ only as the CPU baseline / reference arm: every `oracle` import sits inside cpu_reference()."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    offenders = []
    for fn in ast.walk(tree):
        if isinstance(fn, (ast.FunctionDef, ast.Module)):
            for node in (fn.body if isinstance(fn, ast.Module) else ast.walk(fn)):
                names = []
                if isinstance(node, ast.ImportFrom) and node.module:
                    names = [node.module]
                elif isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                if any(n == "oracle" or n.startswith("oracle.") or n == "tests" or n.startswith("tests.") for n in names):
                    where = fn.name if isinstance(fn, ast.FunctionDef) else "<module>"
                    if where != "cpu_reference":
                        offenders.append((where, names))
    assert not offenders, offenders


def test_bench_synthetic_weights_are_the_survey_init():
    """bench.py builds its weights without oracle code; they must still be SURVEY 8d's `torch.manual_seed(0)` nn.Linear init."""
    import sys
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    from oracle import model_torch as M
    for hidden in (256, 512):
        cfg = nb.default_config(hidden=hidden)
        want = M.flatten_params(M.init_params(M.ModelConfig(hidden=hidden), 0)).numpy()
        assert np.array_equal(bench.synthetic_weights(cfg), want)


def test_built_kernels_contain_the_blackwell_instructions():
    """SASS of the in-tree objects (cuobjdump, no GPU needed): the MLP kernels really issue tcgen05 MMAs (UTCHMMA), commits
    (UTCBAR), tensor-memory loads (LDTM) and -- the TS-mode chain -- tensor-memory STORES (STTM: the activations are the next
    layer's A operand in TMEM), bulk async copies (UBLKCP) and packed fp32x2 adds (FADD2); the production chain kernels
    compile without spills in inference and with at most a handful of spilled words when training
    (profiles/r01_trace_notes.md: register hygiene was worth 15 %)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    build = os.path.join(ROOT, "nerf_rs_b200", "build")
    want = {"mlp_tc3.cu.o": ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "FADD2", "UTCATOMSWS"],
            "mlp_tc2.cu.o": ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "FADD2", "UTCATOMSWS"],
            "mlp_tc.cu.o": ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "RED"]}
    for obj, mnemonics in want.items():
        path = os.path.join(build, obj)
        if not os.path.exists(path):
            pytest.skip(f"{obj} not built")
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        for m in mnemonics:
            assert m in sass, f"{m} missing from {obj}"

    def spills(log, names):
        lines = open(os.path.join(build, log)).read().splitlines()
        out = {}
        for i, ln in enumerate(lines):
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m and any(k in m.group(1) for k in names):
                sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", " ".join(lines[i:i + 4]))
                assert sp, lines[i:i + 4]
                out[m.group(1)] = tuple(int(x) for x in sp.groups())
        return out

    # SS-mode pair kernel (hidden 449..512 and A/B runs): inference forward, training forward, dgrad
    ss = spills("mlp_tc2.cu.log", ("k_chain2ILb0ELb0ELb0ELb0ELb0EE", "k_chain2ILb0ELb1ELb0ELb0ELb0EE", "k_chain2ILb1ELb1ELb0ELb0ELb0EE"))
    assert len(ss) == 3, list(ss)
    for name, (stack, st, ld) in ss.items():
        assert st == 0 and ld == 0, (name, stack, st, ld)
    # TS-mode kernel: <false,false,false> inference, <false,true,false> training forward, <true,true,false> dgrad
    ts = spills("mlp_tc3.cu.log", ("k_chain3ILb0ELb0ELb0EE", "k_chain3ILb0ELb1ELb0EE", "k_chain3ILb1ELb1ELb0EE"))
    assert len(ts) == 3, list(ts)
    for name, (stack, st, ld) in ts.items():
        assert stack <= 16 and st <= 32 and ld <= 64, (name, stack, st, ld)
        if "ILb0ELb0ELb0EE" in name:
            assert st == 0 and ld == 0, (name, stack, st, ld)
