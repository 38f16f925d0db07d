"""Data parallelism on real GPUs (needs >= 2; skipped otherwise): tests/dp_worker.py under torchrun, one process per GPU.
Covers the round-1 advisor finding (ranks used to draw identical rays) and what tools/dp_check.py checked by hand."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_data_parallel_ranks_draw_disjoint_rays_and_match_the_global_batch(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dp_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [l for l in (p.stdout + p.stderr).splitlines() if "OMP_NUM_THREADS" not in l and not l.startswith("*****")]
    tail = "\n".join(lines[-80:])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "dp_worker.log"), "w").write(p.stdout + "\n--- stderr\n" + p.stderr)
    print(tail)
    assert p.returncode == 0 and "DP_CHECK_OK" in p.stdout, tail
