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
    tail = "\n".join((p.stdout + p.stderr).splitlines()[-30:])
    print(tail)
    assert p.returncode == 0 and "DP_CHECK_OK" in p.stdout, tail
