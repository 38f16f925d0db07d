"""Out-of-bounds writes and uninitialised reads, checked by the library itself (compute-sanitizer is closed on the development
pool -- profiles/r02_sanitizer.md): every device buffer gets guard bands and a NaN fill (NERF_B200_GUARD=1, csrc/guard.h), the
whole hot path runs at ragged sizes for every MLP kernel family, then the bands must be intact and every output finite."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_kernel_writes_outside_its_buffers_or_reads_uninitialised_memory():
    env = dict(os.environ, NERF_B200_GUARD="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "guard_worker.py")], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    print(p.stdout[-2000:])
    assert p.returncode == 0 and "GUARDS_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]
