"""Data-parallel check under torchrun (one process per GPU): the logic lives in tests/dp_worker.py (tests/test_gpu_dp.py runs
it whenever two GPUs are visible); this wrapper keeps `torchrun tools/dp_check.py` working."""
import os
import runpy

runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "dp_worker.py"), run_name="__main__")
