"""bench.py's contract on the CPU: the reference arm (--impl reference) runs without a GPU and prints ONE JSON line with the keys
the driver reads; under a multi-rank launch only rank 0 prints; the B200 arm fails loudly without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "training_rays_per_sec" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert "BASELINE configs[1]" in d["config"]["workload"] and "clamped" in d["config"]["note"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["ms_sampler"] > 0 and cb["ms_model_loss_backward_adam"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert p.returncode == 0 and not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = _run(["--steps", "1", "--warmup", "1", "--no-cpu", "--no-extra"], timeout=300)
    assert p.returncode != 0 and not [l for l in p.stdout.splitlines() if l.startswith("{")]
