"""bench.py contract on CPU: the reference arm prints ONE JSON line with the keys the driver reads, ranks > 0 stay
silent, and the GPU arm refuses loudly when there is no CUDA device (there is no CPU fallback to measure)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], {"PPM_BENCH_CPU_ROWS": "4"})
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pixels/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("configs[4]")         # the north-star job is the default workload
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "4 of 1080" in cb["sample"]
    assert cb["extrapolated"] is True and "EXTRAPOLATED" in cb["sample"] and "stratified" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run(["--steps", "1", "--warmup", "0", "--no-cpu"])
    assert p.returncode != 0
    assert p.stdout.strip() == ""                       # no bench line is ever printed from a CPU
    assert "CUDA" in p.stderr or "NODEVICE" in p.stderr
