"""CPU: the parts of the bench.py contract that need no GPU - the reference arm's JSON line (it times the CPU oracle port on the
host cores) and the product arm failing loudly, with no CPU fallback, when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "extend" in d["metric"] and "shadow" in d["metric"] and "Mrays/s" in base["metric"]
    assert "configs[1]" in d["config"]["workload"] and "1M-triangle" in d["config"]["workload"]
    assert d["steps"] == 1 and d["value"] > 0.05 and abs(d["ms_per_step"] * 1e-3 * d["value"] * 1e6 / 2967028 - 1) < 0.01   # one full frame per step
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and "1920x1080" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout) and "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(line.startswith("{") for line in r.stdout.splitlines())          # no result line is printed
