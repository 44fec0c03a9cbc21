"""bench.py contract (CPU part): the reference arm runs without a GPU and prints ONE JSON line with the agreed keys;
the GPU arm refuses to run without a CUDA device instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-rays", "20000", "--cpu-seconds", "0.2")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ray_surface_interactions_per_s" and d["value"] > 0
    assert d["unit"] == "interactions/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["workload"].startswith("C2") and d["config"]["rows"] == 14
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--warmup", "1", "--rays", "1000")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
