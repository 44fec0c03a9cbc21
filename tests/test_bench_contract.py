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


def _check_reference_line(p, kind):
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ray_surface_interactions_per_s" and d["value"] > 0
    assert d["unit"] == "interactions/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["workload"].startswith("C2") and d["config"]["rows"] == 14
    cb = d["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    return d


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """The oracle port as the timed CPU path (what runs when the reference copy is missing)."""
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-rays", "20000", "--cpu-seconds", "0.2",
             "--port-only")
    _check_reference_line(p, "port")


def test_reference_arm_times_the_unmodified_reference_when_it_is_available():
    """kind == "reference": SequentialScene.simulate of the reference's own classes (from /root/reference here, from
    the copy __graft_entry__.build() ships under baseline/_ref on the GPU box), the port as a second figure, and one
    unmodified benchmarks/sim_benchmark.main()."""
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    if not ref_loader.reference_available():
        import pytest
        pytest.skip("no reference checkout or shipped copy in this environment")
    env_small = dict(os.environ, BENCH_REPEATS="2", BENCH_WARMUP="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-rays", "20000", "--cpu-seconds", "0.2"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=env_small)
    d = _check_reference_line(p, "reference")
    cb = d["cpu_baseline"]
    assert cb["port"]["value"] > 0
    sb = cb["sim_benchmark"]
    assert "error" not in sb, sb
    assert set(sb["ms_by_rays"]) == {"4096", "16384", "64000", "128000"} and sb["rays_per_s"] > 0


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--warmup", "1", "--rays", "1000")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
