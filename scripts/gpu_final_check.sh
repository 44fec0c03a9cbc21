#!/bin/bash
# What the driver runs at round end, on one box: GPU tests, smoke(), the default bench line and the reference arm.
python -m pytest tests -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_default.json")); r = d["roofline"]
print({k: d.get(k) for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "dtype", "scaling", "vs_baseline", "gpu_launches")})
print("roofline", {k: r.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "frac_reference_work", "traffic", "kernel_ms")})
print("e2e", d["e2e"])
print("cpu", {k: d["cpu_baseline"].get(k) for k in ("value", "unit", "cores", "kind", "sample")})
print("clocks", d["clocks"])
print("others", list((d.get("other_configs") or {}).keys()))
r2 = json.load(open("gpurun_out/final_reference.json"))
print("ref", {k: r2.get(k) for k in ("impl", "value", "unit", "ms_per_step")}, r2["cpu_baseline"].get("kind"))
PY
