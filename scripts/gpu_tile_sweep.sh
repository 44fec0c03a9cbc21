#!/bin/bash
# A/B of the forward kernel variants (run under gpurun): RTT_FWD_TILE = 0 (per-ray kernel), 1, 2, 4 rays per thread.
set -u
TAG="${1:-sweep}"; OUT=gpurun_out; mkdir -p $OUT
for wl in c2 c1 c4; do
  for t in 3; do
    RTT_FWD_TILE=$t timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-bwd > $OUT/sweep_${wl}_t${t}_$TAG.json 2> $OUT/sweep_${wl}_t${t}_$TAG.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/sweep_${wl}_t${t}_$TAG.json").read())
    print("$wl tile=$t ms=%.3f value=%.4g frac=%.3f" % (d["ms_per_step"], d["value"], d["roofline"]["frac"]))
except Exception as e:
    print("$wl tile=$t failed", e)
PY
  done
done
