#!/bin/bash
# Bench lines of every workload on one box (no CPU legs).  Usage: gpu_bench_all.sh <tag> ["<workloads>"]
set -u
TAG="${1:-all}"; WLS="${2:-c2 c1 c4 c4cam c5 c3}"
OUT=gpurun_out; mkdir -p $OUT
for wl in $WLS; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-config4 --no-other-configs > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err
  echo "$wl exit $? $(python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_${wl}_$TAG.json')); r=d['roofline']; fb=d.get('fwd_bwd') or {}
    adj=(fb.get('adjoint') or {})
    print('ms', round(d['ms_per_step'],3), 'kernel', r.get('kernel'), round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'ref_work', round(r.get('frac_reference_work') or 0,3),
          '| fwd_bwd', round(fb.get('ms_per_step') or 0,2), 'goal', round((fb.get('goal') or {}).get('ms_per_step') or 0,2), 'adj_ms', round(adj.get('kernel_ms') or 0,3), 'adj_frac', round(adj.get('frac') or 0,3))
except Exception as e:
    print('unreadable', e)
PY
)"
done
