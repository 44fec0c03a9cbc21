#!/bin/bash
# A/B of the lean adjoint path (default) against the general adjoint for every ray (RTT_BWD_MINB=8 -> tune bit 8), one box.
# Usage: gpu_lean_ab.sh <tag> "<workloads>"
set -u
TAG="$1"; WLS="$2"
OUT=gpurun_out; mkdir -p $OUT
for wl in $WLS; do
  for v in lean nolean; do
    if [ "$v" = lean ]; then unset RTT_BWD_MINB; else export RTT_BWD_MINB=8; fi
    timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-config4 --no-other-configs > $OUT/lean_${wl}_${v}_$TAG.json 2> $OUT/lean_${wl}_${v}_$TAG.err
    echo "$wl $v exit $? $(python - <<PY
import json
try:
    d=json.load(open('$OUT/lean_${wl}_${v}_$TAG.json')); r=d['roofline']; fb=d.get('fwd_bwd') or {}
    adj=(fb.get('adjoint') or {}); goal=(fb.get('goal') or {})
    print('ms', round(d['ms_per_step'],3), 'kernel_ms', round(r['kernel_ms'],3), 'frac', round(r.get('frac') or 0,3), '| fwd_bwd', round(fb.get('ms_per_step') or 0,2), 'goal', round(goal.get('ms_per_step') or 0,2), 'adj_ms', round(adj.get('kernel_ms') or 0,3), 'adj_frac', round(adj.get('frac') or 0,3))
except Exception as e:
    print('unreadable', e)
PY
)"
  done
done
unset RTT_BWD_MINB
