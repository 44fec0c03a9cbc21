#!/bin/bash
# A/B of adjoint builds (RTT_BWD_MINB -> mode tune bits of rtt_trace_seq_bwd) on one box.
# Usage: gpu_bwd_tune_ab.sh <tag> "<workloads>" <tune> [<tune> ...]   (default = unset; 8 = no lean path; 16 = lean in 256-thread blocks)
TAG="$1"; WLS="$2"; shift 2
for wl in $WLS; do
  for t in "$@"; do
    if [ "$t" = default ]; then unset RTT_BWD_MINB; else export RTT_BWD_MINB=$t; fi
    timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-config4 --no-other-configs > gpurun_out/bt_${wl}_${t}_$TAG.json 2> gpurun_out/bt_${wl}_${t}_$TAG.err
    echo "$wl bwd tune $t exit $? $(python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bt_${wl}_${t}_$TAG.json')); r=d['roofline']; fb=d.get('fwd_bwd') or {}
    adj=(fb.get('adjoint') or {}); goal=(fb.get('goal') or {})
    print('ms', round(d['ms_per_step'],3), 'kernel_ms', round(r['kernel_ms'],3), 'frac', round(r.get('frac') or 0,3), '| fwd_bwd', round(fb.get('ms_per_step') or 0,2), 'goal', round(goal.get('ms_per_step') or 0,2), 'adj_ms', round(adj.get('kernel_ms') or 0,3), 'adj_frac', round(adj.get('frac') or 0,3))
except Exception as e:
    print('unreadable', e)
PY
)"
  done
done
unset RTT_BWD_MINB
