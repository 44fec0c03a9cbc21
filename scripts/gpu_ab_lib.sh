#!/bin/bash
# A/B of library variants on one box: for each workload, the shipped library and raytracetorch_b200/variants/librtt_b200_<name>.so.
# Usage: gpu_ab_lib.sh <tag> "<variant names>" "<workloads>" [extra bench args]
set -u
TAG="$1"; VARS="$2"; WLS="$3"; EXTRA="${4:-}"
OUT=gpurun_out; mkdir -p $OUT
for wl in $WLS; do
  for v in shipped $VARS; do
    if [ "$v" = shipped ]; then unset RTT_B200_LIB; else export RTT_B200_LIB=$PWD/raytracetorch_b200/variants/librtt_b200_$v.so; fi
    timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-config4 --no-other-configs $EXTRA > $OUT/ablib_${wl}_${v}_$TAG.json 2> $OUT/ablib_${wl}_${v}_$TAG.err
    echo "$wl $v exit $? $(python - <<PY
import json
try:
    d=json.load(open('$OUT/ablib_${wl}_${v}_$TAG.json')); r=d['roofline']; fb=d.get('fwd_bwd') or {}
    adj=(fb.get('adjoint') or {})
    print('ms', round(d['ms_per_step'],3), 'kernel_ms', round(r['kernel_ms'],3), '| fwd_bwd', round(fb.get('ms_per_step') or 0,2), 'adj_ms', round(adj.get('kernel_ms') or 0,3), '| nsf', round((d.get('nonseq_fast') or {}).get('kernel_ms') or 0,2))
except Exception as e:
    print('unreadable', e)
PY
)"
  done
done
