#!/bin/bash
# Forward-only A/B of library variants on one box.  Usage: gpu_ab_fwd.sh <tag> "<variant names>" "<workloads>" [reps]
# Every (workload, variant) is run <reps> times, interleaved, so that drift of the box shows as spread, not as a winner.
set -u
TAG="$1"; VARS="$2"; WLS="$3"; REPS="${4:-2}"
OUT=gpurun_out; mkdir -p $OUT
for rep in $(seq 1 $REPS); do
for wl in $WLS; do
  for v in $VARS; do
    if [ "$v" = shipped ]; then unset RTT_B200_LIB; else export RTT_B200_LIB=$PWD/raytracetorch_b200/variants/librtt_b200_$v.so; fi
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e --no-bwd --no-config4 --no-other-configs > $OUT/abf_${wl}_${v}_${TAG}_$rep.json 2> $OUT/abf_${wl}_${v}_${TAG}_$rep.err
    echo "$wl $v rep$rep exit $? $(python - <<PY
import json
try:
    d=json.load(open('$OUT/abf_${wl}_${v}_${TAG}_$rep.json')); r=d['roofline']
    print('ms', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4))
except Exception as e:
    print('unreadable', e)
PY
)"
  done
done
done
