#!/bin/bash
# A/B of forward kernel builds (RTT_FWD_TILE -> mode tune bits of rtt_trace_seq_fwd) on one box.
# Usage: gpu_fwd_tune_ab.sh <tag> "<workloads>" <tune> [<tune> ...]
TAG="$1"; WLS="$2"; shift 2
for wl in $WLS; do
  for t in "$@"; do
    if [ "$t" = default ]; then unset RTT_FWD_TILE; else export RTT_FWD_TILE=$t; fi
    timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-bwd --no-config4 --no-other-configs > gpurun_out/ft_${wl}_${t}_$TAG.json 2> gpurun_out/ft_${wl}_${t}_$TAG.err
    echo "$wl tune $t exit $? $(python -c "
import json
d=json.load(open('gpurun_out/ft_${wl}_${t}_$TAG.json')); print('ms', round(d['ms_per_step'],3), 'kernel', d['roofline'].get('kernel'))")"
  done
done
unset RTT_FWD_TILE
