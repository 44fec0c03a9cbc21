#!/bin/bash
# A/B of the lock-step builds of the non-sequential forward (RTT_NS_TUNE -> mode tune bits of rtt_trace_nonseq_fwd), one box.
# 0 = default kernel; 1 = 256 threads, barrier per trip; 2 = 1024, per trip; 3 = 1024, per trip and per row; 4 / 5 = 512.
TAG="$1"; shift
for t in "$@"; do
  export RTT_NS_TUNE=$t
  timeout 300 python bench.py --workload c5 --steps 3 --warmup 2 --no-cpu --no-e2e --no-config4 --no-other-configs > gpurun_out/ns_${t}_$TAG.json 2> gpurun_out/ns_${t}_$TAG.err
  echo "ns tune $t exit $? $(python -c "
import json
d=json.load(open('gpurun_out/ns_${t}_$TAG.json')); print('ms', round(d['ms_per_step'],2), 'fast', round((d.get('nonseq_fast') or {}).get('kernel_ms') or 0,2))")"
done
unset RTT_NS_TUNE
