#!/bin/bash
# ncu full capture of the non-sequential forward kernel (C5).  Usage: gpu_profile_nonseq.sh <tag> [rays]
set -u
TAG="$1"; RAYS="${2:-10000000}"
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload c5 --rays $RAYS --steps 2 --warmup 1 --no-e2e --no-cpu --no-bwd"
$CMD > $OUT/plain_c5_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_c5_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_trace_nonseq_fwd -s 2 -c 1 -f -o $OUT/prof_c5_$TAG $CMD > $OUT/ncu_full_c5_$TAG.log 2>&1
echo "ncu full c5 exit $?"
