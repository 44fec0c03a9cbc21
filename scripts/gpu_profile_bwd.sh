#!/bin/bash
# ncu full capture of the sequential adjoint kernel for one workload.  Usage: gpu_profile_bwd.sh <tag> <workload> [rays]
set -u
TAG="$1"; WL="$2"; RAYS="${3:-20000000}"
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload $WL --rays $RAYS --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > $OUT/plain_bwd_${WL}_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_bwd_${WL}_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_trace_seq_bwd -s 1 -c 1 -f -o $OUT/prof_bwd_${WL}_$TAG $CMD > $OUT/ncu_bwd_${WL}_$TAG.log 2>&1
echo "ncu bwd $WL exit $?"
