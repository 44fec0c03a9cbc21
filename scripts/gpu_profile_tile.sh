#!/bin/bash
# ncu full capture of the forward kernel for one workload / tile choice.  Usage: gpu_profile_tile.sh <tag> <workload> <tile> [rays]
set -u
TAG="$1"; WL="$2"; TILE="$3"; RAYS="${4:-20000000}"
OUT=gpurun_out; mkdir -p $OUT
export RTT_FWD_TILE=$TILE
CMD="python bench.py --workload $WL --rays $RAYS --steps 2 --warmup 1 --no-e2e --no-cpu --no-bwd"
$CMD > $OUT/plain_${WL}_t${TILE}_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_${WL}_t${TILE}_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_trace_seq_fwd -s 3 -c 1 -f -o $OUT/prof_${WL}_t${TILE}_$TAG $CMD > $OUT/ncu_full_${WL}_t${TILE}_$TAG.log 2>&1
echo "ncu full $WL tile=$TILE exit $?"
