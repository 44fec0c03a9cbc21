#!/bin/bash
# Several ncu --set full captures of the sequential forward kernel in one box visit.
# Usage: gpu_profile_many.sh <tag> "<wl:tune[:keep][:kernel-substring]> ..." [rays]
set -u
TAG="$1"; LIST="$2"; RAYS="${3:-20000000}"
for item in $LIST; do
  IFS=: read -r WL TILE KEEP KSUB <<< "$item"
  bash scripts/gpu_profile_tile.sh "$TAG" "$WL" "$TILE" "$RAYS" "${KEEP:-0}" "${KSUB:-k_trace_seq_fwd}"
done
du -sh gpurun_out
