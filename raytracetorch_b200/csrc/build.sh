#!/bin/bash
# Builds librtt_b200.so (sm_100a) in-tree.  Usage: build.sh [outdir]
# The four translation units compile in parallel (the two kernel variants dominate: ~70 s each).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${1:-$HERE/..}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH ${RTT_NVCC_EXTRA:-}"
mkdir -p "$HERE/build"
$NVCC $COMMON -Xptxas -v -c "$HERE/rtt_kernels_fast.cu"  -o "$HERE/build/rtt_kernels_fast.o"  2> "$HERE/build/ptxas_fast.log" & P1=$!
$NVCC $COMMON -fmad=false -Xptxas -v -c "$HERE/rtt_kernels_exact.cu" -o "$HERE/build/rtt_kernels_exact.o" 2> "$HERE/build/ptxas_exact.log" & P2=$!
$NVCC $COMMON -c "$HERE/rtt_cabi.cu" -o "$HERE/build/rtt_cabi.o" & P3=$!
$NVCC $COMMON -c "$HERE/rtt_goals.cu" -o "$HERE/build/rtt_goals.o" & P4=$!
wait $P1 || { cat "$HERE/build/ptxas_fast.log"; exit 1; }
wait $P2 || { cat "$HERE/build/ptxas_exact.log"; exit 1; }
wait $P3
wait $P4
$NVCC -shared $ARCH -o "$OUT/librtt_b200.so" "$HERE/build/rtt_kernels_fast.o" "$HERE/build/rtt_kernels_exact.o" "$HERE/build/rtt_cabi.o" "$HERE/build/rtt_goals.o" -lcudart
echo "built $OUT/librtt_b200.so"
