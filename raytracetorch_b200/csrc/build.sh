#!/bin/bash
# Builds librtt_b200.so (sm_100a) in-tree.  Usage: build.sh [outdir]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${1:-$HERE/..}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH ${RTT_NVCC_EXTRA:-}"
mkdir -p "$HERE/build"
$NVCC $COMMON -Xptxas -v -c "$HERE/rtt_kernels_fast.cu"  -o "$HERE/build/rtt_kernels_fast.o"  2> "$HERE/build/ptxas_fast.log"  || { cat "$HERE/build/ptxas_fast.log"; exit 1; }
$NVCC $COMMON -fmad=false -Xptxas -v -c "$HERE/rtt_kernels_exact.cu" -o "$HERE/build/rtt_kernels_exact.o" 2> "$HERE/build/ptxas_exact.log" || { cat "$HERE/build/ptxas_exact.log"; exit 1; }
$NVCC $COMMON -c "$HERE/rtt_cabi.cu" -o "$HERE/build/rtt_cabi.o"
$NVCC $COMMON -c "$HERE/rtt_goals.cu" -o "$HERE/build/rtt_goals.o"
$NVCC -shared $ARCH -o "$OUT/librtt_b200.so" "$HERE/build/rtt_kernels_fast.o" "$HERE/build/rtt_kernels_exact.o" "$HERE/build/rtt_cabi.o" "$HERE/build/rtt_goals.o" -lcudart
echo "built $OUT/librtt_b200.so"
