#!/bin/bash
# Build an A/B variant of the library next to the shipped one: scripts/build_variant.sh <name> "<extra nvcc flags>"
# -> raytracetorch_b200/variants/librtt_b200_<name>.so (git-ignored: *.so), objects in csrc/build_<name>/.
set -euo pipefail
NAME="$1"; EXTRA="$2"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/../raytracetorch_b200/csrc" && pwd)"
OUT="$HERE/../variants"; B="$HERE/build_$NAME"
mkdir -p "$OUT" "$B"
# compile a SNAPSHOT of the sources, so that csrc/ can be edited while a variant builds
SROOT="${TMPDIR:-/tmp}/rtt_variant_$NAME"; SNAP="$SROOT/raytracetorch_b200/csrc"; rm -rf "$SROOT"; mkdir -p "$SNAP" "$SROOT/include"
cp "$HERE"/*.cu "$HERE"/*.cuh "$HERE"/*.inl "$HERE"/*.h "$SNAP"/
cp "$HERE"/../../include/*.h "$SROOT/include/"
SRC="$SNAP"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH $EXTRA"
$NVCC $COMMON -c "$SRC/rtt_kernels_fast.cu" -o "$B/fast.o" &
$NVCC $COMMON -fmad=false -c "$SRC/rtt_kernels_exact.cu" -o "$B/exact.o" &
$NVCC $COMMON -c "$SRC/rtt_cabi.cu" -o "$B/cabi.o" &
$NVCC $COMMON -c "$SRC/rtt_goals.cu" -o "$B/goals.o" &
wait
$NVCC -shared $ARCH -o "$OUT/librtt_b200_$NAME.so" "$B/fast.o" "$B/exact.o" "$B/cabi.o" "$B/goals.o" -lcudart
echo "built $OUT/librtt_b200_$NAME.so"
