#!/bin/bash
# The bench line of every workload (device-timed value, fwd+adjoint, e2e, CPU reference on rank 0), one JSON line each,
# appended to gpurun_out/bench_lines_<tag>.jsonl.  Usage: gpu_bench_lines.sh <tag> ["<workloads>"]
set -u
TAG="${1:-r2}"; WLS="${2:-c2 c1 c3 c4 c4cam c5}"
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/bench_lines_$TAG.jsonl
for wl in $WLS; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-seconds 4 > $OUT/line_${wl}_$TAG.json 2> $OUT/line_${wl}_$TAG.err
  echo "$wl exit $?"
  cat $OUT/line_${wl}_$TAG.json >> $OUT/bench_lines_$TAG.jsonl
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/line_reference_$TAG.json 2> $OUT/line_reference_$TAG.err; echo "reference exit $?"
cat $OUT/line_reference_$TAG.json >> $OUT/bench_lines_$TAG.jsonl
wc -l $OUT/bench_lines_$TAG.jsonl
