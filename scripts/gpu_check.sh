#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench lines, ncu launch list (run under gpurun).
# Usage: scripts/gpu_check.sh [tag]
set -u
TAG="${1:-r1}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?" | tee -a $OUT/status_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_c2_$TAG.json 2> $OUT/bench_c2_$TAG.err; echo "bench c2 exit $?" | tee -a $OUT/status_$TAG.txt
cat $OUT/bench_c2_$TAG.json
for wl in c1 c3 c4 c4cam c5; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err; echo "bench $wl exit $?" | tee -a $OUT/status_$TAG.txt
  cat $OUT/bench_${wl}_$TAG.json
done
