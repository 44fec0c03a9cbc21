#!/bin/bash
# Round-2 GPU-box visit: smoke, GPU parity tests, the default bench line and a short reference arm (run under gpurun).
# Usage: scripts/gpu_r2_check.sh [tag] [pytest -k expression]
set -u
TAG="${1:-r2}"
KEXPR="${2:-}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?" | tee -a $OUT/status_$TAG.txt
tail -2 $OUT/smoke_$TAG.log
if [ -n "$KEXPR" ]; then
  timeout 1800 python -m pytest tests -m gpu -q -k "$KEXPR" > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt
else
  timeout 1800 python -m pytest tests -m gpu -q --durations=15 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/status_$TAG.txt
fi
tail -40 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 4 > $OUT/bench_c2_$TAG.json 2> $OUT/bench_c2_$TAG.err; echo "bench c2 exit $?" | tee -a $OUT/status_$TAG.txt
cat $OUT/bench_c2_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-seconds 2 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref exit $?" | tee -a $OUT/status_$TAG.txt
cat $OUT/bench_ref_$TAG.json
