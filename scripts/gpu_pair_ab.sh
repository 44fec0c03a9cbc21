#!/bin/bash
# A/B of the sequential forward builds on one box (run under gpurun): RTT_FWD_TILE selects the build
# (include/rtt_b200.h RTT_MODE_TUNE_*).  Usage: scripts/gpu_pair_ab.sh <tag> "<tunes>" "<workloads>"
set -u
TAG="${1:-ab}"; TUNES="${2:-3 5 16 17 18 19}"; WLS="${3:-c2 c1 c4}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernel_parity.py -m gpu -q -x -k "seq_matches_reference_within_tolerance or maximum_table" > $OUT/pytest_pair_$TAG.log 2>&1; echo "pytest pair exit $?"
tail -5 $OUT/pytest_pair_$TAG.log
for wl in $WLS; do
  for t in $TUNES; do
    RTT_FWD_TILE=$t timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-bwd --no-config4 --no-other-configs > $OUT/ab_${wl}_t${t}_$TAG.json 2> $OUT/ab_${wl}_t${t}_$TAG.err
    echo "$wl tune $t exit $? $(python -c "import json,sys; d=json.load(open('$OUT/ab_${wl}_t${t}_$TAG.json')); r=d['roofline']; print('ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'clk', d['clocks'].get('sm_mhz'))" 2>&1 | tail -1)"
  done
done
