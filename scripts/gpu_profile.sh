#!/bin/bash
# ncu evidence for one workload (run under gpurun, ONE GPU): launch list + full capture of the
# sequential forward and adjoint kernels.  Usage: scripts/gpu_profile.sh <tag> [workload] [rays]
set -u
TAG="${1:-r1}"; WL="${2:-c2}"; RAYS="${3:-20000000}"
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload $WL --rays $RAYS --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > $OUT/plain_${WL}_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_${WL}_$TAG.log; exit 1; }
cat $OUT/plain_${WL}_$TAG.log | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_${WL}_$TAG.csv $CMD > $OUT/ncu_launches_${WL}_$TAG.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_trace_ -s 4 -c 3 -f -o $OUT/prof_${WL}_$TAG $CMD > $OUT/ncu_full_${WL}_$TAG.log 2>&1
echo "ncu full exit $?"
ls -la $OUT | tail -8
