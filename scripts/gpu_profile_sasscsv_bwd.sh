#!/bin/bash
# ncu --set full capture of the sequential adjoint kernel; the per-SASS-instruction source page comes back as CSV.
# Usage: gpu_profile_sasscsv_bwd.sh <tag> <workload> [rays]
set -u
TAG="$1"; WL="$2"; RAYS="${3:-4000000}"
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload $WL --rays $RAYS --steps 2 --warmup 1 --no-e2e --no-cpu --no-config4 --no-other-configs"
$CMD > $OUT/plain_bwd_${WL}_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_bwd_${WL}_$TAG.log; exit 1; }
REP=$OUT/prof_bwd_${WL}_$TAG
ncu --set full --clock-control none --import-source on -k regex:k_trace_seq_bwd -s 1 -c 1 -f -o $REP $CMD > $OUT/ncu_bwd_${WL}_$TAG.log 2>&1
echo "ncu bwd $WL exit $?"
ncu -i $REP.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/sasscsv_bwd_${WL}_$TAG.csv.gz
python scripts/ncu_summary.py $REP.ncu-rep > $OUT/sum_bwd_${WL}_$TAG.txt 2>&1
rm -f $REP.ncu-rep
head -12 $OUT/sum_bwd_${WL}_$TAG.txt
