#!/bin/bash
# ncu full capture of the sequential forward kernel for one workload / kernel build, summarised ON THE BOX (the merged
# gpurun_out/ is limited to 64 MiB, a report is ~15 MB).  Usage: gpu_profile_tile.sh <tag> <workload> <tune> [rays] [keep-report 0|1] [kernel substring]
set -u
TAG="$1"; WL="$2"; TILE="$3"; RAYS="${4:-20000000}"; KEEP="${5:-0}"; KSUB="${6:-k_trace_seq_fwd}"
OUT=gpurun_out; mkdir -p $OUT
export RTT_FWD_TILE=$TILE
CMD="python bench.py --workload $WL --rays $RAYS --steps 2 --warmup 1 --no-e2e --no-cpu --no-bwd --no-config4 --no-other-configs"
$CMD > $OUT/plain_${WL}_t${TILE}_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_${WL}_t${TILE}_$TAG.log; exit 1; }
REP=$OUT/prof_${WL}_t${TILE}_$TAG
ncu --set full --clock-control none --import-source on -k regex:k_trace_seq_fwd -s 3 -c 1 -f -o $REP $CMD > $OUT/ncu_full_${WL}_t${TILE}_$TAG.log 2>&1
echo "ncu full $WL tune=$TILE exit $?"
python scripts/ncu_summary.py $REP.ncu-rep > $OUT/sum_${WL}_t${TILE}_$TAG.txt 2>&1
python scripts/ncu_by_func.py $REP.ncu-rep $KSUB fast > $OUT/func_${WL}_t${TILE}_$TAG.txt 2>&1
python scripts/ncu_by_line.py $REP.ncu-rep $KSUB fast 50 --by-samples > $OUT/samples_${WL}_t${TILE}_$TAG.txt 2>&1
python scripts/ncu_by_line.py $REP.ncu-rep $KSUB fast 50 > $OUT/lines_${WL}_t${TILE}_$TAG.txt 2>&1
ncu -i $REP.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for k,v in zip(h,r):
        if 'stall' in k or 'pipe' in k or 'inst_executed' in k or 'issue' in k: print(k, v)
" > $OUT/raw_${WL}_t${TILE}_$TAG.txt 2>&1
if [ "$KEEP" != "1" ]; then rm -f $REP.ncu-rep; fi
head -24 $OUT/sum_${WL}_t${TILE}_$TAG.txt
