#!/bin/bash
# One ncu --set full capture of one kernel of one bench command, summarised ON THE BOX.
# Usage: gpu_profile_any.sh <tag> <name> <kernel regex> <mangled substring for the line map> <skip> <bench args...>
set -u
TAG="$1"; NAME="$2"; KREGEX="$3"; KSUB="$4"; SKIP="$5"; shift 5
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py $*"
$CMD > $OUT/plain_${NAME}_$TAG.log 2>&1 || { echo "plain run failed: $NAME"; tail -20 $OUT/plain_${NAME}_$TAG.log; exit 1; }
REP=$OUT/prof_${NAME}_$TAG
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c 1 -f -o $REP $CMD > $OUT/ncu_full_${NAME}_$TAG.log 2>&1
echo "ncu full $NAME exit $?"
python scripts/ncu_summary.py $REP.ncu-rep > $OUT/sum_${NAME}_$TAG.txt 2>&1
python scripts/ncu_by_func.py $REP.ncu-rep $KSUB ${VARIANT:-fast} > $OUT/func_${NAME}_$TAG.txt 2>&1
python scripts/ncu_by_line.py $REP.ncu-rep $KSUB ${VARIANT:-fast} 60 --by-samples > $OUT/samples_${NAME}_$TAG.txt 2>&1
python scripts/ncu_by_line.py $REP.ncu-rep $KSUB ${VARIANT:-fast} 60 > $OUT/lines_${NAME}_$TAG.txt 2>&1
ncu -i $REP.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for k,v in zip(h,r):
        if 'stall' in k or 'pipe' in k or 'inst_executed' in k or 'issue' in k or 'local' in k: print(k, v)
" > $OUT/raw_${NAME}_$TAG.txt 2>&1
rm -f $REP.ncu-rep
head -20 $OUT/sum_${NAME}_$TAG.txt
