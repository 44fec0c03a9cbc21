#!/bin/bash
# compute-sanitizer over scripts/sanitize_case.py, ONE tool per gpurun call.  Usage: gpu_sanitize.sh <tag> <memcheck|racecheck|synccheck|initcheck>
set -u
TAG="$1"; TOOL="$2"
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python scripts/sanitize_case.py > $OUT/sanitize_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -30 $OUT/sanitize_plain_$TAG.log; exit 1; }
tail -1 $OUT/sanitize_plain_$TAG.log
timeout 2400 compute-sanitizer --tool $TOOL --print-limit 50 python scripts/sanitize_case.py > $OUT/sanitize_${TOOL}_$TAG.log 2>&1
echo "compute-sanitizer $TOOL exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASE|hazard|Invalid|Error" $OUT/sanitize_${TOOL}_$TAG.log | head -40
