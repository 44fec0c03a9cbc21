#!/bin/bash
# forward+adjoint step time for the adjoint kernel variants (run under gpurun)
set -u
TAG="${1:-bwd}"; OUT=gpurun_out; mkdir -p $OUT
for wl in c2 c4 c1; do
  for m in 2 3; do
    RTT_BWD_MINB=$m timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-e2e > $OUT/bwd_${wl}_m${m}_$TAG.json 2> $OUT/bwd_${wl}_m${m}_$TAG.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bwd_${wl}_m${m}_$TAG.json").read())
    print("$wl minb=$m fwd ms=%.3f fwd_bwd ms=%.3f value=%.4g" % (d["ms_per_step"], d["fwd_bwd"]["ms_per_step"], d["fwd_bwd"]["value"]))
except Exception as e:
    print("$wl minb=$m failed", e)
PY
  done
done
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu > $OUT/bwd_c3_$TAG.json 2> $OUT/bwd_c3_$TAG.err; python -c "
import json; d=json.loads(open('$OUT/bwd_c3_$TAG.json').read()); print('c3 ms', d['ms_per_step'], 'value %.4g' % d['value'], 'bwd kernel ms', d['roofline']['kernel_ms'])"
