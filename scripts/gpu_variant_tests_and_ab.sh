#!/bin/bash
# One call that decides whether a library variant may ship: the whole `-m gpu` suite THROUGH the variant (RTT_B200_LIB),
# then a forward A/B against the shipped library.  Usage: gpu_variant_tests_and_ab.sh [variant name, default i]
# (call s5 of profiles/r2_block_size_ab.md ran this with variant i)
N="${1:-i}"
export V=$PWD/raytracetorch_b200/variants/librtt_b200_$N.so
RTT_B200_LIB=$V timeout 70 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
for wl in c4 c2; do for v in shipped $N; do
  if [ $v = shipped ]; then unset RTT_B200_LIB; else export RTT_B200_LIB=$V; fi
  timeout 40 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-bwd --no-config4 --no-other-configs > gpurun_out/abf_${wl}_${v}_$N.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/abf_${wl}_${v}_$N.json'));print('$wl $v',round(d['ms_per_step'],4))"
done; done
