export V=$PWD/raytracetorch_b200/variants/librtt_b200_i.so
RTT_B200_LIB=$V timeout 70 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
for wl in c4 c2; do for v in shipped i; do
  if [ $v = shipped ]; then unset RTT_B200_LIB; else export RTT_B200_LIB=$V; fi
  timeout 40 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-bwd --no-config4 --no-other-configs > gpurun_out/abf_${wl}_${v}_s5.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/abf_${wl}_${v}_s5.json'));print('$wl $v',round(d['ms_per_step'],4))"
done; done
