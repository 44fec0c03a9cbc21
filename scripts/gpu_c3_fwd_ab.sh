for t in default 5 3 1 18; do
  if [ $t = default ]; then unset RTT_FWD_TILE; else export RTT_FWD_TILE=$t; fi
  python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu > gpurun_out/c3fwd_$t.json 2>gpurun_out/c3fwd_$t.err
  python -c "
import json; d=json.load(open('gpurun_out/c3fwd_$t.json')); print('c3 fwd build $t ms', round(d['ms_per_step'],4))"
done
unset RTT_FWD_TILE
