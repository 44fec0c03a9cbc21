#!/bin/bash
# Weak + strong scaling lines at N GPUs (gpurun --gpus N).  Usage: gpu_scale_check.sh <tag> <N>
set -u
TAG="${1:-s}"; N="${2:-4}"
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $OUT/scale_weak_n${N}_$TAG.json 2> $OUT/scale_weak_n${N}_$TAG.err; echo "weak n=$N exit $?"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-e2e --no-bwd --no-config4 --no-other-configs > $OUT/scale_strong_n${N}_$TAG.json 2> $OUT/scale_strong_n${N}_$TAG.err; echo "strong n=$N exit $?"
timeout 300 $TR bench.py --gpus $N --workload c3 --steps 20 --warmup 3 > $OUT/scale_c3_n${N}_$TAG.json 2> $OUT/scale_c3_n${N}_$TAG.err; echo "c3 n=$N exit $?"
for f in $OUT/scale_*_n${N}_$TAG.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    e=d.get("e2e") or {}
    print({k:d.get(k) for k in ("n_gpus","scaling","ms_per_step","value")}, "e2e", e.get("ms_per_step"), e.get("h2d_gb_per_s_per_rank"), "config4", {k:(d.get("config4") or {}).get(k) for k in ("ms_per_step","value","error")}, d["config"].get("numa_cores"))
except Exception as ex:
    print("unreadable", ex)
PY
done
