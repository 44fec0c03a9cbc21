#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): 2-rank correctness test, weak / strong scaling bench lines, C3 with captured NCCL.
set -u
TAG="${1:-m1}"; N="${2:-2}"
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_$TAG.txt 2>&1
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x > $OUT/pytest_multigpu_$TAG.log 2>&1; echo "pytest multigpu exit $?"; tail -15 $OUT/pytest_multigpu_$TAG.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_c2_n${N}_$TAG.json 2> $OUT/bench_c2_n${N}_$TAG.err; echo "weak c2 n=$N exit $?"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-e2e --no-bwd --no-config4 --no-other-configs > $OUT/bench_c2_strong_n${N}_$TAG.json 2> $OUT/bench_c2_strong_n${N}_$TAG.err; echo "strong c2 n=$N exit $?"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-overlap --no-e2e --no-bwd --no-config4 --no-other-configs > $OUT/bench_c2_nooverlap_n${N}_$TAG.json 2> $OUT/bench_c2_nooverlap_n${N}_$TAG.err; echo "no-overlap c2 n=$N exit $?"
timeout 300 $TR bench.py --gpus $N --workload c3 --steps 20 --warmup 3 > $OUT/bench_c3_n${N}_$TAG.json 2> $OUT/bench_c3_n${N}_$TAG.err; echo "c3 eager-multi n=$N exit $?"
timeout 300 $TR bench.py --gpus $N --workload c3 --steps 20 --warmup 3 --graph-multi > $OUT/bench_c3_graph_n${N}_$TAG.json 2> $OUT/bench_c3_graph_n${N}_$TAG.err; echo "c3 graph-multi n=$N exit $?"
for f in $OUT/bench_*_$TAG.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print({k:d.get(k) for k in ("n_gpus","scaling","ms_per_step","value")}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "config4", {k:(d.get("config4") or {}).get(k) for k in ("ms_per_step","value","error")})
except Exception as e:
    print("unreadable", e)
PY
done
