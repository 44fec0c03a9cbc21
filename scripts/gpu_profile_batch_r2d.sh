# ncu captures of the default forward kernel after the instruction-count session (NARROW / LUT instantiations)
F="--steps 2 --warmup 1 --no-e2e --no-cpu --no-config4 --no-other-configs"
K=k_trace_seq_fwd_tileILi2ELi1ELb0ELi1024ELb0ELb1
bash scripts/gpu_profile_any.sh q2 c2fwd k_trace_seq_fwd ${K}ELi1 3 --workload c2 --rays 20000000 $F --no-bwd
bash scripts/gpu_profile_any.sh q2 c1fwd k_trace_seq_fwd ${K}ELi0 3 --workload c1 --rays 20000000 $F --no-bwd
bash scripts/gpu_profile_any.sh q2 c4fwd k_trace_seq_fwd ${K}ELi0 3 --workload c4 --rays 20000000 $F --no-bwd
bash scripts/gpu_profile_any.sh q2 c4cam k_trace_seq_fwd k_trace_seq_fwd_tileILi2ELi1ELb1ELi1024ELb0ELb1ELi0 3 --workload c4cam --rays 20000000 $F --no-bwd
