bash scripts/gpu_profile_any.sh p6 c3adj k_trace_seq_bwd k_trace_seq_bwd_fastILi4ELb0 2 --workload c3 --rays 10000000 --steps 2 --warmup 1 --no-graph --no-cpu
bash scripts/gpu_profile_any.sh p6 c2adj k_trace_seq_bwd k_trace_seq_bwd_fastILi4ELb0 1 --workload c2 --rays 20000000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-config4 --no-other-configs
bash scripts/gpu_profile_any.sh p6 c4fwd k_trace_seq_fwd k_trace_seq_fwd_tileILi2ELi4 3 --workload c4 --rays 20000000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-bwd
bash scripts/gpu_profile_any.sh p6 c4cam k_trace_seq_fwd k_trace_seq_fwd_tileILi2ELi4 3 --workload c4cam --rays 20000000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-bwd
