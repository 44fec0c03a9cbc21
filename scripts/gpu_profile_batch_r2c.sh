# ncu captures of the final round-2 kernels (defaults after the block-size A/B and the lean adjoint)
F="--steps 2 --warmup 1 --no-e2e --no-cpu --no-config4 --no-other-configs"
bash scripts/gpu_profile_any.sh q1 c2fwd k_trace_seq_fwd k_trace_seq_fwd_tileILi2ELi1ELb0ELi1024 3 --workload c2 --rays 20000000 $F --no-bwd
bash scripts/gpu_profile_any.sh q1 c1fwd k_trace_seq_fwd k_trace_seq_fwd_tileILi2ELi1ELb0ELi1024 3 --workload c1 --rays 20000000 $F --no-bwd
bash scripts/gpu_profile_any.sh q1 c4fwd k_trace_seq_fwd k_trace_seq_fwd_tileILi2ELi1ELb0ELi1024 3 --workload c4 --rays 20000000 $F --no-bwd
bash scripts/gpu_profile_any.sh q1 c4cam k_trace_seq_fwd k_trace_seq_fwd_tileILi2ELi1ELb1ELi1024 3 --workload c4cam --rays 20000000 $F --no-bwd
bash scripts/gpu_profile_any.sh q1 c2adj k_trace_seq_bwd k_trace_seq_bwd_fastILi1ELb0ELi24ELb1ELi1ELi1024 1 --workload c2 --rays 40000000 $F
bash scripts/gpu_profile_any.sh q1 c4adj k_trace_seq_bwd k_trace_seq_bwd_fastILi1ELb0ELi24ELb1ELi1ELi1024 1 --workload c4 --rays 40000000 $F
bash scripts/gpu_profile_any.sh q1 c1adj k_trace_seq_bwd k_trace_seq_bwd_fastILi1ELb0ELi24ELb1ELi1ELi1024 1 --workload c1 --rays 40000000 $F
bash scripts/gpu_profile_any.sh q1 c3adj k_trace_seq_bwd k_trace_seq_bwd_fastILi4ELb0ELi24ELb1ELi2ELi256 2 --workload c3 --rays 10000000 --steps 2 --warmup 1 --no-graph --no-cpu
VARIANT=exact bash scripts/gpu_profile_any.sh q1 c5fwd k_trace_nonseq_fwd k_trace_nonseq_fwd_ls_exactILi1024ELi1 2 --workload c5 --rays 10000000 $F --no-bwd
