VARIANT=exact bash scripts/gpu_profile_any.sh p8 c5fwd k_trace_nonseq_fwd k_trace_nonseq_fwd_exact 2 --workload c5 --rays 10000000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-bwd
bash scripts/gpu_profile_any.sh p8 c1adj k_trace_seq_bwd k_trace_seq_bwd_fastILi4ELb0 1 --workload c1 --rays 20000000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-config4 --no-other-configs
bash scripts/gpu_profile_any.sh p8 c4adj k_trace_seq_bwd k_trace_seq_bwd_fastILi4ELb0 1 --workload c4 --rays 20000000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-config4 --no-other-configs
python -m pytest tests/test_kernel_parity.py -m gpu -q -x -k "maximum_rows_times" 2>&1 | tail -2
