# session 8, call B: does breaking the P-aliases-S dependency pay?  (flag 64 = timing only, results invalid; flag 32 = no clusters)
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "end_to_end_vs or edge or masked_attention or head_dim_64" 2>&1 | tail -1
for f in 32 96 32 96; do timeout 60 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c3b_5stages_flags$f; done
for f in 32 96; do timeout 60 python bench.py --workload c5 --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c5_5stages_flags$f; done
cp tools/_build/librsa_b200_s4.so rectified-spaattn_b200/rsa_b200/librsa_b200.so
for f in 32 96 32 96; do timeout 60 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c3b_4stages_flags$f; done
timeout 120 ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active --clock-control none -k regex:attn_tc5 -c 2 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags 96 2>&1 | grep -E "cycles_elapsed|time_duration|pipe_tensor|pipe_xu"
