# last-heads-first grid order: parity, A/B against the former order (flag 16), DRAM bytes of kernel 4
timeout 100 python -m pytest tests/test_gpu_parity.py -q -x -k "end_to_end_vs or edge or masked_attention or head_dim_64 or ulysses or peer" 2>&1 | tail -1
for f in 0 16; do timeout 60 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --attn-flags $f 2>/dev/null | python tools/ab_line.py c3b_flags$f; done
for f in 0 16; do timeout 60 python bench.py --workload c5 --steps 10 --no-e2e --no-cpu-baseline --no-permute --attn-flags $f 2>/dev/null | python tools/ab_line.py c5_flags$f; done
timeout 100 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:attn_tc5 -c 1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-permute 2>&1 | grep -E "dram__bytes|gpu__time"
