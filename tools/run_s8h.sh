# session 8, call H: polling instead of try_wait + suspend on kernel 4's critical-path waits (64: MMA warp, 128: softmax S wait)
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "end_to_end_vs or masked_attention" 2>&1 | tail -1
for f in 0 64 128 192 0 64 128 192; do timeout 60 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c3b_flags$f; done
for f in 0 64 192; do timeout 60 python bench.py --workload c5 --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c5_flags$f; done
