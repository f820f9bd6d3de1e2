# session 8, call C: bring-up of kernel 4's third skeleton (P through shared memory)
timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "masked_attention" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "end_to_end_vs or edge or head_dim_64 or kernel4" 2>&1 | tail -15
for f in 64 0 64 0; do timeout 60 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c3b_flags$f; done
