# session 8, call I: speculative first-half exponentials under the row-maximum pass (attention flag 64), hand-over pinned
RSA_ATTN_FLAGS=64 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "end_to_end_vs or masked_attention or edge or kernel4 or overflow" 2>&1 | tail -2
for f in 0 64 0 64; do timeout 60 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c3b_flags$f; done
for f in 0 64; do timeout 60 python bench.py --workload c5 --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c5_flags$f; done
for f in 0 64; do timeout 60 python bench.py --workload c2 --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu --attn-flags $f 2>/dev/null | python tools/ab_line.py c2_flags$f; done
