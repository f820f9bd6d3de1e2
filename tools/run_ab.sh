# A/B inside one gpurun call: kernel 4 with one MMA issuer (flags 0) and two (flags 32)
RSA_ATTN_FLAGS=32 timeout 150 python -m pytest tests/test_gpu_parity.py -q -x -k "masked_attention or end_to_end_vs or edge or far_above or dense_limit" 2>&1 | tail -2
for f in 0 32 0 32; do timeout 100 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --attn-flags $f 2>/dev/null | python tools/ab_line.py c3b_flags$f; done
for f in 0 32; do timeout 100 python bench.py --workload c5 --steps 10 --no-e2e --no-cpu-baseline --no-permute --attn-flags $f 2>/dev/null | python tools/ab_line.py c5_flags$f; done
