# A/B inside one gpurun call: kernel 4 with attention flags 0 against $1 (parity under the flag first)
F=${1:-64}
RSA_ATTN_FLAGS=$F timeout 150 python -m pytest tests/test_gpu_parity.py -q -x -k "masked_attention or end_to_end_vs or edge or far_above or dense_limit or head_dim_64" 2>&1 | tail -2
for f in 0 $F 0 $F; do timeout 100 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --attn-flags $f 2>/dev/null | python tools/ab_line.py c3b_flags$f; done
for f in 0 $F; do timeout 100 python bench.py --workload c5 --steps 10 --no-e2e --no-cpu-baseline --no-permute --attn-flags $f 2>/dev/null | python tools/ab_line.py c5_flags$f; done
