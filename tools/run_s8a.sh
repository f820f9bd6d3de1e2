# session 8, call A: GPU suite on the current library, FP8 ceiling proxy, packed 16-bit ex2 rate, POLY share with clusters
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/s8a_gputests.log
timeout 200 python tools/fp8_ceiling.py 2>/dev/null | tail -1 > gpurun_out/r02_fp8_ceiling_proxy.json
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/xu2 tools/xu_rate2.cu && /tmp/xu2 > gpurun_out/r02_xu_rate_packed_ex2.txt 2>&1
for p in 0 1 2 0 1; do RSA_TC5_POLY=$p timeout 60 python bench.py --steps 10 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu 2>/dev/null | python tools/ab_line.py c3b_poly$p; done > gpurun_out/s8a_poly_ab.txt
cat gpurun_out/s8a_gputests.log gpurun_out/s8a_poly_ab.txt gpurun_out/r02_xu_rate_packed_ex2.txt
