# session 8, call G: the final library: GPU suite, smoke, default bench line (with the kernel 0 leg), reference arm
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_c3b_final.json 2> gpurun_out/r02_bench_c3b_final.err; tail -c 600 gpurun_out/r02_bench_c3b_final.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 800 gpurun_out/r02_bench_reference_arm.json
