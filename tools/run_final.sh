# Round-2 measurement set on one B200 (run through gpurun): the default bench line, every workload, kernel 0, head_dim 64.
python bench.py > gpurun_out/r02_bench_c3b.json 2> gpurun_out/r02_bench_c3b.err
for w in c1 c2 c3a c4 c5; do timeout 300 python bench.py --workload $w --steps 10 --no-cpu-baseline > gpurun_out/r02_all_$w.json 2>/dev/null; done
for r in iid cluster; do timeout 200 python bench.py --regime $r --steps 10 --no-cpu-baseline --no-e2e --no-reference-gpu --no-permute > gpurun_out/r02_regime_$r.json 2>/dev/null; done
for w in c3b c3a c4 c2 c5; do timeout 120 python tools/bench_prep.py $w 2>/dev/null | tail -1; done > gpurun_out/r02_bench_kernel0_prep.jsonl
timeout 200 python tools/bench_hd64.py 2>/dev/null | tail -1 > gpurun_out/r02_bench_hd64_cogvideox.jsonl
ls gpurun_out | grep r02_ | tr '\n' ' '
