# Round-2 final measurement set on one B200 (run through gpurun): GPU tests, the default bench line, every workload, the two
# other input regimes, kernel 0 and head_dim 64 tools, the ncu launch list of the bench command and one `ncu --set full`
# capture of every kernel of the call.
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_c3b.json 2> gpurun_out/r02_bench_c3b.err
for w in c1 c2 c3a c4 c5; do timeout 300 python bench.py --workload $w --steps 10 --no-cpu-baseline > gpurun_out/r02_all_$w.json 2>/dev/null; done
for r in iid cluster; do timeout 200 python bench.py --regime $r --steps 10 --no-cpu-baseline --no-e2e --no-reference-gpu --no-permute > gpurun_out/r02_regime_$r.json 2>/dev/null; done
for w in c3b c3a c4 c2 c5; do timeout 120 python tools/bench_prep.py $w 2>/dev/null | tail -1; done > gpurun_out/r02_bench_kernel0_prep.jsonl
timeout 200 python tools/bench_hd64.py 2>/dev/null | tail -1 > gpurun_out/r02_bench_hd64_cogvideox.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"attn_tc5|pool_stats|block_s|rect_c|pair_schedule|transpose" -c 7 -f -o gpurun_out/r02_prof_c3b_final python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu > gpurun_out/r02_prof.log 2>&1
ls gpurun_out | grep r02_ | tr '\n' ' '
