# Round-end measurement set on one B200 (run through gpurun): GPU tests, the default bench line, every workload, the ncu
# launch list of the bench command and one `ncu --set full` capture of each kernel of the call.
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py > gpurun_out/s6_final_c3b.json 2> gpurun_out/s6_final_c3b.err
for w in c1 c2 c3a c4 c5; do timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline > gpurun_out/s6_all_$w.json 2>/dev/null; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s6_launches_c3b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-permute > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"attn_tc5|pool_stats|block_s|rect_c|pair_schedule" -c 6 -f -o gpurun_out/s6_prof_c3b python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-permute > gpurun_out/s6_prof.log 2>&1
ls gpurun_out | grep s6_ | tr '\n' ' '
