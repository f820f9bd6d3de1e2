# session 8, call D: where does the third skeleton wait?  ncu --set full with source for both skeletons
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_tc5 -c 1 -f -o gpurun_out/s8d_k4_smemp python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-permute --no-reference-gpu > gpurun_out/s8d.log 2>&1
tail -2 gpurun_out/s8d.log
