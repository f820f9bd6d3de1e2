# stage times after a kernel change, with the GPU test suite first (one gpurun call)
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for w in c3b c5 c4 c2; do timeout 100 python bench.py --workload $w --steps 10 --no-e2e --no-cpu-baseline --no-permute 2>/dev/null | python tools/stage_line.py $w; done
