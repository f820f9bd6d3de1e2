"""Is a READ-ONLY HBM stream slower right after a power-capped tensor kernel?  torch.sum over 1 GiB of bf16 / fp32, cold
and hot (after 20 full attention calls), next to a read+write copy and kernel 2."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "rectified-spaattn_b200"))
sys.argv = [sys.argv[0]]
import torch
import bench
from rsa_b200 import ops
dev = torch.device("cuda:0")
wp = bench.workload_params("c3b")
q, k, v = bench.synth_heads_device(wp["heads"], 0, wp["s"], "walk", dev)
plan = ops.Plan(q, k, v, bench.product_geometry(wp), wp["top_k"], bench.P_REMAIN, ops.gilbert_block_neighbors(*wp["grid"]))
x = torch.randn(1 << 28, device=dev)          # 1 GiB fp32
y = torch.empty_like(x)
def t(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def heat():
    for _ in range(20):
        plan.run()
res = {}
for name, fn, byts in (("sum_fp32", lambda: x.sum(), x.numel() * 4), ("max_fp32", lambda: x.max(), x.numel() * 4),
                       ("copy", lambda: y.copy_(x), 2 * x.numel() * 4), ("pool_stats", plan.pool_stats, 3 * q.numel() * 2)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    cold = t(fn, 30)
    heat()
    hot = t(fn, 30)
    res[name] = {"cold_GBps": byts / cold / 1e6, "hot_GBps": byts / hot / 1e6, "hot_over_cold": cold / hot}
    torch.cuda.synchronize()
    import time; time.sleep(1.0)
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(REPO, "gpurun_out", "hot_readonly.json"), "w"), indent=1)
