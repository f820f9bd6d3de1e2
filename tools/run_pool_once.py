import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "rectified-spaattn_b200"))
sys.argv = [sys.argv[0]]
import torch, bench
from rsa_b200 import ops
dev = torch.device("cuda:0")
wp = bench.workload_params("c3b")
q, k, v = bench.synth_heads_device(wp["heads"], 0, wp["s"], "walk", dev)
plan = ops.Plan(q, k, v, bench.product_geometry(wp), wp["top_k"], bench.P_REMAIN, ops.gilbert_block_neighbors(*wp["grid"]))
for _ in range(4):
    plan.pool_stats()
torch.cuda.synchronize()
