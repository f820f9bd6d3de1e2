"""How many (32-row warp, kept block) steps of kernel 4 hold only negligible probabilities?  For sampled query tiles of a
bench.py workload: S = q~ K^T over the tile's kept blocks in the order kernel 4 walks them, running row maximum, and the
share of (warp, block) steps in which every row's block maximum lies more than T (log2 units) below its running maximum
-- every p of such a step is < 2^-T of what the row has already accumulated."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "rectified-spaattn_b200")]
name = sys.argv[1] if len(sys.argv) > 1 else "c3b"
regime = sys.argv[2] if len(sys.argv) > 2 else "walk"
sys.argv = [sys.argv[0]]
import bench  # noqa: E402
from rsa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
wp = bench.workload_params(name)
geo = bench.product_geometry(wp)
nbr = ops.gilbert_block_neighbors(*wp["grid"])
heads = 2
q, k, v = bench.synth_heads_device(heads, 0, wp["s"], regime, dev)
plan = ops.Plan(q, k, v, geo, wp["top_k"], bench.P_REMAIN, nbr)
plan.run()
torch.cuda.synchronize()
vw = plan.view()
cnt = vw["kept_cnt"].cpu()
sched = vw["sched_idx"].cpu().long() & 0xFFFF
nq = geo.nq_blocks
scale = 128 ** -0.5 * 1.44269504
res = {T: [0, 0] for T in (24, 32, 40, 48, 64)}
tiles_all = {T: [0, 0] for T in res}
g = torch.Generator().manual_seed(0)
for hi in range(heads):
    for tile in torch.randint(0, nq - 1, (24,), generator=g).tolist():
        n = int(cnt[hi, tile])
        blocks = sched[hi, tile, :n].to(dev)
        qt = (q[0, hi, tile * 128: tile * 128 + 128].float() * scale).to(torch.bfloat16).float()
        rows = (blocks[:, None] * 128 + torch.arange(128, device=dev)[None]).flatten()
        rows = rows.clamp(max=wp["s"] - 1)
        s = (qt @ k[0, hi, rows].float().T).view(128, n, 128)
        bmax = s.max(dim=2).values                               # [128 rows, n blocks]
        run = torch.cummax(bmax, dim=1).values
        prev = torch.cat([bmax[:, :1], run[:, :-1]], dim=1)      # running maximum BEFORE the block
        gap = prev - bmax                                        # how far the block's maximum lies below it
        for T in res:
            neg = (gap > T).view(4, 32, n).all(dim=1)            # per warp
            neg[:, 0] = False
            res[T][0] += int(neg.sum()); res[T][1] += 4 * n
            allw = neg.all(dim=0)
            tiles_all[T][0] += int(allw.sum()); tiles_all[T][1] += n
print(json.dumps({"workload": name, "regime": regime,
                  "share_of_warp_block_steps_negligible": {T: r[0] / r[1] for T, r in res.items()},
                  "share_of_tile_block_steps_negligible_for_all_4_warps": {T: r[0] / r[1] for T, r in tiles_all.items()}}))
