"""head_dim 64 (CogVideoX1.5-5B, the reference's scripts/main_cogvideox.py defaults: 768x1280, 81 frames -> latent grid
11x48x80 = 42 240 visual tokens + 226 text tokens, 48 heads x 64, sa_drop_rate 0.75, p_remain_rates 0.3) on the GPU box.
Times the whole call (kernels 2-4) and kernel 4 alone through its 64-column instantiation and, for comparison, through
the 128-column instantiation reading the same tensors (TMA zero fill; attention flag bit 2), with RSA_TC5_POLY as set in
the environment.  One JSON line."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "rectified-spaattn_b200")]
import bench  # noqa: E402
from rsa_b200 import geometry as G  # noqa: E402
from rsa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
t, h, w, text, heads = 11, 48, 80, 226, 48
nv = t * h * w
s = nv + text
top_k = int((1 - 0.75) * (nv // 128))
geo = G.cogvideo(s, text)
nbr = ops.gilbert_block_neighbors(t, h, w)
q, k, v = bench.synth_heads_device(heads, 0, s, "walk", dev, d=64)
plan = ops.Plan(q, k, v, geo, top_k, 0.3, nbr, private_workspace=True)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {"shape": f"CogVideoX1.5 {t}x{h}x{w} + {text} text = {s} tokens x {heads} heads x 64, top_k {top_k}",
       "poly": os.environ.get("RSA_TC5_POLY", "default")}
outs = {}
for name, flag in (("d64", 0), ("d128_zero_fill", 4)):
    ops.set_attention_flags(flag)
    res[name + "_call_ms"] = timed(plan.run)
    res[name + "_kernel4_ms"] = timed(plan.sparse_attention)
    outs[name] = plan.run().clone()
ops.set_attention_flags(0)
kept = int(plan.view()["kept_cnt"].sum().item())
res["kept_pairs"] = kept
if kept:
    res["kept_density"] = kept / (heads * geo.n_blocks * geo.n_blocks)
    for name in ("d64", "d128_zero_fill"):
        res[name + "_tflops_on_kept_pairs_d64_flop"] = kept * 4.0 * 128 * 128 * 64 / res[name + "_kernel4_ms"] / 1e9
res["dense_equiv_tflops_d64"] = 4.0 * s * s * 64 * heads / res["d64_call_ms"] / 1e9
res["bit_identical"] = bool(torch.equal(outs["d64"].view(torch.int16), outs["d128_zero_fill"].view(torch.int16)))
res["max_abs_diff"] = float((outs["d64"].float() - outs["d128_zero_fill"].float()).abs().max())
print(json.dumps(res))
