"""Kernel 0 measurement (run on the GPU box): the reference processor's pre-attention op sequence in PyTorch eager
(rectified_hunyuan_attn.py:448-479, what `_processors.py` mirrors) against rsa_qkv_prep, on a bench.py workload.
Prints one JSON line: times, the HBM roofline of the fused kernel and what it saves per attention call."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "rectified-spaattn_b200")]
sys.argv, argv = [sys.argv[0]], sys.argv[1:]
import bench  # noqa: E402
from oracle import make_golden as MG  # noqa: E402  (torch_prep: the literal op sequence)
from rsa_b200 import ops  # noqa: E402

name = argv[0] if argv else "c3b"
wp = bench.workload_params(name)
dev = torch.device("cuda:0")
heads, s, nv = wp["heads"], wp["s"], wp["nv"]
g = torch.Generator(device=dev).manual_seed(0)
src = [torch.randn(1, s, heads * 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(3)]
wan = wp["fam"] == "wan"      # Wan: RMSNorm across heads, rotary embedding on every token (complex, float64 in the reference)
nw = heads * 128 if wan else 128
wq = (1 + 0.1 * torch.randn(nw, generator=g, device=dev)).to(torch.bfloat16)
wk = (1 + 0.1 * torch.randn(nw, generator=g, device=dev)).to(torch.bfloat16)
ang = torch.outer(torch.arange(nv, dtype=torch.float32, device=dev),
                  1.0 / (256.0 ** (torch.arange(0, 128, 2, device=dev) / 128)))
cos, sin = ang.cos().repeat_interleave(2, 1).contiguous(), ang.sin().repeat_interleave(2, 1).contiguous()
freqs = torch.polar(torch.ones_like(ang, dtype=torch.float64), ang.double())[None, None] if wan else None
geo = bench.product_geometry(wp)
t, h, w = wp["grid"]
nbr = ops.gilbert_block_neighbors(t, h, w)
q, k, v = (torch.empty(1, heads, s, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
plan = ops.Plan(q, k, v, geo, wp["top_k"], bench.P_REMAIN, nbr)


def timed(fn, n):
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


def fused():
    if wp["text"] and geo.gap:          # ragged visual segment: two sources (dual-stream form)
        plan.qkv_prep(*(x[:, :nv] for x in src), dst_row=0, q_weight=wq, k_weight=wk, rope=(cos, sin))
        plan.qkv_prep(*(x[:, nv:] for x in src), dst_row=nv, q_weight=wq, k_weight=wk)
    else:
        plan.qkv_prep(*src, dst_row=0, q_weight=wq, k_weight=wk, rope=(cos, sin), rope_rows=nv)


def eager():
    if wan:
        tq = MG.torch_prep_wan(src[0], heads, wq, 1e-6, freqs)
        tk = MG.torch_prep_wan(src[1], heads, wk, 1e-6, freqs)
        return tq, tk, src[2].unflatten(2, (heads, -1)).transpose(1, 2)
    tq = MG.torch_prep(src[0], heads, wq, 1e-6, cos, sin, nv)
    tk = MG.torch_prep(src[1], heads, wk, 1e-6, cos, sin, nv)
    tv = src[2].unflatten(2, (heads, -1)).transpose(1, 2)
    return tq, tk, tv


ms_fused = timed(fused, 10)
ms_pool = timed(plan.pool_stats, 10)
ms_eager = timed(eager, 3)
tq, tk, tv = eager()
fused()
torch.cuda.synchronize()
same = float((q == tq).float().mean())
b_alg = 6 * s * heads * 128 * 2
peaks = bench.measured_peaks()
print(json.dumps({
    "workload": name, "tokens": s, "heads": heads, "fused_ms": ms_fused, "kernel2_alone_ms": ms_pool,
    "pytorch_eager_sequence_ms": ms_eager, "algorithmic_bytes": b_alg,
    "achieved_gbs": b_alg / (ms_fused * 1e-3) / 1e9, "peak_gbs": peaks["hbm"], "peak_source": peaks["source"],
    "frac": b_alg / (ms_fused * 1e-3) / 1e9 / peaks["hbm"],
    "q_identical_to_eager": same, "form": "wan (norm across heads, complex fp64 RoPE in the reference)" if wan else "joint (norm per head, cos/sin RoPE)",
    "note": "fused = head split + RMSNorm + RoPE + re-layout + block pooling; eager = the reference processor's op "
            "sequence (its V is a view; the attention call then runs kernel 2 on top)"}))
