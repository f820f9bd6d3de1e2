"""Dense mode of the path (run on the GPU box): `fullattn(mode="flash")` -- what the reference processors call on warm-up
steps / layers and for text rows (attn.py:60-120 -> flash_attn_varlen_func) -- runs here on kernel 4 with every block
kept.  Times it at the C3a token count on 8 heads against flash-attn 2 (the reference's dependency, if importable) and
PyTorch SDPA on the same tensors.  One JSON line."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "rectified-spaattn_b200")]
from rectified_spaattn.attn import fullattn  # noqa: E402

dev = torch.device("cuda:0")
heads, s, valid = 8, 115456, 115400
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.randn(1, heads, s, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(3))
cu = [0, valid, s]
flop = 4.0 * valid * valid * 128 * heads


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


ms_ours, o_ours = timed(lambda: fullattn(q, k, v, mode="flash", cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=s,
                                         max_seqlen_kv=s, batch_size=1))
res = {"shape": f"1 x {heads} heads x {s} tokens ({valid} valid) x 128", "dense_tflop": flop / 1e12,
       "ours_kernel4_dense_ms": ms_ours, "ours_tflops": flop / ms_ours / 1e9}
qv, kv_, vv = (t[:, :, :valid] for t in (q, k, v))
ms_sdpa, o_sdpa = timed(lambda: torch.nn.functional.scaled_dot_product_attention(qv, kv_, vv))
res.update(sdpa_ms=ms_sdpa, sdpa_tflops=flop / ms_sdpa / 1e9,
           max_abs_vs_sdpa=float((o_ours[:, :, :valid].float() - o_sdpa.float()).abs().max()))
try:
    from flash_attn import flash_attn_func
    qf, kf, vf = (t.transpose(1, 2).contiguous() for t in (qv, kv_, vv))
    ms_fa, _ = timed(lambda: flash_attn_func(qf, kf, vf))
    res.update(flash_attn2_ms=ms_fa, flash_attn2_tflops=flop / ms_fa / 1e9)
except Exception as e:  # noqa: BLE001
    res["flash_attn2"] = f"unavailable: {type(e).__name__}: {e}"
print(json.dumps(res))
