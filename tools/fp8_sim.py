"""SURVEY 8f rank 2 (FP8 Q/K/V path for kernel 4): accuracy study on the CPU before any kernel is written.

Block-sparse attention of one head on the three synthetic regimes of SURVEY 8d (random 25 % block mask + diagonal),
fp32 reference against e4m3 variants with per-128-token-block absmax scales (what kernel 2 could produce in the same
pass).  north_star's bar for the attention output is max-abs-err <= 2e-2 and cosine >= 0.999.  Result (see
profiles/r01_fp8_simulation.txt): no e4m3 variant meets it on the headline ("walk") regime -- the logits are large
there (|q||k| grows with the block centroids), so e4m3's 2^-4 relative step becomes an absolute error of several units
in the exponent; only a SageAttention-style two-sided mean removal with exact rank-1 corrections comes close, and that
is a different kernel.  Hence rank 2 is not built this round; the bf16 path stays the product.
Run: python tools/fp8_sim.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rsa_oracle as O  # noqa: E402  (synthetic inputs only)

E4M3_MAX = 448.0


def q8(x, blk=128):
    n, d = x.shape
    xb = x.view(n // blk, blk, d)
    sc = xb.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-12) / E4M3_MAX
    return ((xb / sc).to(torch.float8_e4m3fn).float() * sc).view(n, d)


def run(regime, seq=4096, dens=0.25):
    q, k, v = (torch.from_numpy(a[0, 0]) for a in O.synth_qkv(1, seq, 128, regime, 11))
    nb = seq // 128
    g = torch.Generator().manual_seed(0)
    m = (torch.rand(nb, nb, generator=g) < dens) | torch.eye(nb, dtype=torch.bool)
    big = m.repeat_interleave(128, 0).repeat_interleave(128, 1)
    scale = 128 ** -0.5

    def attn(s_unscaled, v_, p_mode):
        s = (s_unscaled * scale).masked_fill(~big, -1e30)
        p = torch.exp(s - s.max(-1, keepdim=True).values)
        l = p.sum(-1, keepdim=True)
        if p_mode == "bf16":
            p = p.bfloat16().float()
        else:                                  # e4m3 with the 2^8 offset trick (uses the format's range)
            p = (p * 256).to(torch.float8_e4m3fn).float() / 256
        return (p @ v_) / l

    ref_s = q @ k.T
    ref = (torch.softmax((ref_s * scale).masked_fill(~big, -1e30), -1)) @ v

    def rep(name, o):
        err = (o - ref).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(o.ravel(), ref.ravel(), dim=0).item()
        ok = "meets" if err <= 2e-2 and cos >= 0.999 else "FAILS"
        print(f"{regime:8s} {name:44s} max-abs {err:.4f}  cosine {cos:.6f}  {ok} the bar")

    rep("bf16 P (the product path's rounding)", attn(ref_s, v, "bf16"))
    rep("e4m3 Q,K per-block scales; bf16 P,V", attn(q8(q) @ q8(k).T, v, "bf16"))
    rep("e4m3 P,V; bf16 Q,K", attn(ref_s, q8(v), "e4m3"))
    rep("e4m3 Q,K,P,V", attn(q8(q) @ q8(k).T, q8(v), "e4m3"))
    km = k.mean(0, keepdim=True)
    rep("e4m3 Q, K - mean(K) (global smoothing)", attn(q8(q) @ q8(k - km).T, v, "bf16"))
    # two-sided per-block mean removal with exact rank-1 corrections:
    # q.k = (q - qb).(k - kb) [e4m3] + qb.k + q.kb - qb.kb [fp32]
    qb = q.view(nb, 128, 128).mean(1).repeat_interleave(128, 0)
    kb = k.view(nb, 128, 128).mean(1).repeat_interleave(128, 0)
    s2 = q8(q - qb) @ q8(k - kb).T + qb @ k.T + q @ kb.T - qb @ kb.T
    rep("e4m3 residuals + exact block-mean corrections", attn(s2, v, "bf16"))


if __name__ == "__main__":
    torch.manual_seed(0)
    for r in ("iid", "walk", "cluster"):
        run(r)
