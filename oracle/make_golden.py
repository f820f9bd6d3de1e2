"""ORACLE support (test infrastructure): generate tests/golden/*.npz|json by running the UNMODIFIED reference
(/root/reference) on CPU in the build container.  Re-run with `python -m oracle.make_golden`.

What is recorded (inputs are regenerated from seeds by oracle.rsa_oracle.synth_qkv, so only outputs are stored):
  gilbert.json        linear_to_hilbert / hilbert_to_linear / neighbour matrices of utils/jenga_gilbert.py
                      (full arrays for small grids, SHA-1 digests for the BASELINE.json grids)
  mask_<case>.npz     outputs of each family's _build_block_index_with_importance_optimized run on fp32 CPU
                      tensors (one_hot mask, probs, nogapr) captured from inside block_sparse_attention_combined,
                      the rectification factors R and C recovered from the same call, and the visual-row output
  kernel_fp16.npz     the literal Triton kernel (_triton_block_sparse_attention_onehot) under TRITON_INTERPRET=1
                      in fp16 on a random block mask (the interpreter has no bf16)

Ragged HunyuanVideo case (hunyuan_ragged): the reference raises on it (:356), so it is run on the explicitly padded
layout [visual | zero rows up to the block boundary | text] with num_true shifted by the pad; mask, probs, nogapr, R
and C then come from the unmodified reference code, and the two attention stand-ins below additionally skip the pad
keys (the extension's rule, geometry_hunyuan).  The stored output has the pad rows removed.

Stand-ins for third-party pieces that cannot run on CPU (named so the goldens stay honest):
  * flash_attn_varlen_func (flash-attn, unpinned in the reference README; 2.8.3 here) -> per-sequence dense softmax
  * the Triton kernel inside the *combined* call -> dense masked softmax (its literal semantics are pinned
    separately by kernel_fp16.npz)
"""
from __future__ import annotations

import os

os.environ.setdefault("TRITON_INTERPRET", "1")

import contextlib
import hashlib
import json
import sys

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _REPO)

from oracle import ref_loader  # noqa: E402
from oracle import rsa_oracle as O  # noqa: E402

GOLD = os.path.join(_REPO, "tests", "golden")


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ------------------------------------------------------------------------------------------------ gilbert
SMALL_GRIDS = [(4, 16, 16), (3, 5, 7), (2, 6, 10), (1, 16, 16), (5, 6, 4), (2, 3, 130), (1, 1, 9)]
BIG_GRIDS = [(21, 30, 52), (21, 45, 80), (32, 45, 80), (33, 45, 80), (1, 256, 256), (11, 48, 80)]


def make_gilbert(big=True):
    g = ref_loader.load(["jenga_gilbert"])["jenga_gilbert"]
    out = {"small": [], "big": []}
    for (t, h, w) in SMALL_GRIDS:
        with ref_loader.quiet():
            l2h, h2l = g.gilbert_mapping(t, h, w)
            nbr = g.gilbert_block_neighbor_mapping(t, h, w)
        out["small"].append(dict(grid=[t, h, w], l2h=list(map(int, l2h)), h2l=list(map(int, h2l)),
                                 nbr=np.packbits(nbr.numpy().astype(np.uint8)).tolist(),
                                 nbr_shape=list(nbr.shape)))
        # alternative axis order exercised by the same code path
        with ref_loader.quiet():
            l2h2, h2l2 = g.gilbert_mapping(t, h, w, axis_order=("t", "h", "w"))
        out["small"][-1]["h2l_thw"] = list(map(int, h2l2))
    if big:
        for (t, h, w) in BIG_GRIDS:
            with ref_loader.quiet():
                l2h, h2l = g.gilbert_mapping(t, h, w)
                nbr = g.gilbert_block_neighbor_mapping(t, h, w)
            out["big"].append(dict(grid=[t, h, w], sha_h2l=sha(np.asarray(h2l, dtype=np.int64)),
                                   sha_l2h=sha(np.asarray(l2h, dtype=np.int64)),
                                   sha_nbr=sha(nbr.numpy().astype(np.uint8)), nbr_nnz=int(nbr.sum()),
                                   h2l_head=list(map(int, h2l[:6]))))
            print("gilbert", (t, h, w), out["big"][-1]["sha_h2l"], flush=True)
    with open(os.path.join(GOLD, "gilbert.json"), "w") as f:
        json.dump(out, f)


# ------------------------------------------------------------------------------------- mask builder + R, C
HOLE = None  # (first, end) pad rows inside the sequence that no query may attend; set for the ragged case only


def _varlen_dense(q, k, v, cu_q, cu_k):
    """Stand-in for flash_attn_varlen_func on [(tokens), H, D] tensors."""
    out = torch.zeros_like(q)
    for i in range(len(cu_q) - 1):
        q0, q1, k0, k1 = int(cu_q[i]), int(cu_q[i + 1]), int(cu_k[i]), int(cu_k[i + 1])
        if q1 <= q0:
            continue
        keys = torch.arange(k0, k1)
        if HOLE is not None:
            keys = keys[(keys < HOLE[0]) | (keys >= HOLE[1])]
        qi = q[q0:q1].transpose(0, 1).float()
        ki = k[keys].transpose(0, 1).float()
        vi = v[keys].transpose(0, 1).float()
        s = (qi @ ki.transpose(1, 2)) * (q.shape[-1] ** -0.5)
        out[q0:q1] = (torch.softmax(s, -1) @ vi).transpose(0, 1).to(q.dtype)
    return out


def _dense_masked_kernel(q, k, v, seqlens, block_mask, sm_scale, bm=128, bn=128):
    """Stand-in for the Triton launch inside the combined call (dense masked softmax, fp32)."""
    b, h, s, d = q.shape
    # the stand-in computes with d^-1/2 (rsa_oracle.masked_attention); hold the reference to passing exactly that
    assert abs(float(sm_scale) - d ** -0.5) < 1e-12, (sm_scale, d)
    out = torch.zeros_like(q)
    for bi in range(b):
        for hi in range(h):
            o = O.masked_attention(q[bi, hi].float().numpy(), k[bi, hi].float().numpy(), v[bi, hi].float().numpy(),
                                   block_mask[bi, hi].numpy(), int(seqlens[bi]), s, hole=HOLE)
            out[bi, hi] = torch.from_numpy(o).to(q.dtype)
    return out


from oracle.cases import CASES, EXTRA_CASES, MID_CASES, case_inputs, mid_case_inputs  # noqa: E402

MID_PROB_ROWS = 48   # rows of the probability matrix a mid-size fixture keeps (evenly spaced; the full matrix is MBs)


def run_reference_case(name):
    """CASES / EXTRA_CASES: everything, including the visual-row output.  MID_CASES (`mid`): the mask-building stages
    only -- the attention stand-in is not run (it would take minutes at 140 000 tokens and is size-independent), the
    probabilities are kept for MID_PROB_ROWS rows plus every row's sum."""
    mid = name in MID_CASES
    fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v = (mid_case_inputs if mid else case_inputs)(name)
    modname = {"wan": "rectified_wan21_attn", "hunyuan": "rectified_hunyuan_attn", "flux": "rectified_flux_attn",
               "cogvideo": "rectified_cogvideo_attn"}[fam]
    mods = ref_loader.load([modname, "jenga_gilbert"])
    ref, g = mods[modname], mods["jenga_gilbert"]
    with ref_loader.quiet():
        nbr = g.gilbert_block_neighbor_mapping(t, h, w)
    cap = {}
    orig_build = ref._build_block_index_with_importance_optimized

    def build_spy(*a, **kw):
        r = orig_build(*a, **kw)
        cap["mask"], cap["probs"], cap["nogapr"] = [x.clone() for x in r]
        return r

    def fullattn_cpu(q_, k_, v_, mode, drop_rate=0, attn_mask=None, causal=False, cu_seqlens_q=None,
                     cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None, batch_size=1):
        pre = lambda x: x.transpose(1, 2).reshape(x.shape[0] * x.shape[2], x.shape[1], x.shape[3])
        x = _varlen_dense(pre(q_), pre(k_), pre(v_), cu_seqlens_q, cu_seqlens_kv)
        x = x.view(batch_size, max_seqlen_q, x.shape[-2], x.shape[-1])
        return x.transpose(1, 2)

    ref._build_block_index_with_importance_optimized = build_spy
    ref.fullattn = fullattn_cpu
    global HOLE
    HOLE = None
    s_in = s
    gap = (-nv) % 128 if fam == "hunyuan" else 0
    if gap:
        q, k, v = (np.concatenate([x[:, :, :nv], np.zeros_like(x[:, :, :gap]), x[:, :, nv:]], axis=2) for x in (q, k, v))
        HOLE = (nv, nv + gap)
        s = s + gap
        ntrue_d = ntrue_d + gap
    tq, tk, tv = (torch.from_numpy(x) for x in (q, k, v))
    ffb = (nv + 127) // 128 // t if fam == "wan" else None
    num_true = nv + ntrue_d

    def call(kernel):
        ref._triton_block_sparse_attention_onehot = kernel
        kq, kk, kv_ = tq.clone(), tk.clone(), tv.clone()
        if fam == "wan":
            return ref.rectified_block_sparse_attention(kq, kk, kv_, None, top_k, block_neighbor_list=nbr,
                                                        p_remain_rates=p, first_frame_blocks=ffb)
        if fam == "hunyuan":
            am = torch.zeros(1, 1, 1, s, dtype=torch.bool)
            am[..., :num_true] = True
            cu = torch.tensor([0, num_true, s], dtype=torch.int32)
            return ref.rectified_block_sparse_attention(kq, kk, kv_, am, top_k, cu_seqlens_q=cu, cu_seqlens_kv=cu,
                                                        max_seqlen_q=s, max_seqlen_kv=s, block_neighbor_list=nbr,
                                                        p_remain_rates=p)
        if fam == "flux":
            cuq = torch.tensor([0, s, s], dtype=torch.int32)
            return ref.rectified_block_sparse_attention(kq, kk, kv_, None, top_k, cu_seqlens_q=cuq, cu_seqlens_kv=cuq,
                                                        max_seqlen_q=s, max_seqlen_kv=s, block_neighbor_list=nbr,
                                                        p_remain_rates=p, text_length=text_len)
        cuq = torch.tensor([0, s, s], dtype=torch.int32)
        return ref.rectified_block_sparse_attention(kq, kk, kv_, None, top_k, cu_seqlens_q=cuq, cu_seqlens_kv=cuq,
                                                    max_seqlen_q=s, max_seqlen_kv=s, block_neighbor_list=nbr,
                                                    p_remain_rates=p, text_length=text_len)

    zeros = lambda q_, *a, **kw: torch.zeros_like(q_)
    ones = lambda q_, *a, **kw: torch.ones_like(q_)
    out_c = call(zeros)          # visual rows hold C (block-constant)
    out_rc = call(ones)          # visual rows hold R + C
    out = out_c if mid else call(_dense_masked_kernel)
    nq = cap["mask"].shape[2]
    rows_vis = min(nq * 128, s)
    hd = q.shape[-1]                                               # 128, or 64 for cog_d64
    oc = out_c.reshape(1, s, heads, hd)[0].permute(1, 0, 2)       # [H, S, D]
    orc = out_rc.reshape(1, s, heads, hd)[0].permute(1, 0, 2)
    idx = torch.arange(0, rows_vis, 128)
    c = oc[:, idx]                                                 # [H, NQ, D]
    r = (orc[:, idx] - c).mean(dim=-1)                             # [H, NQ]  (R + C - C, 128 identical columns)
    out_rows = out[0]
    if gap:
        out_rows = torch.cat([out_rows[:nv], out_rows[nv + gap:]], dim=0)
        assert out_rows.shape[0] == s_in
    HOLE = None
    if mid:
        pr = cap["probs"][0].numpy().astype(np.float32)             # [H, NQ, n_ent]
        rows = np.unique(np.linspace(0, nq - 1, MID_PROB_ROWS).astype(np.int64))
        np.savez_compressed(
            os.path.join(GOLD, f"mask_{name}.npz"),
            mask=np.packbits(cap["mask"][0].numpy().astype(np.uint8)), mask_shape=np.array(cap["mask"][0].shape),
            prob_rows=rows, probs=pr[:, rows], prob_sum=pr.sum(axis=2, dtype=np.float64).astype(np.float32),
            nogapr=np.packbits(cap["nogapr"][0].numpy().astype(np.uint8)),
            nogapr_shape=np.array(cap["nogapr"][0].shape),
            R=r.numpy().astype(np.float32), C=c.numpy().astype(np.float32),
            nbr=np.packbits(nbr.numpy().astype(np.uint8)), nbr_shape=np.array(nbr.shape))
        print("mid case", name, "NQ", nq, "mask density", float(cap["mask"].float().mean()), "nogapr",
              float(cap["nogapr"].float().mean()), "R min/mean", float(r.min()), float(r.mean()), flush=True)
        return
    np.savez_compressed(
        os.path.join(GOLD, f"mask_{name}.npz"),
        mask=np.packbits(cap["mask"][0].numpy().astype(np.uint8)), mask_shape=np.array(cap["mask"][0].shape),
        probs=cap["probs"][0].numpy().astype(np.float32),
        nogapr=np.packbits(cap["nogapr"][0].numpy().astype(np.uint8)),
        nogapr_shape=np.array(cap["nogapr"][0].shape),
        R=r.numpy().astype(np.float32), C=c.numpy().astype(np.float32),
        out=out_rows.numpy().astype(np.float32),
        nbr=np.packbits(nbr.numpy().astype(np.uint8)), nbr_shape=np.array(nbr.shape))
    print("case", name, "mask density", float(cap["mask"].float().mean()), "nogapr", float(cap["nogapr"].float().mean()),
          "R min/mean", float(r.min()), float(r.mean()), flush=True)


# ------------------------------------------------------------------------------- literal Triton kernel, fp16
def make_kernel_fp16():
    class _NullDev(contextlib.nullcontext):
        def __init__(self, *_a, **_k):
            super().__init__()

    ref = ref_loader.load(["rectified_wan21_attn"])["rectified_wan21_attn"]
    torch.manual_seed(11)
    h, s_valid, s_pad = 2, 1000, 1024
    q = torch.randn(1, h, s_pad, 128).half()
    k = torch.randn(1, h, s_pad, 128).half()
    v = torch.randn(1, h, s_pad, 128).half()
    q[:, :, s_valid:] = 0
    k[:, :, s_valid:] = 0
    v[:, :, s_valid:] = 0
    mask = torch.rand(1, h, 8, 8) < 0.4
    mask |= torch.eye(8, dtype=torch.bool)
    seqlens = torch.tensor([s_valid], dtype=torch.int32)
    saved = torch.cuda.device
    torch.cuda.device = _NullDev
    try:
        o = ref._triton_block_sparse_attention_onehot(q, k, v, seqlens, mask, 128 ** -0.5, 128, 128)
    finally:
        torch.cuda.device = saved
    np.savez_compressed(os.path.join(GOLD, "kernel_fp16.npz"), q=q.numpy(), k=k.numpy(), v=v.numpy(),
                        mask=mask.numpy(), seqlen=np.array(s_valid), out=o.numpy())
    print("kernel_fp16 done", float(o.float().abs().mean()), flush=True)


# ------------------------------------------------------------------ pre-attention sequence (kernel 0), bf16 CPU
def prep_inputs(seed=31, rows=390, heads=2, n_rope=300):
    """Seeded inputs shared with the tests: projection outputs, norm weights, rotary tables."""
    g = torch.Generator().manual_seed(seed)
    src = [torch.randn(1, rows, heads * 128, generator=g).mul(1.5).to(torch.bfloat16) for _ in range(3)]
    wq = (1 + 0.1 * torch.randn(128, generator=g)).to(torch.bfloat16)
    wk = (1 + 0.1 * torch.randn(128, generator=g)).to(torch.bfloat16)
    pos = torch.arange(n_rope, dtype=torch.float32)
    inv = 1.0 / (256.0 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
    ang = torch.outer(pos, inv)
    cos, sin = ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)     # [n_rope, 128]
    return src, wq, wk, cos, sin, n_rope


def torch_prep(x, heads, weight, eps, cos, sin, n_rope):
    """The literal op sequence of rectified_hunyuan_attn.py:448-479 with diffusers' RMSNorm.forward and
    apply_rotary_emb(use_real=True, use_real_unbind_dim=-1) written out (diffusers is not installed)."""
    x = x.unflatten(2, (heads, -1)).transpose(1, 2)
    if weight is not None:
        variance = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(variance + eps)
        x = x.to(weight.dtype) * weight
    if n_rope:
        xr = x[:, :, :n_rope]
        c, s_ = cos[None, None], sin[None, None]
        x_real, x_imag = xr.reshape(*xr.shape[:-1], -1, 2).unbind(-1)
        x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
        xr = (xr.float() * c + x_rot.float() * s_).to(x.dtype)
        x = torch.cat([xr, x[:, :, n_rope:]], dim=2)
    return x


def torch_prep_wan(x, heads, weight, eps, freqs):
    """The literal op sequence of rectified_wan21_attn.py:419-441: RMSNorm over the inner dim (diffusers RMSNorm.forward
    written out), head split, complex rotary embedding in float64.  freqs: complex [1, 1, S, 64]."""
    if weight is not None:
        variance = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(variance + eps)
        x = x.to(weight.dtype) * weight
    x = x.unflatten(2, (heads, -1)).transpose(1, 2)
    if freqs is not None:
        xr = torch.view_as_complex(x.to(torch.float64).unflatten(3, (-1, 2)))
        x = torch.view_as_real(xr * freqs).flatten(3, 4).type_as(x)
    return x


def torch_prep_cog(x, heads, weight, bias, eps, cos, sin, n_rope):
    """rectified_cogvideo_attn.py:443-469: head split, torch.nn.LayerNorm(head_dim), rotary embedding on the video tokens."""
    x = x.unflatten(2, (heads, -1)).transpose(1, 2)
    x = torch.nn.functional.layer_norm(x, (x.shape[-1],), weight, bias, eps)
    return torch_prep(x.transpose(1, 2).flatten(2), heads, None, eps, cos, sin, n_rope)


def cog_prep_params():
    g = torch.Generator().manual_seed(33)
    w = [(1 + 0.1 * torch.randn(128, generator=g)).to(torch.bfloat16) for _ in range(2)]
    b = [(0.05 * torch.randn(128, generator=g)).to(torch.bfloat16) for _ in range(2)]
    return w, b


def make_prep():
    src, wq, wk, cos, sin, n_rope = prep_inputs()
    cw, cb = cog_prep_params()
    qc = torch_prep_cog(src[0], 2, cw[0], cb[0], 1e-6, cos, sin, n_rope)
    kc = torch_prep_cog(src[1], 2, cw[1], cb[1], 1e-6, cos, sin, n_rope)
    np.savez_compressed(os.path.join(GOLD, "prep_cog.npz"), q=qc.contiguous().view(torch.int16).numpy(),
                        k=kc.contiguous().view(torch.int16).numpy())
    # Wan form on the same sources: weights over the inner dim, rotary embedding on every token (complex, float64)
    g = torch.Generator().manual_seed(32)
    rows = src[0].shape[1]
    wq_in = (1 + 0.1 * torch.randn(256, generator=g)).to(torch.bfloat16)
    wk_in = (1 + 0.1 * torch.randn(256, generator=g)).to(torch.bfloat16)
    ang = torch.outer(torch.arange(rows, dtype=torch.float64), 1.0 / (256.0 ** (torch.arange(0, 128, 2, dtype=torch.float64) / 128)))
    freqs = torch.polar(torch.ones_like(ang), ang)[None, None]
    qw = torch_prep_wan(src[0], 2, wq_in, 1e-6, freqs)
    kw = torch_prep_wan(src[1], 2, wk_in, 1e-6, freqs)
    np.savez_compressed(os.path.join(GOLD, "prep_wan.npz"), q=qw.contiguous().view(torch.int16).numpy(),
                        k=kw.contiguous().view(torch.int16).numpy())
    q = torch_prep(src[0], 2, wq, 1e-6, cos, sin, n_rope)
    k = torch_prep(src[1], 2, wk, 1e-6, cos, sin, n_rope)
    v = torch_prep(src[2], 2, None, 1e-6, None, None, 0)
    np.savez_compressed(os.path.join(GOLD, "prep_hunyuan.npz"), q=q.contiguous().view(torch.int16).numpy(),
                        k=k.contiguous().view(torch.int16).numpy(), v=v.contiguous().view(torch.int16).numpy())
    print("prep golden", tuple(q.shape), float(q.float().abs().mean()), flush=True)


# ------------------------------------------------------------------- recursion helpers and the public API surface
def make_helpers():
    """gilbert_helpers.json: sgn / in_bounds / gilbert_xyz2d_r of the reference on seeded random frames (the three
    frame forms of gilbert_xyz2d :34-54 plus the fourth permutation, shifted origins, index offsets, flipped major axis).
    api_signatures.json: names, parameter names, kinds and defaults of every function and class (__init__, __call__)
    the reference's hot-path modules define -- what "drop-in" means for the mirror modules of this repo."""
    import inspect
    import random

    g = ref_loader.load(["jenga_gilbert"])["jenga_gilbert"]
    rnd = random.Random(5)
    rows = []
    while len(rows) < 200:
        w, h, d = (rnd.randint(1, 9) for _ in range(3))
        a, b, c = rnd.choice([((w, 0, 0), (0, h, 0), (0, 0, d)), ((0, h, 0), (w, 0, 0), (0, 0, d)),
                              ((0, 0, d), (w, 0, 0), (0, h, 0)), ((0, 0, d), (0, h, 0), (w, 0, 0))])
        o = [rnd.randint(-5, 5) for _ in range(3)]
        pt = (o[0] + rnd.randrange(w), o[1] + rnd.randrange(h), o[2] + rnd.randrange(d))
        ci = rnd.randint(0, 1000)
        if rnd.random() < 0.3:
            ax = [i for i in range(3) if a[i] != 0][0]
            o[ax] += a[ax] - 1
            a = tuple(-v for v in a)
        outside = (pt[0] + rnd.choice((-11, 11)), pt[1], pt[2])
        frame = [*o, *a, *b, *c]
        rows.append({"args": [ci, *pt, *frame], "index": g.gilbert_xyz2d_r(ci, *pt, *frame),
                     "in_bounds": bool(g.in_bounds(*pt, *frame)), "outside": list(outside),
                     "outside_in_bounds": bool(g.in_bounds(*outside, *frame))})
    with open(os.path.join(GOLD, "gilbert_helpers.json"), "w") as f:
        json.dump({"sgn": {str(v): g.sgn(v) for v in (-7, -1, 0, 1, 12)}, "cases": rows}, f)

    mods = ["rectified_wan21_attn", "rectified_wan22_attn", "rectified_hunyuan_attn", "rectified_flux_attn",
            "rectified_cogvideo_attn", "attn", "gapr_mask", "attn_processor", "jenga_gilbert"]
    ref = ref_loader.load(mods)

    def sig(fn):
        return [[p.name, p.kind.name, None if p.default is inspect._empty else repr(p.default)]
                for p in inspect.signature(fn).parameters.values()]

    api = {}
    for name, m in ref.items():
        entry = {}
        for k, v in vars(m).items():
            if k.startswith("__") or getattr(v, "__module__", None) != m.__name__:
                continue
            if inspect.isfunction(v):
                entry[k] = {"kind": "function", "params": sig(v)}
            elif inspect.isclass(v):
                entry[k] = {"kind": "class", "init": sig(v.__init__)}
                if "__call__" in vars(v):
                    entry[k]["call"] = sig(v.__call__)
        api[name] = entry
    with open(os.path.join(GOLD, "api_signatures.json"), "w") as f:
        json.dump(api, f, indent=1)
    print("helpers", len(rows), "api", {k: len(v) for k, v in api.items()}, flush=True)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1:] or ["gilbert", "masks", "mid", "kernel", "prep", "helpers"]
    if "gilbert" in what:
        make_gilbert()
    if "masks" in what:
        every = list(CASES) + list(EXTRA_CASES)
        for n in every:
            if n in what or not any(w in every for w in what):   # `masks <case> ...` regenerates only those
                run_reference_case(n)
    if "mid" in what:
        for n in MID_CASES:
            if n in what or not any(w in MID_CASES for w in what):
                run_reference_case(n)
    if "kernel" in what:
        make_kernel_fp16()
    if "prep" in what:
        make_prep()
    if "helpers" in what:
        make_helpers()
