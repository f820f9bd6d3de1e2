"""ORACLE support (test infrastructure): the reference's OWN hot-path code on CPU tensors -- bench.py's CPU arm.

`rectified_block_sparse_attention` of the unmodified reference (baseline/_ref, staged by oracle/stage_reference.py, or
/root/reference) runs on fp32 CPU tensors as it is: its PyTorch mask builder, GAPR test, IPAR, sort / cumsum / scatter,
R and C, the epilogue, cat / permute / reshape (rectified_hunyuan_attn.py:283-389, rectified_wan21_attn.py:276-357,
rectified_flux_attn.py:282-376, gapr_mask.py:4-42).  Two pieces cannot run without CUDA and are restated:
  * the Triton launch `_triton_block_sparse_attention_onehot` (wan21 :108-168; torch.cuda.device + triton.jit)
        -> oracle.rsa_oracle.masked_attention (dense masked softmax per query block; pinned against the literal kernel
           under TRITON_INTERPRET by tests/golden/kernel_fp16.npz)
  * `fullattn(..., "flash")` for the text rows (attn.py:60-120 -> flash_attn_varlen_func, CUDA only)
        -> oracle.rsa_oracle.dense_attention
fp32 instead of the scripts' bf16: bf16 matmuls on CPU are ~90x slower than fp32 (178 ms vs 2 ms at C1), so
fp32 is the FAVOURABLE setting for this baseline.

HunyuanVideo 129 frames (118 800 visual tokens, not a multiple of 128): the reference raises on it (hunyuan :356), so
the tensors are laid out [visual | zero rows up to the block boundary | text] first -- the rule its Wan / CogVideoX
paths apply to ragged inputs -- and the pad keys are skipped in the two restated pieces, exactly as in
oracle/make_golden.py's `hunyuan_ragged` fixture.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ref_loader
from . import rsa_oracle as O

MOD = {"wan": "rectified_wan21_attn", "hunyuan": "rectified_hunyuan_attn", "flux": "rectified_flux_attn",
       "cogvideo": "rectified_cogvideo_attn"}


def available():
    return ref_loader.reference_root() is not None


class ReferenceOnCpu:
    """One family's reference module with the two CUDA-only pieces restated; `call` = one attention call."""

    def __init__(self, fam):
        self.fam = fam
        self.ref = ref_loader.load([MOD[fam]])[MOD[fam]]
        self.hole = None
        self.stage_s = {}
        ref, me = self.ref, self

        def kernel(q, k, v, seqlens, block_mask, sm_scale, bm=128, bn=128):
            import time
            t0 = time.perf_counter()
            b, h, s, d = q.shape
            out = torch.zeros_like(q)
            for bi in range(b):
                for hi in range(h):
                    o = O.masked_attention(q[bi, hi].float().numpy(), k[bi, hi].float().numpy(),
                                           v[bi, hi].float().numpy(), block_mask[bi, hi].numpy(), int(seqlens[bi]), s,
                                           hole=me.hole)
                    out[bi, hi] = torch.from_numpy(o).to(q.dtype)
            me.stage_s["kernel"] = me.stage_s.get("kernel", 0.0) + time.perf_counter() - t0
            return out

        def fullattn_cpu(q_, k_, v_, mode, drop_rate=0, attn_mask=None, causal=False, cu_seqlens_q=None,
                         cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None, batch_size=1):
            # varlen semantics of attn.py:107-120: sequence i = query rows [cu_q[i], cu_q[i+1]) x keys [cu_k[i], cu_k[i+1])
            b, h, sq, d = q_.shape
            out = torch.zeros_like(q_)
            for i in range(len(cu_seqlens_q) - 1):
                q0, q1 = int(cu_seqlens_q[i]), int(cu_seqlens_q[i + 1])
                k0, k1 = int(cu_seqlens_kv[i]), int(cu_seqlens_kv[i + 1])
                if q1 <= q0 or i > 0:      # the second "sequence" is the padding attending itself: discarded by callers
                    continue
                for hi in range(h):
                    o = O.dense_attention(q_[0, hi, q0:q1].float().numpy(), k_[0, hi, k0:k1].float().numpy(),
                                          v_[0, hi, k0:k1].float().numpy(), k1 - k0, me.hole)
                    out[0, hi, q0:q1] = torch.from_numpy(o).to(q_.dtype)
            return out

        ref._triton_block_sparse_attention_onehot = kernel
        ref.fullattn = fullattn_cpu

    def call(self, q, k, v, nbr, top_k, p_remain, num_true=None, text_len=None, first_frame_blocks=None):
        """q, k, v: fp32 CPU tensors [1, H, S, D] as the caller holds them -> [1, S, H*D]."""
        ref, fam = self.ref, self.fam
        s = q.shape[2]
        self.hole = None
        if fam == "wan":
            return ref.rectified_block_sparse_attention(q, k, v, None, top_k, block_neighbor_list=nbr,
                                                        p_remain_rates=p_remain, first_frame_blocks=first_frame_blocks)
        if fam == "hunyuan":
            nv = s - 256
            gap = (-nv) % 128
            if gap:                      # the 129-frame shape: run the reference on the explicitly padded layout
                pad = lambda x: torch.cat([x[:, :, :nv], torch.zeros_like(x[:, :, :gap]), x[:, :, nv:]], dim=2)
                q, k, v = pad(q), pad(k), pad(v)
                self.hole = (nv, nv + gap)
            sp, nt = s + gap, num_true + gap
            am = torch.zeros(1, 1, 1, sp, dtype=torch.bool)
            am[..., :nt] = True
            cu = torch.tensor([0, nt, sp], dtype=torch.int32)
            out = ref.rectified_block_sparse_attention(q, k.clone(), v.clone(), am, top_k, cu_seqlens_q=cu,
                                                       cu_seqlens_kv=cu, max_seqlen_q=sp, max_seqlen_kv=sp,
                                                       block_neighbor_list=nbr, p_remain_rates=p_remain)
            if gap:
                out = torch.cat([out[:, :nv], out[:, nv + gap:]], dim=1)
            return out
        cu = torch.tensor([0, s, s], dtype=torch.int32)
        return ref.rectified_block_sparse_attention(q, k, v, None, top_k, cu_seqlens_q=cu, cu_seqlens_kv=cu,
                                                    max_seqlen_q=s, max_seqlen_kv=s, block_neighbor_list=nbr,
                                                    p_remain_rates=p_remain, text_length=text_len)
