"""Drop-in mirror of the reference rectified_spaattn/rectified_wan21_attn.py (Wan2.1 / Wan2.2 self-attention over
video tokens).  Same public names and signatures; the PyTorch-eager mask builder, the Triton kernel and the
elementwise epilogue of the reference are replaced by the sm_100a kernels behind librsa_b200.so.

reference symbol                                    here
_triton_block_sparse_attention_onehot  (:108-168)   kernel 4 on a dense block mask (rsa_masked_attention)
_build_block_index_with_importance_optimized (:171) kernels 2 + 3a + 3b
block_sparse_attention_combined        (:276-357)   rsa_rectified_attention (kernels 2, 3a, 3b, 3c, 4)
rectified_block_sparse_attention       (:361-386)   same wrapper (p_remain_rates -> prob_threshold)
"""
from typing import Optional

import torch

from rsa_b200 import geometry as _G
from rsa_b200 import ops as _ops

from . import _common
from .attn import fullattn  # noqa: F401  (re-exported like the reference)
from .gapr_mask import estimate_pr_gain  # noqa: F401


def _triton_block_sparse_attention_onehot(q, k, v, seqlens, block_mask, sm_scale, block_size_M=128, block_size_N=128):
    """q,k,v [B,H,S,D] bf16; seqlens [B] valid KV length; block_mask [B,H,NQ,NB] bool -> [B,H,S,D].
    (name kept from the reference; the kernel is tcgen05/TMEM/TMA CUDA, not Triton)"""
    assert q.shape[-1] == k.shape[-1] == v.shape[-1]
    # the reference asserts Lk in {16, 32, 64, 128} (:121); 128 (every BASELINE model) and 64 (CogVideoX) are built
    assert k.shape[-1] in {64, 128}, "head_dim must be 128 or 64"
    _common.check_blocks(block_size_M, block_size_N)
    lens = _common.host_ints(seqlens)
    if len(lens) not in (1, q.shape[0]):
        raise ValueError("seqlens must hold one valid KV length per batch element")
    if len(set(lens)) == 1:
        return _ops.masked_attention(q, k, v, block_mask, lens[0], sm_scale)
    # the reference kernel loads seqlens[off_hz // H] per batch element (:36-37, :86): one launch per element, each
    # writing its slice of the result
    out = torch.empty_like(q)
    for b, n in enumerate(lens):
        _ops.masked_attention(q[b:b + 1], k[b:b + 1], v[b:b + 1], block_mask[b:b + 1], n, sm_scale, out=out[b:b + 1])
    return out


def _build_block_index_with_importance_optimized(query, key, top_k, block_size_M=128, block_size_N=128,
                                                 text_start_block=None, text_end_block=None, num_blocks=None,
                                                 prob_threshold=0.7, block_neighbor_list=None,
                                                 first_frame_blocks=None):
    _common.check_blocks(block_size_M, block_size_N)
    s = key.shape[2]
    if s % 128:
        raise RuntimeError("inputs must be padded to a multiple of 128 tokens (the reference reshapes to blocks)")
    geo = _G.wan(s, first_frame_blocks or 0)
    return _common.build_index(query, key, geo, top_k, prob_threshold, block_neighbor_list)


def block_sparse_attention_combined(query, key, value, attn_mask, top_k, block_size_M=128, block_size_N=128,
                                    cu_seqlens_q=None, cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None,
                                    prob_threshold=0.5, block_neighbor_list=None, shape_xfuse=False,
                                    first_frame_blocks=None,
                                    mask_cache=None):
    """[B,H,S,D] -> [B,S,H*D] ([B,S,H,D] if shape_xfuse).  S need not be a multiple of 128: the kernels treat
    the ragged tail as the zero rows the reference pads with (:299-302)."""
    _common.check_blocks(block_size_M, block_size_N)
    geo = _G.wan(query.shape[2], first_frame_blocks or 0)
    return _common.run(query, key, value, geo, top_k, prob_threshold, block_neighbor_list, shape_xfuse,
                       mask_cache)


def rectified_block_sparse_attention(query, key, value, attn_mask, top_k, block_size_M=128, block_size_N=128,
                                     cu_seqlens_q=None, cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None,
                                     block_neighbor_list=None, shape_xfuse=False, p_remain_rates=0.5,
                                     first_frame_blocks=None,
                                     mask_cache=None):
    return block_sparse_attention_combined(
        query, key, value, attn_mask, top_k, block_size_M, block_size_N, cu_seqlens_q, cu_seqlens_kv,
        max_seqlen_q, max_seqlen_kv, block_neighbor_list=block_neighbor_list, shape_xfuse=shape_xfuse,
        prob_threshold=p_remain_rates, first_frame_blocks=first_frame_blocks, mask_cache=mask_cache)


from . import _processors as _P  # noqa: E402


class RectifiedWanT2VSpaAttnProcessor2_0(_P.WanProcessorBase):
    """Wan2.1 T2V self-attention processor (reference :389-509): sparse from layer 2 on and from call 10 on
    (5 denoising steps x cond/uncond), dense ("flash") before; `current_step` wraps at 100."""

    def sparse_now(self):
        return self.processor_id >= 2 and self.current_step >= 10


class RectifiedWanI2VSpaAttnProcessor2_0(_P.WanProcessorBase):
    """Wan2.1 I2V (reference :512-632): sparse from layer 2 on at every step; adds the CLIP image cross-attention."""

    def sparse_now(self):
        return self.processor_id >= 2
