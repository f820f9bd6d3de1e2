"""Drop-in mirror of the reference rectified_spaattn/rectified_hunyuan_attn.py (HunyuanVideo joint text+video
attention: video tokens first, 256 text tokens last).  See rectified_wan21_attn.py for the symbol map; the
text-row flash-attn call (reference :371-380) and the K/V masked_fill_ (:307-308) are folded into the kernels
(dense query tiles, kv_len bound) -- key/value are NOT modified in place.
"""
import torch

from rsa_b200 import geometry as _G
from rsa_b200 import ops as _ops

from . import _common
from .attn import fullattn  # noqa: F401
from .gapr_mask import estimate_pr_gain  # noqa: F401
from .rectified_wan21_attn import _triton_block_sparse_attention_onehot  # noqa: F401  (identical kernel)


def _geometry(seq, cu_seqlens_q, num_true):
    if num_true is None:
        cu = _common.host_ints(cu_seqlens_q)
        if cu is None:
            raise RuntimeError("HunyuanVideo path needs cu_seqlens_q = [0, num_true, S] (the reference's "
                               "no-cu_seqlens branch leaves `attenable` undefined, rectified_hunyuan_attn.py:318-323)")
        num_true = cu[1]
        if cu[2] != seq:
            raise ValueError("cu_seqlens_q[2] must equal the sequence length")
    return _G.hunyuan(seq, int(num_true))


def _build_block_index_with_importance_optimized(query, key, top_k, block_size_M=128, block_size_N=128,
                                                 text_start_block=None, text_end_block=None, num_blocks=None,
                                                 prob_threshold=0.7, block_neighbor_list=None, attenable=None):
    _common.check_blocks(block_size_M, block_size_N)
    s = key.shape[2]
    a = int(attenable)
    geo = _G.hunyuan(s, s - 256 + a)
    return _common.build_index(query, key, geo, top_k, prob_threshold, block_neighbor_list)


def block_sparse_attention_combined(query, key, value, attn_mask, top_k, block_size_M=128, block_size_N=128,
                                    cu_seqlens_q=None, cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None,
                                    prob_threshold=0.5, block_neighbor_list=None, shape_xfuse=False, num_true=None):
    _common.check_blocks(block_size_M, block_size_N)
    geo = _geometry(query.shape[2], cu_seqlens_q, num_true)
    return _common.run(query, key, value, geo, top_k, prob_threshold, block_neighbor_list, shape_xfuse)


def rectified_block_sparse_attention(query, key, value, attn_mask, top_k, block_size_M=128, block_size_N=128,
                                     cu_seqlens_q=None, cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None,
                                     block_neighbor_list=None, shape_xfuse=False, p_remain_rates=0.5, num_true=None):
    """`num_true` (host int) is an optional extension: it avoids reading cu_seqlens_q back from the device."""
    return block_sparse_attention_combined(
        query, key, value, attn_mask, top_k, block_size_M, block_size_N, cu_seqlens_q, cu_seqlens_kv,
        max_seqlen_q, max_seqlen_kv, block_neighbor_list=block_neighbor_list, shape_xfuse=shape_xfuse,
        prob_threshold=p_remain_rates, num_true=num_true)
