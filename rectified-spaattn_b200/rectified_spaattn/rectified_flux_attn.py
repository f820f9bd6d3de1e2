"""Drop-in mirror of the reference rectified_spaattn/rectified_flux_attn.py (Flux.1-dev joint attention: image
tokens first, `text_length` text tokens last, no padding).  See rectified_wan21_attn.py for the symbol map."""
import torch

from rsa_b200 import geometry as _G
from rsa_b200 import ops as _ops

from . import _common
from .attn import fullattn  # noqa: F401
from .gapr_mask import estimate_pr_gain  # noqa: F401
from .rectified_wan21_attn import _triton_block_sparse_attention_onehot  # noqa: F401


def _build_block_index_with_importance_optimized(query, key, top_k, block_size_M=128, block_size_N=128,
                                                 text_start_block=None, text_end_block=None, num_blocks=None,
                                                 prob_threshold=0.7, block_neighbor_list=None, attenable=None):
    _common.check_blocks(block_size_M, block_size_N)
    geo = _G.flux(key.shape[2], int(attenable))
    return _common.build_index(query, key, geo, top_k, prob_threshold, block_neighbor_list)


def block_sparse_attention_combined(query, key, value, attn_mask, top_k, block_size_M=128, block_size_N=128,
                                    cu_seqlens_q=None, cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None,
                                    prob_threshold=0.5, block_neighbor_list=None, text_length=256, shape_xfuse=False):
    _common.check_blocks(block_size_M, block_size_N)
    cu = _common.host_ints(cu_seqlens_kv)
    kv_len = cu[1] if cu is not None else None     # seqlens = cu_seqlens_kv[1:2]  (reference :307)
    geo = _G.flux(query.shape[2], int(text_length), kv_len)
    return _common.run(query, key, value, geo, top_k, prob_threshold, block_neighbor_list, shape_xfuse)


def rectified_block_sparse_attention(query, key, value, attn_mask, top_k, block_size_M=128, block_size_N=128,
                                     cu_seqlens_q=None, cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None,
                                     block_neighbor_list=None, shape_xfuse=False, p_remain_rates=0.5, text_length=256):
    return block_sparse_attention_combined(
        query, key, value, attn_mask, top_k, block_size_M, block_size_N, cu_seqlens_q, cu_seqlens_kv,
        max_seqlen_q, max_seqlen_kv, block_neighbor_list=block_neighbor_list, shape_xfuse=shape_xfuse,
        prob_threshold=p_remain_rates, text_length=text_length)
