"""Shared host logic of the per-model entry points (not part of the reference's surface)."""
import torch

from rsa_b200 import geometry as G
from rsa_b200 import ops


def check_blocks(block_size_M, block_size_N):
    if block_size_M != 128 or block_size_N != 128:
        raise NotImplementedError("only block_size_M = block_size_N = 128 is built (every reference script uses 128)")


def host_ints(t):
    """cu_seqlens may arrive as a device tensor (reference processors build it with torch.tensor(..., device=...));
    reading it costs one host sync, which callers avoid by passing Python ints."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return [int(x) for x in t.tolist()]
    return [int(x) for x in t]


def run(query, key, value, geo, top_k, prob_threshold, block_neighbor_list, shape_xfuse, mask_cache=None):
    return ops.rectified_attention(query, key, value, geo, top_k, prob_threshold, block_neighbor_list, shape_xfuse,
                                   mask_cache=mask_cache)


def build_index(query, key, geo, top_k, prob_threshold, block_neighbor_list):
    """Stages 2, 3a, 3b on their own -> (one_hot [B,H,NQ,NB] bool, probs [B,H,NQ,n_ent] fp32, nogapr bool)."""
    b, h, s, d = key.shape
    if query.shape[2] != s:
        # the reference passes query[:, :, :normal_tokens]; the kernels index the visual rows themselves
        pad = torch.zeros(b, h, s - query.shape[2], d, dtype=query.dtype, device=query.device)
        query = torch.cat([query, pad], dim=2)
    plan = ops.Plan(query, key, key, geo, top_k, prob_threshold, block_neighbor_list, debug_dump_probs=True)
    plan.pool_stats()
    plan.block_scores()
    plan.block_select()
    vw = plan.view()
    nq, nb = geo.nq_blocks, geo.n_blocks
    one_hot = plan.dense_mask()[:, :nq].reshape(b, h, nq, nb).clone()
    probs = vw["probs"].reshape(b, h, nq, -1).clone()
    nogapr = vw["nogapr"].bool().reshape(b, h, nq, nq).clone()
    return one_hot, probs, nogapr
