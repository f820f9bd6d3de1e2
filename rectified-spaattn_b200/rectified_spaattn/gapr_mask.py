"""Mirror of the reference rectified_spaattn/gapr_mask.py.  In this implementation the GAPR statistics are fused
into kernels 2 and 3a (csrc/pool_stats.cu, csrc/block_scores.cu); `estimate_pr_gain` is kept importable for
callers that address the stage on its own and runs those two kernels on a WAN-geometry plan."""
import torch

from rsa_b200 import geometry as _G
from rsa_b200 import ops as _ops


def estimate_pr_gain(Q_blocks, K_blocks, q_pools=None, k_pools=None, attention_scores=None):
    """Q_blocks, K_blocks: (B, H, NB, 128, D) bf16 CUDA -> ~gapr_mask (B, H, NQ, NK) bool (gapr_mask.py:4-42).
    The pooled arguments are accepted for signature compatibility; the kernels recompute them in fp32."""
    b, h, nq, bs, d = Q_blocks.shape
    if K_blocks.shape != Q_blocks.shape or bs != 128:
        raise ValueError("estimate_pr_gain expects equal Q/K block shapes with 128-token blocks")
    q = Q_blocks.reshape(b, h, nq * bs, d)
    k = K_blocks.reshape(b, h, nq * bs, d)
    plan = _ops.Plan(q, k, k, _G.wan(nq * bs), 1, 0.0)
    plan.pool_stats()
    plan.block_scores()
    return plan.view()["nogapr"].bool().view(b, h, nq, nq).clone()
