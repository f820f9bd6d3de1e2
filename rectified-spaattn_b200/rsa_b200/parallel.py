"""Multi-GPU host logic for the rectified sparse-attention path: one process per GPU (torch.distributed, NCCL over
NVLink 5 / NVSwitch on the B200 box; gloo in the CPU tests).

The path is independent per (batch, head) end to end -- pooling, scoring, selection, attention and rectification
all run per head (reference grid axis B*H, rectified_wan21_attn.py:138) -- so the natural sharding is
HEAD-PARALLEL with no collective on the data path: every rank runs the same call on its own contiguous slice of
heads.  The reference itself has no multi-GPU path for one attention call (it only fans prompts out over GPUs,
eval/video/experiments/multigpu_hunyuan.py:287-298); its `shape_xfuse` flag (rectified_wan21_attn.py:290,
:353-357) is the hook an Ulysses-style sequence-parallel wrapper would use, and that wrapper is what
`ulysses_attention` provides: one all-to-all turns sequence-sharded Q/K/V into head-sharded full sequences, the
local call runs unchanged, one all-to-all brings the output back.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_heads(heads: int, world: int, rank: int):
    """(first_head, n_local_heads) of `rank`: contiguous, remainder heads go to the lowest ranks."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(heads, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def seq_to_head_shard(x: torch.Tensor, group=None) -> torch.Tensor:
    """[B, H, S/P, D] on every rank (sequence-sharded, all heads) -> [B, H/P, S, D] (all tokens, own heads).

    One all_to_all_single: rank r sends its tokens of head-group g to rank g.  H must divide by P and every rank
    must hold the same number of tokens (pad the sequence before sharding, as the Wan/CogVideoX paths do)."""
    world = dist.get_world_size(group)
    if world == 1:
        return x
    b, h, s_loc, d = x.shape
    if h % world:
        raise ValueError(f"{h} heads do not split over {world} ranks")
    hl = h // world
    # send buffer ordered by destination rank: [P, B, H/P, S/P, D]
    send = x.reshape(b, world, hl, s_loc, d).permute(1, 0, 2, 3, 4).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # recv[src] = tokens of rank src's sequence slice for my heads -> concatenate along the sequence
    return recv.permute(1, 2, 0, 3, 4).reshape(b, hl, world * s_loc, d)


def head_to_seq_shard(o: torch.Tensor, group=None) -> torch.Tensor:
    """[B, S, H/P, D] (all tokens, own heads; the `shape_xfuse=True` layout) -> [B, S/P, H, D] (own tokens, all
    heads).  Inverse exchange of seq_to_head_shard for the attention output."""
    world = dist.get_world_size(group)
    if world == 1:
        return o
    b, s, hl, d = o.shape
    if s % world:
        raise ValueError(f"{s} tokens do not split over {world} ranks")
    s_loc = s // world
    send = o.reshape(b, world, s_loc, hl, d).permute(1, 0, 2, 3, 4).contiguous()  # [P(dst), B, S/P, H/P, D]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # recv[src] = my tokens for rank src's heads -> concatenate along the head axis in rank order
    return recv.permute(1, 2, 0, 3, 4).reshape(b, s_loc, world * hl, d)


def ulysses_attention(q, k, v, attention_fn, group=None):
    """Sequence-parallel wrapper: q, k, v are [B, H, S/P, D] shards; `attention_fn(q, k, v)` takes [B, H/P, S, D]
    and returns [B, S, H/P, D] (call the per-model entry point with shape_xfuse=True).  Returns [B, S/P, H*D]."""
    ql, kl, vl = (seq_to_head_shard(t, group) for t in (q, k, v))
    o = attention_fn(ql, kl, vl)
    o = head_to_seq_shard(o, group)
    b, s_loc, h, d = o.shape
    return o.reshape(b, s_loc, h * d)
