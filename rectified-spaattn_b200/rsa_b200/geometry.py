"""Per-model block geometry of one attention call, reduced to the integers rsa_attn_desc carries.

Mirrors how each reference family derives seqlens / attenable / normal_blocks / text_end_block / padding:
  wan       rectified_wan21_attn.py:299-313
  hunyuan   rectified_hunyuan_attn.py:313-332   (needs S % 128 == 0; the reference raises otherwise, :356)
  flux      rectified_flux_attn.py:307-317
  cogvideo  rectified_cogvideo_attn.py:307-322
"""
from __future__ import annotations

from dataclasses import dataclass

BLOCK = 128


@dataclass(frozen=True)
class BlockGeometry:
    family: int            # native.FAMILY_WAN / FAMILY_JOINT
    seq: int
    n_blocks: int
    nq_blocks: int
    text_keys: int
    kv_len: int
    kv_zero_from: int
    text_end_block: int
    text_q_valid: int
    first_frame_blocks: int = 0


def _ceil_blocks(n):
    return (n + BLOCK - 1) // BLOCK


def wan(seq, first_frame_blocks=0):
    nb = _ceil_blocks(seq)
    return BlockGeometry(0, seq, nb, nb, 0, seq, seq, nb, 0, int(first_frame_blocks or 0))


def hunyuan(seq, num_true):
    if seq % BLOCK:
        # rectified_hunyuan_attn.py:356 reshapes value to (B,H,-1,128,D): a ragged S raises there
        raise RuntimeError(f"HunyuanVideo rectified attention needs S % 128 == 0, got S={seq}")
    nb = seq // BLOCK
    nq = nb - 256 // BLOCK
    attenable = 256 - (seq - num_true)
    if nq <= 0 or attenable < 1 or num_true > seq:
        raise ValueError(f"inconsistent HunyuanVideo geometry: S={seq}, num_true={num_true}")
    return BlockGeometry(1, seq, nb, nq, attenable, num_true, num_true, _ceil_blocks(num_true),
                         max(0, num_true - nq * BLOCK))


def flux(seq, text_length, kv_len=None):
    if seq % BLOCK:
        raise RuntimeError(f"Flux rectified attention needs S % 128 == 0, got S={seq}")
    kv_len = seq if kv_len is None else int(kv_len)
    nb = seq // BLOCK
    nq = nb - text_length // BLOCK
    if nq <= 0 or text_length < 1:
        raise ValueError(f"inconsistent Flux geometry: S={seq}, text_length={text_length}")
    return BlockGeometry(1, seq, nb, nq, text_length, kv_len, seq, _ceil_blocks(kv_len), seq - nq * BLOCK)


def cogvideo(seq, text_length):
    nb = _ceil_blocks(seq)
    pad = nb * BLOCK - seq
    nq = nb - (text_length + pad) // BLOCK
    if nq <= 0 or text_length < 1 or nq * BLOCK + text_length > nb * BLOCK:
        raise ValueError(f"inconsistent CogVideoX geometry: S={seq}, text_length={text_length}")
    return BlockGeometry(1, seq, nb, nq, text_length, seq, seq, nb, seq - nq * BLOCK)
