"""Per-model block geometry of one attention call, reduced to the integers rsa_attn_desc carries.

Mirrors how each reference family derives seqlens / attenable / normal_blocks / text_end_block / padding:
  wan       rectified_wan21_attn.py:299-313
  hunyuan   rectified_hunyuan_attn.py:313-332   (the reference needs S % 128 == 0 and raises otherwise, :356; here a
            ragged visual segment -- 129 frames = 118 800 tokens -- is completed with zero rows, see hunyuan())
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
    vis_len: int = 0       # joint family: visual tokens when they do not fill whole blocks (0 = nq_blocks * 128)

    @property
    def gap(self):
        """Zero rows that complete the last visual block in the padded layout (not present in memory)."""
        return self.nq_blocks * BLOCK - self.vis_len if self.vis_len else 0


def _ceil_blocks(n):
    return (n + BLOCK - 1) // BLOCK


def wan(seq, first_frame_blocks=0):
    nb = _ceil_blocks(seq)
    return BlockGeometry(0, seq, nb, nb, 0, seq, seq, nb, 0, int(first_frame_blocks or 0))


def hunyuan(seq, num_true, text_len=256):
    """HunyuanVideo: `text_len` (256) text tokens last, `num_true` = visual + valid text tokens.

    Block-aligned visual segment (S % 128 == 0): exactly the reference's live branch (:313-332).
    Ragged visual segment (the 129-frame shape, 118 800 + 256): the reference raises (:356).  Extension, following the
    rule its Wan/CogVideoX paths use for ragged inputs (wan21 :299-302): the last visual block is completed with
    zero rows -- pooled as zeros with divisor 128, never attended as keys, never written as queries -- and the text
    blocks follow block-aligned.  kv_len / kv_zero_from / text_end_block are positions in that padded layout."""
    nv = seq - text_len
    attenable = text_len - (seq - num_true)
    if nv <= 0 or attenable < 1 or num_true > seq:
        raise ValueError(f"inconsistent HunyuanVideo geometry: S={seq}, num_true={num_true}")
    nq = _ceil_blocks(nv)
    gap = nq * BLOCK - nv
    nb = nq + _ceil_blocks(text_len)
    valid_v = num_true + gap                      # end of the valid keys in the padded layout
    return BlockGeometry(1, seq, nb, nq, attenable, valid_v, valid_v, _ceil_blocks(valid_v), num_true - nv,
                         vis_len=nv if gap else 0)


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
