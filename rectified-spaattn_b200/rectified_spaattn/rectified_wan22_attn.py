"""Drop-in mirror of the reference rectified_spaattn/rectified_wan22_attn.py: Wan2.2 re-uses the Wan2.1 hot path
(reference :12 `from .rectified_wan21_attn import rectified_block_sparse_attention`); only processors differ."""
from .rectified_wan21_attn import rectified_block_sparse_attention  # noqa: F401
