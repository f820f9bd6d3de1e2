"""Drop-in mirror of the reference rectified_spaattn/rectified_wan22_attn.py: Wan2.2 re-uses the Wan2.1 hot path
(reference :12 `from .rectified_wan21_attn import rectified_block_sparse_attention`); only the processors differ
(cos/sin RoPE tables on the [B, S, H, D] layout, other warm-up gates, 80-call cycle for the A14B pair)."""
from . import _processors as _P
from .attn import fullattn  # noqa: F401
from .rectified_wan21_attn import rectified_block_sparse_attention  # noqa: F401


class RectifiedWanTI2VSpaAttnProcessor2_0(_P.WanProcessorBase):
    """Wan2.2 TI2V-5B (reference :15-163): sparse from layer 2 and call 10 on; cycle of 100 calls."""

    rope = "cos_sin"

    def sparse_now(self):
        return self.processor_id >= 2 and self.current_step >= 10


class _Wan22A14B(_P.WanProcessorBase):
    """Wan2.2 T2V / I2V A14B (two transformers of 40 layers share one processor numbering 0..79): layers
    0, 1, 40, 41 stay dense, sparse from call `warm_steps` on; cycle of 80 calls (reference :166-288, :291-414)."""

    rope = "cos_sin"
    steps_per_cycle = 80

    def __init__(self, mode, select_block_num, block_neighbor_list, p_remain_rates, processor_id=0,
                 first_frame_blocks=0, warm_steps=0):
        super().__init__(mode, select_block_num, block_neighbor_list, p_remain_rates, processor_id,
                         first_frame_blocks)
        self.warm_steps = warm_steps

    def sparse_now(self):
        return self.processor_id not in (0, 1, 40, 41) and self.current_step >= self.warm_steps


class RectifiedWanT2VSpaAttnProcessor2_0(_Wan22A14B):
    pass


class RectifiedWanI2VSpaAttnProcessor2_0(_Wan22A14B):
    pass
