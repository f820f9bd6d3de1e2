"""A stand-in for diffusers.models.attention_processor.Attention (diffusers is not installed): just the attributes
the reference processors touch: to_q/to_k/to_v, heads, norm_q/norm_k, add_{q,k,v}_proj,
norm_added_{q,k}, to_out[0..1], to_add_out, is_cross_attention, prepare_attention_mask."""
import torch
from torch import nn


class FakeAttention(nn.Module):
    def __init__(self, dim, heads, head_dim=128, added=False, norm="head", image_ctx=False):
        super().__init__()
        inner = heads * head_dim
        self.heads = heads
        self.is_cross_attention = False
        self.to_q, self.to_k, self.to_v = (nn.Linear(dim, inner) for _ in range(3))
        nd = inner if norm == "inner" else head_dim         # Hunyuan/Flux/Cog norm per head; Wan over the inner dim
        if norm == "layer":                                  # CogVideoX: LayerNorm(head_dim) with bias
            self.norm_q, self.norm_k = nn.LayerNorm(nd, eps=1e-6), nn.LayerNorm(nd, eps=1e-6)
            for m in (self.norm_q, self.norm_k):
                nn.init.normal_(m.weight, 1.0, 0.1)
                nn.init.normal_(m.bias, 0.0, 0.05)
        else:
            self.norm_q, self.norm_k = nn.RMSNorm(nd, eps=1e-6), nn.RMSNorm(nd, eps=1e-6)
        self.add_q_proj = self.add_k_proj = self.add_v_proj = None
        self.norm_added_q = self.norm_added_k = None
        self.to_add_out = None
        if added:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = (nn.Linear(dim, inner) for _ in range(3))
            self.norm_added_q, self.norm_added_k = nn.RMSNorm(nd, eps=1e-6), nn.RMSNorm(nd, eps=1e-6)
            self.to_add_out = nn.Linear(inner, dim)
        if image_ctx:                                        # Wan I2V: only K/V projections of the CLIP tokens
            self.add_k_proj, self.add_v_proj = nn.Linear(dim, inner), nn.Linear(dim, inner)
            self.norm_added_k = nn.RMSNorm(inner, eps=1e-6)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        return attention_mask
