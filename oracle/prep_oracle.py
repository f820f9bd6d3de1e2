"""ORACLE (test infrastructure, not product code) -- the pre-attention sequence of the reference's joint-text processors.

Restates rectified_hunyuan_attn.py:448-479 (identical in rectified_flux_attn.py) for one source tensor:
  * head split          query.unflatten(2, (heads, -1)).transpose(1, 2)                          :448-450
  * QK normalisation    attn.norm_q(query), attn.norm_k(key)                                     :453-456
  * rotary embedding    apply_rotary_emb(query[:, :, :n_rope], image_rotary_emb), rest untouched :459-478
The two library functions are NOT under /root/reference: they come from diffusers (requirements.txt:18 pins
diffusers==0.34.0; absent from this image).  Their published algorithms, restated here:
  * diffusers.models.normalization.RMSNorm.forward (elementwise_affine, bf16 weight):
        variance = x.to(float32).pow(2).mean(-1, keepdim=True)
        x = x * rsqrt(variance + eps)            # fp32
        x = x.to(weight.dtype) * weight          # rounded to bf16, multiplied, rounded to bf16
  * diffusers.models.embeddings.apply_rotary_emb(x, (cos, sin), use_real=True, use_real_unbind_dim=-1):
        x_real, x_imag = x.reshape(..., -1, 2).unbind(-1)
        x_rotated = stack([-x_imag, x_real], -1).flatten(3)
        out = (x.float() * cos + x_rotated.float() * sin).to(x.dtype)
Parity pin: tests/golden/prep_hunyuan.npz holds the output of the literal PyTorch op sequence above (a torch.nn.Module
replica of RMSNorm + the expression of apply_rotary_emb) run in bf16 on CPU by oracle/make_golden.py; diffusers itself is
not available to run, so this pin is against the restated library code ("parity unpinned" against the binary wheel).

Only tests/ may import this."""
from __future__ import annotations

import numpy as np


def bf16_round(x):
    """fp32 -> nearest-even bfloat16, returned as fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def _fma32(a, b, c):
    # fp32 fma: the product of two fp32 values is exact in float64
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def mean_square_kernel_order(x):
    """mean(x^2) over the last axis (128) in the CUDA kernel's order: lane l of 16 owns columns 8l..8l+7 (one fma
    chain), then a butterfly over the lanes (xor 1, 2, 4, 8)."""
    x = np.asarray(x, dtype=np.float32)
    lanes = x.reshape(*x.shape[:-1], 16, 8)
    ss = np.zeros(lanes.shape[:-1], dtype=np.float32)
    for c in range(8):
        ss = _fma32(lanes[..., c], lanes[..., c], ss)
    for o in (1, 2, 4, 8):
        idx = np.arange(16) ^ o
        ss = (ss + ss[..., idx]).astype(np.float32)
    return (ss[..., 0] * np.float32(1.0 / 128.0)).astype(np.float32)


def rms_norm(x, weight, eps):
    """x [..., 128] (bf16 values held in fp32), weight [128] (bf16 values) -> bf16 values."""
    var = mean_square_kernel_order(x)
    rinv = (1.0 / np.sqrt(var.astype(np.float64) + np.float64(np.float32(eps)))).astype(np.float32)
    xn = bf16_round((np.asarray(x, np.float32) * rinv[..., None]).astype(np.float32))
    return bf16_round((xn * np.asarray(weight, np.float32)).astype(np.float32))


def layer_norm(x, weight, bias, eps):
    """torch.nn.LayerNorm(128) on bf16 values (CogVideoX norm_q / norm_k, rectified_cogvideo_attn.py:452-455): statistics
    and affine in fp32, one rounding to bf16.  Kernel order: per lane a sequential sum of its 8 columns, a butterfly over
    the 16 lanes; two-pass variance (fma chain of the centred values, same butterfly)."""
    x = np.asarray(x, dtype=np.float32)
    lanes = x.reshape(*x.shape[:-1], 16, 8)
    sm = np.zeros(lanes.shape[:-1], dtype=np.float32)
    for c in range(8):
        sm = (sm + lanes[..., c]).astype(np.float32)
    for o in (1, 2, 4, 8):
        sm = (sm + sm[..., np.arange(16) ^ o]).astype(np.float32)
    mean = (sm[..., 0] * np.float32(1.0 / 128.0)).astype(np.float32)
    d = (lanes - mean[..., None, None]).astype(np.float32)
    ss = np.zeros(lanes.shape[:-1], dtype=np.float32)
    for c in range(8):
        ss = _fma32(d[..., c], d[..., c], ss)
    for o in (1, 2, 4, 8):
        ss = (ss + ss[..., np.arange(16) ^ o]).astype(np.float32)
    var = (ss[..., 0] * np.float32(1.0 / 128.0)).astype(np.float32)
    rstd = (1.0 / np.sqrt(var.astype(np.float64) + np.float64(np.float32(eps)))).astype(np.float32)
    t = (d.reshape(x.shape) * rstd[..., None]).astype(np.float32)
    return bf16_round(_fma32(np.broadcast_to(np.asarray(weight, np.float32), x.shape), t,
                             np.broadcast_to(np.asarray(bias, np.float32), x.shape)))


def mean_square_row_kernel_order(x):
    """mean(x^2) over the last axis (heads*128 channels of a token) in row_rms_kernel's order: lane l of 32 runs one fma
    chain over columns 8(l + 32 j) .. +7 for j = 0, 1, ..; then a butterfly over the lanes (xor 1, 2, 4, 8, 16)."""
    x = np.asarray(x, dtype=np.float32)
    inner = x.shape[-1]
    ss = np.zeros(x.shape[:-1] + (32,), dtype=np.float32)
    for j in range((inner + 255) // 256):
        for lane in range(32):
            c0 = 8 * (lane + 32 * j)
            if c0 >= inner:
                continue
            for c in range(8):
                ss[..., lane] = _fma32(x[..., c0 + c], x[..., c0 + c], ss[..., lane])
    for o in (1, 2, 4, 8, 16):
        idx = np.arange(32) ^ o
        ss = (ss + ss[..., idx]).astype(np.float32)
    return (ss[..., 0] / np.float32(inner)).astype(np.float32)


def rms_norm_across_heads(x, weight, eps):
    """Wan: RMSNorm over all heads*128 channels of a token (rectified_wan21_attn.py:423-426; diffusers RMSNorm with
    dim = inner_dim).  x [..., inner] bf16 values, weight [inner] -> bf16 values."""
    var = mean_square_row_kernel_order(x)
    rinv = (1.0 / np.sqrt(var.astype(np.float64) + np.float64(np.float32(eps)))).astype(np.float32)
    xn = bf16_round((np.asarray(x, np.float32) * rinv[..., None]).astype(np.float32))
    return bf16_round((xn * np.asarray(weight, np.float32)).astype(np.float32))


def rotary(x, cos, sin):
    """x [..., S, 128], cos/sin [S, 128] fp32 -> bf16 values; pairs are (x[2i], x[2i+1])."""
    x = np.asarray(x, np.float32)
    rot = np.empty_like(x)
    rot[..., 0::2] = -x[..., 1::2]
    rot[..., 1::2] = x[..., 0::2]
    a = (x * cos).astype(np.float32)
    b = (rot * sin).astype(np.float32)
    return bf16_round((a + b).astype(np.float32))


def prep(src, heads, weight=None, eps=1e-6, cos=None, sin=None, rope_rows=0, bias=None):
    """src [B, rows, heads*128] -> [B, heads, rows, 128] after head split, RMSNorm (if weight: [128] = per head as in
    HunyuanVideo / Flux, [heads*128] = across heads as in Wan, applied BEFORE the split) and rotary embedding on the
    first rope_rows tokens."""
    b, rows, _ = src.shape
    src = np.asarray(src, np.float32)
    if weight is not None and np.asarray(weight).size == heads * 128:
        src = rms_norm_across_heads(src, weight, eps)
        weight = None
    x = src.reshape(b, rows, heads, 128).transpose(0, 2, 1, 3)
    if weight is not None and bias is not None:
        x = layer_norm(x, weight, bias, eps)
    elif weight is not None:
        x = rms_norm(x, weight, eps)
    if rope_rows:
        x = x.copy()
        x[:, :, :rope_rows] = rotary(x[:, :, :rope_rows], cos[:rope_rows], sin[:rope_rows])
    return x
