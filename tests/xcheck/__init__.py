"""TEST INFRASTRUCTURE: the mma.sync (HMMA) cross-check implementation of kernel 4 in its own library
(tests/xcheck/librsa_xcheck.so, built by rectified-spaattn_b200/build_native.py::build_xcheck).  The product library has
one attention kernel and no run-time switch; the tests run this one over the kept-block lists, R and C the PRODUCT's
stages left in a plan's workspace, or over lists built from a dense block mask, and compare outputs."""
import ctypes as C
import math
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librsa_xcheck.so")
_lib = None


class Args(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
                ("batch", C.c_int32), ("heads", C.c_int32),
                ("qs", C.c_int64 * 3), ("ks", C.c_int64 * 3), ("vs", C.c_int64 * 3), ("os", C.c_int64 * 3),
                ("seq_q", C.c_int32), ("seq_kv", C.c_int32), ("kv_len", C.c_int32), ("q_valid", C.c_int32),
                ("vis_len", C.c_int32), ("nq_vis", C.c_int32), ("gap", C.c_int32), ("nqt", C.c_int32), ("nb", C.c_int32),
                ("kept_idx", C.c_void_p), ("kept_cnt", C.c_void_p), ("R", C.c_void_p), ("C", C.c_void_p),
                ("scale_log2", C.c_float), ("q_round", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: python rectified-spaattn_b200/build_native.py builds it")
        L = C.CDLL(LIB_PATH)
        L.rsa_xcheck_last_error.restype = C.c_char_p
        L.rsa_xcheck_args_size.restype = C.c_size_t
        assert L.rsa_xcheck_args_size() == C.sizeof(Args)
        L.rsa_xcheck_attention.argtypes = [C.POINTER(Args), C.c_void_p]
        _lib = L
    return _lib


def _launch(a, dev):
    with torch.cuda.device(dev):
        rc = lib().rsa_xcheck_attention(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"rsa_xcheck_attention failed ({rc}): {lib().rsa_xcheck_last_error().decode()}")


def _scale_log2(d):
    return float(np.float32((1.0 / math.sqrt(d)) * 1.44269504))


def sparse_attention(plan):
    """The cross-check kernel over the lists, R and C in `plan`'s workspace (run the product's stages first) -> a new
    [B, S, H, D] tensor in the product's output layout.  bf16, head_dim 128, block-aligned visual segment only."""
    b, h, s, d = plan.shape
    assert d == 128 and plan.q.dtype == torch.bfloat16
    g = plan.desc
    vw = plan.view()
    out = torch.zeros((b, s, h, d), dtype=plan.q.dtype, device=plan.device)
    o4 = out.permute(0, 2, 1, 3)
    a = Args()
    a.q, a.k, a.v, a.o = plan.q.data_ptr(), plan.k.data_ptr(), plan.v.data_ptr(), out.data_ptr()
    a.batch, a.heads = b, h
    for name, t in (("qs", plan.q), ("ks", plan.k), ("vs", plan.v), ("os", o4)):
        arr = getattr(a, name)
        for i in range(3):
            arr[i] = t.stride(i)
    joint = g.family == 1
    vis_len = min(g.vis_len if g.vis_len > 0 else g.nq_blocks * 128, s) if joint else s
    nq_vis = (vis_len + 127) // 128
    gap = nq_vis * 128 - vis_len if joint else 0
    a.seq_q = a.seq_kv = s
    a.kv_len = g.kv_len
    a.q_valid = min(g.nq_blocks * 128 + g.text_q_valid, s + gap) if joint else s
    a.vis_len, a.nq_vis, a.gap = vis_len, nq_vis, gap
    a.nqt = a.nb = g.n_blocks
    a.kept_idx, a.kept_cnt = vw["kept_idx"].data_ptr(), vw["kept_cnt"].data_ptr()
    a.R, a.C = vw["R"].data_ptr(), vw["C"].data_ptr()
    a.scale_log2, a.q_round = _scale_log2(d), 1
    _launch(a, plan.device)
    return out


def masked_attention(q, k, v, block_mask, kv_len):
    """The surface of _triton_block_sparse_attention_onehot through the cross-check kernel: q, k, v [B, H, S, 128] bf16,
    block_mask bool [B, H, NQ, NB] -> [B, H, S, 128].  The lists are built here, in PyTorch."""
    b, h, s, d = q.shape
    assert d == 128 and q.dtype == torch.bfloat16
    q, k, v = (t.contiguous() for t in (q, k, v))
    nqb, nkb = block_mask.shape[-2:]
    m = block_mask.reshape(b * h, nqb, nkb).to(q.device).bool().clone()
    m[:, :, (kv_len + 127) // 128:] = False
    cnt = m.sum(dim=2).to(torch.int32).contiguous()
    order = torch.argsort((~m).to(torch.int8), dim=2, stable=True)          # kept block indices first, ascending
    idx = order.to(torch.int16).contiguous()
    out = torch.zeros_like(q)
    a = Args()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.batch, a.heads = 1, b * h
    for name, t in (("qs", q), ("ks", k), ("vs", v), ("os", out)):
        arr = getattr(a, name)
        arr[0], arr[1], arr[2] = 0, t.stride(1), t.stride(2)
    a.seq_q, a.seq_kv, a.kv_len, a.q_valid = s, k.shape[2], int(kv_len), s
    a.vis_len, a.nq_vis, a.gap = 1 << 30, 1 << 20, 0
    a.nqt, a.nb = nqb, nkb
    a.kept_idx, a.kept_cnt, a.R, a.C = idx.data_ptr(), cnt.data_ptr(), None, None
    a.scale_log2, a.q_round = _scale_log2(d), 1
    _launch(a, q.device)
    torch.cuda.synchronize()
    return out
