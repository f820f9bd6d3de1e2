"""ORACLE support (test infrastructure): run the UNMODIFIED reference (Triton JIT kernel + flash-attn + PyTorch eager
mask builder, bf16) on the B200 box.

  python -m oracle.ref_on_gpu golden          -> gpurun_out/golden_gpu_<case>.npz   (outputs for tests/golden/)
  python -m oracle.ref_on_gpu time c2 c3a ... -> gpurun_out/ref_timing.json         (the "bar to beat")
  python -m oracle.ref_on_gpu golden_c3a      -> gpurun_out/golden_gpu_c3a.npz      (full-size fixture: packed mask +
                                                                                     three output rows per block)

Needs baseline/_ref (oracle/stage_reference.py) or /root/reference.
"""
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "rectified-spaattn_b200"))

from oracle import ref_loader  # noqa: E402
from oracle.cases import CASES, case_inputs  # noqa: E402

OUT = os.path.join(REPO, "gpurun_out")
MOD = {"wan": "rectified_wan21_attn", "hunyuan": "rectified_hunyuan_attn", "flux": "rectified_flux_attn",
       "cogvideo": "rectified_cogvideo_attn"}


def call_ref(ref, fam, q, k, v, nbr, top_k, p, s, num_true, text_len, ffb):
    dev = q.device
    if fam == "wan":
        return ref.rectified_block_sparse_attention(q, k, v, None, top_k, block_neighbor_list=nbr, p_remain_rates=p,
                                                    first_frame_blocks=ffb)
    if fam == "hunyuan":
        am = (torch.arange(s, device=dev) < num_true).view(1, 1, 1, s)
        cu = torch.tensor([0, num_true, s], dtype=torch.int32, device=dev)
        return ref.rectified_block_sparse_attention(q, k, v, am, top_k, cu_seqlens_q=cu, cu_seqlens_kv=cu,
                                                    max_seqlen_q=s, max_seqlen_kv=s, block_neighbor_list=nbr,
                                                    p_remain_rates=p)
    cu = torch.tensor([0, s, s], dtype=torch.int32, device=dev)
    return ref.rectified_block_sparse_attention(q, k, v, None, top_k, cu_seqlens_q=cu, cu_seqlens_kv=cu,
                                                max_seqlen_q=s, max_seqlen_kv=s, block_neighbor_list=nbr,
                                                p_remain_rates=p, text_length=text_len)


def golden():
    dev = torch.device("cuda:0")
    from oracle import gilbert_oracle as GO
    for name in CASES:
        fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v = case_inputs(name)
        ref = ref_loader.load([MOD[fam]])[MOD[fam]]
        nbr = torch.from_numpy(GO.gilbert_block_neighbors(t, h, w))
        tq, tk, tv = (torch.from_numpy(x).to(dev).to(torch.bfloat16) for x in (q, k, v))
        ffb = (nv + 127) // 128 // t if fam == "wan" else None
        cap = {}
        orig = ref._build_block_index_with_importance_optimized

        def spy(*a, **kw):
            r = orig(*a, **kw)
            cap["mask"] = r[0].clone()
            return r

        ref._build_block_index_with_importance_optimized = spy
        out = call_ref(ref, fam, tq.clone(), tk.clone(), tv.clone(), nbr, top_k, p, s, nv + ntrue_d, text_len, ffb)
        ref._build_block_index_with_importance_optimized = orig
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(OUT, f"golden_gpu_{name}.npz"),
                            out=out[0].float().cpu().numpy().astype(np.float16),
                            mask=np.packbits(cap["mask"][0].cpu().numpy().astype(np.uint8)),
                            mask_shape=np.array(cap["mask"][0].shape))
        print("golden", name, tuple(out.shape), float(out.float().abs().mean()), flush=True)


C3A_HEADS = (0, 23)          # heads of bench.py's generator the full-size fixture holds
C3A_ROWS = (0, 64, 127)      # output rows kept per 128-token block


def golden_c3a():
    """HunyuanVideo 128 frames (C3a, the largest shape the reference runs): the unmodified reference in bf16 on two
    heads of bench.py's inputs.  Kept: its block mask (packed bits) and rows C3A_ROWS of every block of its output."""
    sys.argv = [sys.argv[0]]
    import bench
    from oracle import gilbert_oracle as GO  # noqa: F401
    from rsa_b200 import ops
    dev = torch.device("cuda:0")
    wp = bench.workload_params("c3a")
    ref = ref_loader.load([MOD["hunyuan"]])[MOD["hunyuan"]]
    t, h, w = wp["grid"]
    nbr = ops.gilbert_block_neighbors(t, h, w)
    qs, ks, vs = zip(*(bench.synth_heads_device(1, hd, wp["s"], "walk", dev) for hd in C3A_HEADS))
    q, k, v = (torch.cat(x, dim=1) for x in (qs, ks, vs))
    cap = {}
    orig = ref._build_block_index_with_importance_optimized

    def spy(*a, **kw):
        r = orig(*a, **kw)
        cap["mask"] = r[0].clone()
        return r

    ref._build_block_index_with_importance_optimized = spy
    run = lambda: call_ref(ref, "hunyuan", q, k.clone(), v.clone(), nbr, wp["top_k"], bench.P_REMAIN, wp["s"],
                           wp["num_true"], wp["text"], None)
    out = run()
    ref._build_block_index_with_importance_optimized = orig
    # the reference's own R and C (bf16 arithmetic, hunyuan :348-357): with its kernel replaced by a constant the visual
    # rows of the result are C (kernel = 0) and R + C (kernel = 1), both constant per 128-row block
    kernel = ref._triton_block_sparse_attention_onehot
    ref._triton_block_sparse_attention_onehot = lambda q_, *a, **kw: torch.zeros_like(q_)
    out_c = run()
    ref._triton_block_sparse_attention_onehot = lambda q_, *a, **kw: torch.ones_like(q_)
    out_rc = run()
    ref._triton_block_sparse_attention_onehot = kernel
    torch.cuda.synchronize()
    nq = wp["nv"] // 128
    first = torch.arange(0, nq * 128, 128, device=dev)
    c_ref = out_c[0].view(wp["s"], len(C3A_HEADS), 128)[first]                       # [NQ, H, 128] bf16
    r_ref = (out_rc[0].view(wp["s"], len(C3A_HEADS), 128)[first].float() - c_ref.float()).mean(dim=-1)   # [NQ, H]
    o = out[0].view(wp["s"], len(C3A_HEADS), 128)
    rows = (torch.arange(0, wp["s"], 128, device=dev)[:, None] + torch.tensor(C3A_ROWS, device=dev)[None]).flatten()
    np.savez_compressed(os.path.join(OUT, "golden_gpu_c3a.npz"), rows=rows.cpu().numpy(),
                        out=o[rows].float().cpu().numpy().astype(np.float16),
                        R=r_ref.cpu().numpy().astype(np.float32),
                        C_bf16=c_ref.contiguous().view(torch.int16).cpu().numpy(),
                        mask=np.packbits(cap["mask"][0].cpu().numpy().astype(np.uint8)),
                        mask_shape=np.array(cap["mask"][0].shape), heads=np.array(C3A_HEADS))
    print("golden c3a", tuple(out.shape), tuple(cap["mask"].shape), float(cap["mask"].float().mean()), flush=True)


def timing(names):
    sys.argv = [sys.argv[0]]
    import bench
    dev = torch.device("cuda:0")
    from rsa_b200 import ops
    res = {}
    for name in names:
        wp = bench.workload_params(name)
        wp["name"] = name
        fam = wp["fam"]
        ref = ref_loader.load([MOD[fam]])[MOD[fam]]
        t, h, w = wp["grid"]
        nbr = ops.gilbert_block_neighbors(t, h, w)
        q, k, v = bench.synth_heads_device(wp["heads"], 0, wp["s"], "walk", dev)
        stage = {}
        orig_build = ref._build_block_index_with_importance_optimized
        orig_kernel = ref._triton_block_sparse_attention_onehot

        def timed_wrap(key, fn):
            def inner(*a, **kw):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = fn(*a, **kw)
                torch.cuda.synchronize()
                stage.setdefault(key, []).append((time.perf_counter() - t0) * 1e3)
                if key == "kernel":
                    stage["kept_pairs"] = int(a[4].sum().item())
                return r
            return inner

        def run(instrument):
            ref._build_block_index_with_importance_optimized = timed_wrap("mask_build", orig_build) if instrument else orig_build
            ref._triton_block_sparse_attention_onehot = timed_wrap("kernel", orig_kernel) if instrument else orig_kernel
            kq, kk, kv = q, k.clone(), v.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = call_ref(ref, fam, kq, kk, kv, nbr, wp["top_k"], bench.P_REMAIN, wp["s"], wp["num_true"], wp["text"],
                         wp["ffb_blocks"])
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1), o

        try:
            for _ in range(2):
                run(False)
            tot = sorted(run(False)[0] for _ in range(5))
            stage.clear()
            for _ in range(3):
                run(True)
            res[name] = dict(ms_per_call=tot[len(tot) // 2], ms_all=tot,
                             mask_build_ms=sorted(stage["mask_build"])[1], kernel_ms=sorted(stage["kernel"])[1],
                             kept_pairs=stage.get("kept_pairs"),
                             dense_equiv_tflops=wp["dense_flop_per_head"] * wp["heads"] / (tot[len(tot) // 2] * 1e-3) / 1e12)
        except Exception as e:  # noqa: BLE001
            res[name] = dict(error=repr(e)[:400])
        finally:
            ref._build_block_index_with_importance_optimized = orig_build
            ref._triton_block_sparse_attention_onehot = orig_kernel
        print("ref timing", name, res[name], flush=True)
        del q, k, v
        torch.cuda.empty_cache()
    with open(os.path.join(OUT, "ref_timing.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1] == "golden":
        golden()
    elif sys.argv[1] == "golden_c3a":
        golden_c3a()
    else:
        timing(sys.argv[2:] or ["c2"])
