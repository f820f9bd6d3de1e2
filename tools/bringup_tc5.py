"""Bring-up diagnostics for the tcgen05 attention kernel (run on the GPU box): stage-by-stage comparison of the
kernel's TMEM dumps (S of the first block, un-normalised O, l, m) with fp32 PyTorch, then full-output checks against
the mma.sync cross-check kernel and SDPA.  Not a test: prints numbers, never asserts."""
import ctypes as C
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "rectified-spaattn_b200"))
from rsa_b200 import native, ops  # noqa: E402

dev = torch.device("cuda:0")
L = native.lib()


sys.path.insert(0, os.path.join(REPO, "tests"))
import xcheck  # noqa: E402  (the tests' mma.sync kernel: its own library)


def run(q, k, v, mask, kv_len, impl):
    o = (ops.masked_attention if impl == 0 else xcheck.masked_attention)(q, k, v, mask, kv_len)
    torch.cuda.synchronize()
    return o


def stats(name, got, ref):
    d = (got.float() - ref.float()).abs()
    cos = torch.nn.functional.cosine_similarity(got.float().flatten(), ref.float().flatten(), dim=0).item()
    print(f"  {name}: max-abs {d.max().item():.4e}  mean-abs {d.mean().item():.4e}  ref-mean-abs "
          f"{ref.float().abs().mean().item():.4e}  cos {cos:.6f}  nan {int(torch.isnan(got.float()).sum())}", flush=True)


def one_block():
    print("== one block, S=128, 1 head", flush=True)
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(1, 1, 128, 128, generator=g).to(torch.bfloat16).to(dev) for _ in range(3))
    mask = torch.ones(1, 1, 1, 1, dtype=torch.bool, device=dev)
    dbg = torch.zeros(34048, dtype=torch.float32, device=dev)
    L.rsa_debug_set_attention_dump(C.c_void_p(dbg.data_ptr()))
    o = run(q, k, v, mask, 128, 0)
    L.rsa_debug_set_attention_dump(None)
    qf, kf, vf = q[0, 0].float(), k[0, 0].float(), v[0, 0].float()
    s_ref = qf @ kf.T
    stats("S = Q K^T", dbg[:16384].view(128, 128), s_ref)
    stats("S^T (if transposed)", dbg[:16384].view(128, 128).T, s_ref)
    sc = 128 ** -0.5 * 1.4426950408889634
    m = dbg[32768 + 128: 32768 + 256]
    stats("m", m, s_ref.max(dim=1).values * sc)
    p = torch.exp2(s_ref * sc - m[:, None])
    stats("l", dbg[32768: 32768 + 128], p.sum(dim=1))
    o_raw_ref = p.to(torch.bfloat16).float() @ vf
    stats("O raw = P V", dbg[16384:32768].view(128, 128), o_raw_ref)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    stats("out vs sdpa", o, ref)
    if (dbg[16384:32768].view(128, 128) - o_raw_ref).abs().max() > 0.1:
        got = dbg[16384:32768].view(128, 128)
        # which columns / rows are off?
        err = (got - o_raw_ref).abs()
        print("  O err per 16-col group:", [round(err[:, c:c + 16].max().item(), 3) for c in range(0, 128, 16)])
        print("  O err per 16-row group:", [round(err[r:r + 16].max().item(), 3) for r in range(0, 128, 16)])
    if (dbg[:16384].view(128, 128) - s_ref).abs().max() > 0.1:
        got = dbg[:16384].view(128, 128)
        err = (got - s_ref).abs()
        print("  S err per 16-col group:", [round(err[:, c:c + 16].max().item(), 3) for c in range(0, 128, 16)])
        print("  S err per 16-row group:", [round(err[r:r + 16].max().item(), 3) for r in range(0, 128, 16)])


def multi_block(h, s, s_valid, dens, seed):
    print(f"== h={h} S={s} valid={s_valid} density={dens}", flush=True)
    g = torch.Generator().manual_seed(seed)
    q, k, v = (torch.randn(1, h, s, 128, generator=g).to(torch.bfloat16).to(dev) for _ in range(3))
    nb = (s + 127) // 128
    mask = (torch.rand(1, h, nb, nb, generator=g) < dens) | torch.eye(nb, dtype=torch.bool)
    mask = mask.to(dev)
    o0 = run(q, k, v, mask, s_valid, 0)
    o1 = run(q, k, v, mask, s_valid, 1)
    stats("tcgen05 vs mma.sync", o0, o1)
    if dens >= 1.0:
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k[:, :, :s_valid].float(),
                                                               v[:, :, :s_valid].float())
        stats("tcgen05 vs sdpa", o0, ref)
        stats("mma.sync vs sdpa", o1, ref)


def perf(h, s, dens):
    print(f"== perf h={h} S={s} density={dens}", flush=True)
    g = torch.Generator(device=dev).manual_seed(1)
    q, k, v = (torch.randn(1, h, s, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(3))
    nb = (s + 127) // 128
    mask = (torch.rand(1, h, nb, nb, generator=g, device=dev) < dens) | torch.eye(nb, dtype=torch.bool, device=dev)
    pairs = int(mask.sum().item())
    for impl in (0, 1):
        for _ in range(2):
            run(q, k, v, mask, s, impl)
        fn = ops.masked_attention if impl == 0 else xcheck.masked_attention
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(q, k, v, mask, s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"  impl {impl}: {ms:.3f} ms (incl. mask->lists)  {pairs * 8388608 / ms / 1e9:.1f} TFLOP/s on kept pairs",
              flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["one", "multi", "perf"]
    if "one" in what:
        one_block()
    if "multi" in what:
        multi_block(1, 256, 256, 1.0, 1)
        multi_block(2, 1024, 1024, 1.0, 2)
        multi_block(3, 1000, 1000, 0.4, 5)
        multi_block(2, 4096, 4000, 0.3, 6)
    if "perf" in what:
        perf(8, 16384, 0.25)
        perf(4, 65536, 0.22)


def trace(h=8, s=16384, dens=0.25, flags=0):
    """Clock trace of one CTA in the middle of a busy launch (slots: see attn_tc5.cu RSA_TRACE)."""
    print(f"== trace h={h} S={s} density={dens}", flush=True)
    g = torch.Generator(device=dev).manual_seed(1)
    q, k, v = (torch.randn(1, h, s, 128, generator=g, device=dev).to(torch.bfloat16) for _ in range(3))
    nb = (s + 127) // 128
    mask = (torch.rand(1, h, nb, nb, generator=g, device=dev) < dens) | torch.eye(nb, dtype=torch.bool, device=dev)
    dbg = torch.zeros(34048, dtype=torch.float32, device=dev)
    run(q, k, v, mask, s, 0)
    L.rsa_debug_set_attention_dump(C.c_void_p(dbg.data_ptr()))
    L.rsa_debug_set_attention_flags(flags)
    run(q, k, v, mask, s, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.masked_attention(q, k, v, mask, s)
    e1.record()
    torch.cuda.synchronize()
    L.rsa_debug_set_attention_flags(0)
    L.rsa_debug_set_attention_dump(None)
    print(f"  debug kernel, flags={flags}: {e0.elapsed_time(e1) / 5:.3f} ms per call "
          f"({int(mask.sum()) * 8388608 / (e0.elapsed_time(e1) / 5) / 1e9:.0f} TFLOP/s-equivalent)")
    t = dbg[33024:].view(64, 16).cpu().double()
    names = ["sm:wait_s", "sm:got_s", "sm:ld_done", "sm:max_done", "sm:half0", "sm:done", "", "", "mma:qk_issued",
             "mma:got_p0", "mma:pv_enter", "mma:pv_issued", "mma:k_in", "mma:v_in", "mma:qk_enter"]
    print("  step " + " ".join(f"{n:>13}" for n in names if n))
    for i in range(4, 16):
        print(f"  {i:4d} " + " ".join(f"{int(t[i, j]):13d}" for j, n in enumerate(names) if n))
    a, b = t[8:24], t[9:25]
    f = lambda x: f"{x.mean():.0f}"
    print(f"  period {f(b[:,1]-a[:,1])}; wait_s {f(a[:,1]-a[:,0])}; ld {f(a[:,2]-a[:,1])}; max {f(a[:,3]-a[:,2])}; "
          f"exp half0 {f(a[:,4]-a[:,3])}; exp half1 {f(a[:,5]-a[:,4])}; half0->mma got_p0 {f(a[:,9]-a[:,4])}; "
          f"got_p0->pv_issued {f(a[:,11]-a[:,9])}; pv_issued->qk_issued(next) {f(b[:,8]-a[:,11])}; "
          f"qk_issued->got_s {f(b[:,1]-b[:,8])}")


if __name__ == "__main__" and "trace" in sys.argv[1:]:
    trace()
if __name__ == "__main__" and "ablate" in sys.argv[1:]:
    trace(4, 65536, 0.22, flags=0)
    trace(4, 65536, 0.22, flags=1)
    trace(4, 65536, 0.22, flags=3)
