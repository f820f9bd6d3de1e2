"""GPU parity tests (run on the B200 box with `-m gpu`): every stage of the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): permutation bit-exact; pooled statistics, scores, GAPR mask bit-exact (same fp32
operation order as the oracle); block mask / kept lists / n_needed bit-exact when the oracle's selection is fed the
kernel's own fp32 probabilities (expf differs by ulps between CPU and GPU); probabilities, R, C within fp32
tolerance; attention output max-abs-err <= 2e-2 and cosine >= 0.999."""
import numpy as np
import pytest
import torch

from helpers import assert_prep_close, cos_sim, load_case, product_geometry
from oracle import cases as C
from oracle import gilbert_oracle as GO
from oracle import rsa_oracle as O

pytestmark = pytest.mark.gpu

ATOL_OUT, COS_OUT = 2e-2, 0.999


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsa_b200 import native
    assert native.lib().rsa_device_ok() == 1, "not an sm_100 device"
    return torch.device("cuda:0")


def _plan(case, dev, dump=True, impl=None):
    from rsa_b200 import ops
    q, k, v = (torch.from_numpy(case[n]).to(dev).to(torch.bfloat16) for n in ("q", "k", "v"))
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    nbr = torch.from_numpy(case["nbr"])
    return ops.Plan(q, k, v, geo, case["top_k"], case["p"], nbr, debug_dump_probs=dump)


# ------------------------------------------------------------------------------------------------ kernel 1
@pytest.mark.parametrize("shape", [(1, 1024, 3072), (2, 390, 1536), (1, 4096, 128), (1, 7, 8)])
def test_permute_bit_exact(dev, shape):
    from rsa_b200 import ops
    b, n, c = shape
    g = torch.Generator().manual_seed(n)
    x = torch.randn(b, n, c, generator=g).to(torch.bfloat16)
    idx = torch.randperm(n, generator=g)
    out = ops.permute_rows(x.to(dev), idx.to(dev)).cpu()
    ref = GO.permute_rows(x.view(torch.int16).numpy(), idx.numpy())
    assert np.array_equal(out.view(torch.int16).numpy(), ref)
    inv = torch.empty_like(idx)
    inv[idx] = torch.arange(n)
    back = ops.permute_rows(out.to(dev), inv.to(dev)).cpu()
    assert torch.equal(back.view(torch.int16), x.view(torch.int16))          # round trip


def test_permute_gilbert_fp32_rows(dev):
    """RoPE tables are fp32 [N, 128] and are permuted with the same index (scripts/main_hunyuan.py:89)."""
    from rsa_b200 import ops
    t, h, w = 4, 16, 16
    l2h, h2l = ops.gilbert_mapping(t, h, w)
    x = torch.randn(t * h * w, 128)
    out = ops.permute_rows(x.to(dev), h2l.to(dev)).cpu()
    assert torch.equal(out, x[h2l])
    assert torch.equal(ops.permute_rows(out.to(dev), l2h.to(dev)).cpu(), x)


# -------------------------------------------------------------------------------------------- kernels 2, 3a
@pytest.mark.parametrize("name", list(C.CASES))
def test_pool_scores_gapr_bit_exact(dev, name):
    case = load_case(name)
    plan = _plan(case, dev)
    plan.pool_stats()
    plan.block_scores()
    torch.cuda.synchronize()
    vw = plan.view()
    geo = case["ogeo"]
    nq, nv = geo.nq_blocks, geo.nq_blocks * 128
    for hi in range(case["heads"]):
        q, k, v = O.padded_inputs(case["q"][0, hi], case["k"][0, hi], case["v"][0, hi], geo)
        qp, dq = O.pool_stats(q, geo.seq + geo.gap, nq)
        kp, dk = O.pool_stats(k, geo.kv_zero_from, nq)
        vp, _ = O.pool_stats(v, geo.kv_zero_from, geo.n_blocks, want_mad=False)
        assert np.array_equal(vw["q_pool"][hi].cpu().numpy(), qp)
        assert np.array_equal(vw["q_mad"][hi].cpu().numpy(), dq)
        assert np.array_equal(vw["k_cat"][hi, :nq].cpu().numpy(), kp)
        assert np.array_equal(vw["k_mad"][hi].cpu().numpy(), dk)
        assert np.array_equal(vw["v_pool"][hi].cpu().numpy(), vp)
        kt = None
        if geo.family == "joint":
            kt = k[nv: nv + geo.text_keys]
            assert np.array_equal(vw["k_cat"][hi, nq:].cpu().numpy(), kt)
        a, nogapr = O.block_scores(qp, dq, kp, dk, kt)
        assert np.array_equal(vw["scores"][hi].cpu().numpy(), a)
        assert np.array_equal(vw["nogapr"][hi].cpu().numpy().astype(bool), nogapr)


# ------------------------------------------------------------------------------------------ kernels 3b, 3c
@pytest.mark.parametrize("name", list(C.CASES))
def test_select_and_rectify(dev, name):
    case = load_case(name)
    plan = _plan(case, dev)
    plan.pool_stats()
    plan.block_scores()
    plan.block_select()
    plan.rect_c()
    torch.cuda.synchronize()
    vw = plan.view()
    geo = case["ogeo"]
    nq, nb = geo.nq_blocks, geo.n_blocks
    mask_all = plan.dense_mask().cpu().numpy()
    for hi in range(case["heads"]):
        _, st = O.head_forward(case["q"][0, hi], case["k"][0, hi], case["v"][0, hi], geo, case["nbr"],
                               return_stages=True)
        p_gpu = vw["probs"][hi].cpu().numpy()
        np.testing.assert_allclose(p_gpu, st["probs"], rtol=3e-5, atol=1e-9)
        # selection is a discrete function of P: feed the kernel's own P to the oracle -> bit-exact
        m, n = O.select_blocks(p_gpu, geo, case["nbr"])
        assert np.array_equal(vw["n_needed"][hi].cpu().numpy(), n)
        assert np.array_equal(mask_all[hi, :nq], m)
        # kept lists = ascending set bits restricted to blocks that hold valid keys
        cnt = vw["kept_cnt"][hi].cpu().numpy()
        idx = vw["kept_idx"][hi].cpu().numpy().astype(np.int64) & 0xFFFF
        kvb = (geo.kv_len + 127) // 128
        for i in range(nq):
            want = np.nonzero(m[i, :kvb])[0]
            assert cnt[i] == len(want) and np.array_equal(idx[i, : cnt[i]], want)
        for i in range(nq, nb):                       # text query blocks: dense rows, R = 1, C = 0
            assert cnt[i] == kvb and np.array_equal(idx[i, :kvb], np.arange(kvb))
        r, c, part, w = O.rectify_factors(p_gpu, m, st["nogapr"], st["vp"], geo)
        np.testing.assert_allclose(vw["R"][hi, :nq].cpu().numpy(), r, rtol=0, atol=2e-6)
        np.testing.assert_allclose(vw["C"][hi, :nq].cpu().numpy(), c, rtol=1e-4, atol=2e-6)
        assert np.array_equal(vw["w_skip"][hi].cpu().numpy(), w)
        assert np.all(vw["R"][hi, nq:].cpu().numpy() == 1.0) and np.all(vw["C"][hi, nq:].cpu().numpy() == 0.0)
        # and the kernel's mask agrees with the pure-oracle mask except on ulp-level ties (none in these seeds)
        assert np.array_equal(m, st["mask"])


# ------------------------------------------------------------------------------------------------ kernel 4
def _impls():
    return [pytest.param(0, id="tcgen05"), pytest.param(1, id="mma_crosscheck")]


@pytest.mark.parametrize("impl", _impls())
def test_masked_attention_random_mask(dev, impl):
    """Kernel 4 alone: the surface of _triton_block_sparse_attention_onehot, ragged S, random 40 % mask (impl 1: the
    tests' own mma.sync kernel, tests/xcheck, held to the same oracle -- it is the checker at the full sizes)."""
    import xcheck
    from rsa_b200 import ops
    g = torch.Generator().manual_seed(5)
    h, s, s_valid = 3, 1000, 1000
    q, k, v = (torch.randn(1, h, s, 128, generator=g).to(torch.bfloat16) for _ in range(3))
    mask = torch.rand(1, h, 8, 8, generator=g) < 0.4
    mask |= torch.eye(8, dtype=torch.bool)
    fn = ops.masked_attention if impl == 0 else xcheck.masked_attention
    out = fn(q.to(dev), k.to(dev), v.to(dev), mask.to(dev), s_valid).float().cpu().numpy()
    for hi in range(h):
        ref = O.masked_attention(q[0, hi].float().numpy(), k[0, hi].float().numpy(), v[0, hi].float().numpy(),
                                 mask[0, hi].numpy(), s_valid, s, q_dtype="bf16")
        assert np.abs(out[0, hi] - ref).max() <= ATOL_OUT
        assert cos_sim(out[0, hi], ref) >= COS_OUT


@pytest.mark.parametrize("impl", _impls())
@pytest.mark.parametrize("name", list(C.CASES))
def test_end_to_end_vs_oracle(dev, name, impl):
    import xcheck
    case = load_case(name)
    if impl == 1 and case["ogeo"].gap:
        pytest.skip("the mma.sync cross-check kernel handles block-aligned visual segments only")
    plan = _plan(case, dev, dump=False)
    out = plan.run()                                    # [1, S, H, D]
    if impl == 1:                                       # the tests' own kernel 4 over the product's lists, R and C
        out = xcheck.sparse_attention(plan)
    out = out.float().cpu().numpy()
    ref = O.forward(case["q"], case["k"], case["v"], case["ogeo"], case["nbr"], q_dtype="bf16").reshape(out.shape)
    err = np.abs(out - ref).max()
    assert err <= ATOL_OUT, f"{name}: max-abs-err {err}"
    assert cos_sim(out, ref) >= COS_OUT


@pytest.mark.parametrize("name", [n for n in C.CASES if n not in C.REFERENCE_NEEDS_PADDED_LAYOUT])
def test_against_unmodified_reference_on_b200(dev, name, gold_dir):
    """tests/golden/golden_gpu_<case>.npz: output and block mask of the UNMODIFIED reference (bf16 mask arithmetic,
    Triton JIT kernel, flash-attn text rows) run on a B200 by oracle/ref_on_gpu.py.  The reference's bf16 mask differs
    from the fp32 one on near-tie entries by construction (iid inputs, whose pooled scores are almost
    uniform, are the worst case: 93 %), so: the masks must agree on >= 90 % of the entries, and on query blocks whose mask row agrees the outputs must meet the north-star tolerance."""
    import os
    case = load_case(name)
    g = np.load(os.path.join(gold_dir, f"golden_gpu_{name}.npz"))
    mask_ref = np.unpackbits(g["mask"])[: int(np.prod(g["mask_shape"]))].reshape(g["mask_shape"]).astype(bool)
    plan = _plan(case, dev, dump=False)
    out = plan.run().float().cpu().numpy()[0]                         # [S, H, D]
    ref = g["out"].astype(np.float32).reshape(out.shape)
    geo = case["ogeo"]
    nq = geo.nq_blocks
    mask = plan.dense_mask().cpu().numpy()[:, :nq]
    agree = (mask == mask_ref)
    assert agree.mean() >= 0.90, f"{name}: mask agreement {agree.mean():.4f}"
    rows_ok = agree.all(axis=2)                                       # [H, NQ]
    assert rows_ok.mean() >= 0.5, f"{name}: only {rows_ok.mean():.2f} of the query blocks comparable"
    rows = min(nq * 128, geo.seq)
    keep = np.repeat(rows_ok, 128, axis=1)[:, :rows].T                # [rows, H]
    d = np.abs(out[:rows] - ref[:rows])[keep]
    # Measured (tools/ref_gpu_golden_stats.py): six cases max <= 1.2e-2; flux_small / hunyuan_mid have isolated blocks
    # at 4.7e-2 / 3.1e-2 where the reference's bf16 GAPR test or bf16 R, C differ from fp32 -- the fp32 oracle shows
    # the same distance to the reference there (4.6e-2 / 3.0e-2), i.e. it is the reference's rounding, not this kernel.
    assert np.mean(d <= ATOL_OUT) >= 0.999, f"{name}: {np.mean(d <= ATOL_OUT):.5f} of the elements within {ATOL_OUT}"
    assert d.max() <= 6e-2, f"{name}: max-abs-err {d.max()}"
    assert cos_sim(out[:rows][keep], ref[:rows][keep]) >= COS_OUT
    t0, nt = nq * 128, geo.text_q_valid                               # text rows: dense in both implementations
    if nt:
        assert np.abs(out[t0: t0 + nt] - ref[t0: t0 + nt]).max() <= ATOL_OUT


@pytest.mark.parametrize("name", ["wan_c1", "hunyuan_mid", "cog_small", "hunyuan_ragged", "wan_300"])
def test_pair_schedule_is_a_permutation_with_common_prefix(dev, name):
    """Kernel 4 walks each kept list as [blocks all four tiles of its 2-CTA cluster keep] + [other blocks the tile pair
    keeps] + [the rest], each part ascending; the order of blocks does not change the attention result, the common
    prefixes let one K/V tile serve two tiles (shared-memory stage) or four (TMA multicast across the cluster)."""
    if name in C.MID_CASES:
        fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v = C.mid_case_inputs(name)
        case = dict(fam=fam, grid=(t, h, w), nv=nv, s=s, text_len=text_len, ntrue_d=ntrue_d, heads=heads, top_k=top_k, p=p,
                    q=q, k=k, v=v, nbr=GO.gilbert_block_neighbors(t, h, w))
    else:
        case = load_case(name)
    plan = _plan(case, dev, dump=False)
    plan.run()
    torch.cuda.synchronize()
    vw = plan.view()
    cnt = vw["kept_cnt"].cpu().numpy()
    kept = vw["kept_idx"].cpu().numpy().astype(np.int64) & 0xFFFF
    sched = vw["sched_idx"].cpu().numpy().astype(np.int64) & 0xFFFF
    nsh = vw["pair_shared"].cpu().numpy()
    nquad = vw["quad_shared"].cpu().numpy()
    bh, nqt = cnt.shape
    npairs = (nqt + 1) // 2
    nq_vis = plan.desc.nq_blocks if plan.desc.family == 1 else nqt
    vis_pairs = min(nq_vis, nqt) // 2                      # pairs of whole visual tiles: p and p ^ 1 are partners
    seen_quads = 0
    for h in range(bh):
        for p in range(npairs):
            t0, t1 = 2 * p, 2 * p + 1
            sets = [set(kept[h, t, : cnt[h, t]]) if t < nqt else set() for t in (t0, t1)]
            com = sets[0] & sets[1]
            quad = set()
            if p < vis_pairs and (p ^ 1) < vis_pairs:
                pt0 = 2 * (p ^ 1)
                quad = com & set(kept[h, pt0, : cnt[h, pt0]]) & set(kept[h, pt0 + 1, : cnt[h, pt0 + 1]])
                seen_quads += 1
            assert nsh[h, p] == len(com) and nquad[h, p] == len(quad)
            for t, mine in zip((t0, t1), sets):
                if t >= nqt:
                    continue
                row = list(sched[h, t, : cnt[h, t]])
                nq4, n2 = len(quad), len(com)
                assert row[:nq4] == sorted(quad)
                assert row[nq4:n2] == sorted(com - quad)
                assert row[n2:] == sorted(mine - com)
    if vis_pairs >= 2:
        assert seen_quads > 0


@pytest.mark.parametrize("impl", _impls())
def test_dense_limit_equals_sdpa(dev, impl):
    """top_k >= NB => every block kept => R = 1, C = 0 => plain dense attention."""
    import xcheck
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    g = torch.Generator().manual_seed(9)
    q, k, v = (torch.randn(1, 2, 640, 128, generator=g).to(torch.bfloat16).to(dev) for _ in range(3))
    plan = ops.Plan(q, k, v, G.wan(640), 99, 0.3, None)
    out = plan.run().view(1, 640, 2, 128)
    if impl == 1:
        out = xcheck.sparse_attention(plan)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).transpose(1, 2)
    assert (out.float() - ref).abs().max().item() <= ATOL_OUT


def test_entry_points_keep_reference_signatures(dev):
    """The per-model modules expose the reference's names/kwargs and return [B, S, H*D]."""
    from rectified_spaattn import rectified_cogvideo_attn as cog
    from rectified_spaattn import rectified_flux_attn as flux
    from rectified_spaattn import rectified_hunyuan_attn as hun
    from rectified_spaattn import rectified_wan21_attn as wan
    g = torch.Generator().manual_seed(1)
    mk = lambda s: tuple(torch.randn(1, 2, s, 128, generator=g).to(torch.bfloat16).to(dev) for _ in range(3))
    nbr = torch.from_numpy(GO.gilbert_block_neighbors(4, 16, 16))
    q, k, v = mk(1000)
    o = wan.rectified_block_sparse_attention(q, k, v, None, 2, block_neighbor_list=nbr, p_remain_rates=0.3,
                                             first_frame_blocks=2)
    assert o.shape == (1, 1000, 256)
    q, k, v = mk(1280)
    cu = torch.tensor([0, 1224, 1280], dtype=torch.int32, device=dev)
    am = (torch.arange(1280, device=dev) < 1224).view(1, 1, 1, -1)
    k0 = k.clone()
    o = hun.rectified_block_sparse_attention(q, k, v, attn_mask=am, top_k=2, cu_seqlens_q=cu, cu_seqlens_kv=cu,
                                             max_seqlen_q=1280, max_seqlen_kv=1280, block_neighbor_list=nbr,
                                             p_remain_rates=0.3)
    assert o.shape == (1, 1280, 256) and torch.equal(k, k0)
    assert torch.all(o[:, 1224:] == 0)
    q, k, v = mk(1536)
    cu = torch.tensor([0, 1536, 1536], dtype=torch.int32, device=dev)
    o = flux.rectified_block_sparse_attention(q, k, v, None, 1, cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=1536,
                                              max_seqlen_kv=1536, block_neighbor_list=nbr, p_remain_rates=0.3,
                                              text_length=512, shape_xfuse=True)
    assert o.shape == (1, 1536, 2, 128)
    q, k, v = mk(1250)
    o = cog.rectified_block_sparse_attention(q, k, v, None, 2, cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=1250,
                                             max_seqlen_kv=1250, block_neighbor_list=nbr, p_remain_rates=0.3,
                                             text_length=226)
    assert o.shape == (1, 1250, 256)
    with pytest.raises(NotImplementedError):
        wan.rectified_block_sparse_attention(q, k, v, None, 2, block_size_M=64)


# ------------------------------------------------------------------------------ host-buffer (pipelined) call
@pytest.mark.parametrize("name,hc", [("hunyuan_small", 1), ("wan_ragged", 2), ("cog_small", 1), ("hunyuan_mid", 3)])
def test_host_buffer_call_equals_device_call(dev, name, hc):
    """rsa_rectified_attention_host pipelines chunks of heads (H2D | kernels | D2H); heads are independent, so the
    result must be bit-identical to one device-resident call over all heads -- including a tail chunk (3 heads in
    chunks of 2) and chunks larger than the head count."""
    from rsa_b200 import ops
    case = load_case(name)
    # wan_ragged runs with 3 heads (its 2 + a copy of head 0) so that chunks of 2 leave a tail chunk of 1
    q, k, v = (torch.from_numpy(np.concatenate([case[n], case[n][:, :1]], axis=1) if hc == 2 else case[n])
               .to(torch.bfloat16) for n in ("q", "k", "v"))
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    nbr = torch.from_numpy(case["nbr"])
    want = ops.rectified_attention(q.to(dev), k.to(dev), v.to(dev), geo, case["top_k"], case["p"], nbr).cpu()
    hq, hk, hv = (t.pin_memory() for t in (q, k, v))
    got = ops.rectified_attention_host(hq, hk, hv, geo, case["top_k"], case["p"], nbr, heads_per_chunk=hc)
    torch.cuda.synchronize()
    assert not got.is_cuda and got.is_pinned() and got.shape == want.shape
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    # twice in a row on the same scratch (slot reuse across calls), and a [B,H,S,D]-strided host output
    b, h, s, d = q.shape
    out_hm = torch.empty(b, h, s, d, dtype=torch.bfloat16).pin_memory()
    from rsa_b200 import native as N
    import ctypes as CT
    desc = ops._fill_desc(N.AttnDesc(), (b, h, s, d), [ops._strides3(t) for t in (hq, hk, hv, out_hm)], geo,
                          case["top_k"], case["p"], ops._device_neighbors(nbr, dev))
    L = N.lib()
    need = L.rsa_host_call_scratch_bytes(CT.byref(desc), hc)
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    for _ in range(2):
        N.check(L.rsa_rectified_attention_host(CT.byref(desc), hq.data_ptr(), hk.data_ptr(), hv.data_ptr(),
                                               out_hm.data_ptr(), hc, scratch.data_ptr(), need,
                                               CT.c_void_p(torch.cuda.current_stream().cuda_stream)), "host call")
    torch.cuda.synchronize()
    assert torch.equal(out_hm.permute(0, 2, 1, 3).reshape(want.shape).view(torch.int16), want.view(torch.int16))


def test_entry_point_accepts_pinned_host_tensors(dev):
    """The per-family public call with page-locked host tensors returns a pinned host tensor (GPU arithmetic);
    pageable host tensors are refused rather than silently serialised."""
    from rectified_spaattn import rectified_hunyuan_attn as hun
    case = load_case("hunyuan_small")
    q, k, v = (torch.from_numpy(case[n]).to(torch.bfloat16) for n in ("q", "k", "v"))
    nbr = torch.from_numpy(case["nbr"])
    s, nt = case["s"], case["nv"] + case["ntrue_d"]
    kw = dict(attn_mask=None, top_k=case["top_k"], cu_seqlens_q=[0, nt, s], cu_seqlens_kv=[0, nt, s], max_seqlen_q=s,
              max_seqlen_kv=s, block_neighbor_list=nbr, p_remain_rates=case["p"])
    want = hun.rectified_block_sparse_attention(q.to(dev), k.to(dev), v.to(dev), **kw).cpu()
    got = hun.rectified_block_sparse_attention(q.pin_memory(), k.pin_memory(), v.pin_memory(), **kw)
    torch.cuda.synchronize()
    assert got.is_pinned() and torch.equal(got.view(torch.int16), want.view(torch.int16))
    with pytest.raises(RuntimeError):
        hun.rectified_block_sparse_attention(q, k, v, **kw)


def test_selection_ties_take_lowest_indices(dev):
    """Equal probabilities straddling the cut: the kernel sorts values only and recovers the selected set from the
    n-th value, taking tied entries in ascending index order -- the oracle's stable sort.  Keys are one block repeated
    (all columns tie) or two alternating blocks (two tie groups)."""
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    g = torch.Generator().manual_seed(3)
    nblk = 12
    for period, top_k, p in ((1, 1, 0.3), (2, 1, 0.55), (3, 5, 0.1), (1, 1, 0.999)):
        q = torch.randn(1, 2, nblk * 128, 128, generator=g)
        kb = torch.randn(1, 2, period, 128, 128, generator=g)
        k = kb.repeat(1, 1, nblk // period, 1, 1).reshape(1, 2, nblk * 128, 128)
        v = torch.randn(1, 2, nblk * 128, 128, generator=g)
        q, k, v = (t.to(torch.bfloat16).to(dev) for t in (q, k, v))
        plan = ops.Plan(q, k, v, G.wan(nblk * 128), top_k, p, None, debug_dump_probs=True)
        plan.pool_stats()
        plan.block_scores()
        plan.block_select()
        torch.cuda.synchronize()
        vw = plan.view()
        geo = O.geometry_wan(nblk * 128, top_k, p, 0)
        mask = plan.dense_mask().cpu().numpy()
        for hi in range(2):
            pg = vw["probs"][hi].cpu().numpy()
            assert len(np.unique(pg[0])) <= period          # the columns really tie
            m, n = O.select_blocks(pg, geo, None)
            assert np.array_equal(vw["n_needed"][hi].cpu().numpy(), n)
            assert np.array_equal(mask[hi], m)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_host_buffer_call_tapered_tail(dev, dtype):
    """With heads >= 3 * heads_per_chunk the last heads_per_chunk heads go one per chunk (the pipeline's tail is the
    last chunk's kernels + copy-out): 7 heads in chunks of 2 = chunks of 2, 2, 1 (body) + 1, 1 (tail).  Bit-identical
    to the device-resident call, in both dtypes."""
    from rsa_b200 import ops
    case = load_case("hunyuan_small")
    q, k, v = (torch.from_numpy(np.concatenate([case[n]] * 4, axis=1)[:, :7]).to(dtype) for n in ("q", "k", "v"))
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    nbr = torch.from_numpy(case["nbr"])
    want = ops.rectified_attention(q.to(dev), k.to(dev), v.to(dev), geo, case["top_k"], case["p"], nbr).cpu()
    got = ops.rectified_attention_host(*(t.pin_memory() for t in (q, k, v)), geo, case["top_k"], case["p"], nbr,
                                       heads_per_chunk=2)
    torch.cuda.synchronize()
    assert got.dtype == dtype and torch.equal(got.view(torch.int16), want.view(torch.int16))


# ----------------------------------------------------------------- batch > 1, strided views, CUDA-graph capture
@pytest.mark.parametrize("name", ["hunyuan_small", "wan_ragged"])
def test_batch_and_strided_views(dev, name):
    """B = 2 with the two batch entries swapped heads, passed as NON-contiguous [B, H, S, D] views of a [B, S, H, D]
    buffer (what `unflatten(2, (heads, -1)).transpose(1, 2)` of a projection output is): each (batch, head) must equal
    the single-batch contiguous result bit for bit."""
    from rsa_b200 import ops
    case = load_case(name)
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    nbr = torch.from_numpy(case["nbr"])
    q, k, v = (torch.from_numpy(case[n]).to(torch.bfloat16).to(dev) for n in ("q", "k", "v"))   # [1, 2, S, D]
    want = ops.rectified_attention(q, k, v, geo, case["top_k"], case["p"], nbr, shape_xfuse=True)  # [1, S, 2, D]

    def as_view(x):      # [2, 2, S, D] view over [2, S, 2, D]: batch 1 holds the heads in swapped order
        both = torch.cat([x, x.flip(1)], dim=0)
        return both.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)

    qv, kv, vv = as_view(q), as_view(k), as_view(v)
    assert not qv.is_contiguous()
    got = ops.rectified_attention(qv, kv, vv, geo, case["top_k"], case["p"], nbr, shape_xfuse=True)  # [2, S, 2, D]
    assert torch.equal(got[0].view(torch.int16), want[0].view(torch.int16))
    assert torch.equal(got[1].flip(1).view(torch.int16), want[0].view(torch.int16))


def test_call_is_cuda_graph_capturable(dev):
    """No allocation, no host synchronisation, tensor maps passed by value: the whole call records into a CUDA graph
    and replays on new input values (the reference has >= 4 host syncs per call and cannot be captured)."""
    from rsa_b200 import ops
    case = load_case("hunyuan_mid")
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    nbr = torch.from_numpy(case["nbr"])
    q, k, v = (torch.from_numpy(case[n]).to(torch.bfloat16).to(dev) for n in ("q", "k", "v"))
    plan = ops.Plan(q, k, v, geo, case["top_k"], case["p"], nbr)
    want = plan.run().clone()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        plan.run()                                   # warm-up on the capture stream
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=s):
            plan.run()
    plan.out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(plan.out.view(torch.int16), want.view(torch.int16))
    # new values in the captured input buffers -> the replay computes the new result
    q2 = torch.from_numpy(load_case("hunyuan_mid")["q"]).flip(1).to(torch.bfloat16).to(dev)
    q.copy_(q2)
    graph.replay()
    torch.cuda.synchronize()
    want2 = ops.Plan(q2, k, v, geo, case["top_k"], case["p"], nbr).run()
    torch.cuda.synchronize()
    assert torch.equal(plan.out.view(torch.int16), want2.view(torch.int16))


# ------------------------------------------------------------------- kernel 0: fused pre-attention sequence
def test_qkv_prep_against_oracle_and_torch(dev):
    """Kernel 0 on the golden inputs (390 ragged rows, 2 heads, rotary on the first 300): bit patterns against
    oracle/prep_oracle.py (which follows the kernel's reduction order; only rsqrt differs) and against the PyTorch op
    sequence of the reference's processor run on the GPU.  Bar (helpers.assert_prep_close): every element within one bf16 ulp of its
    rotation pair's magnitude, >= 99.9 % identical; V (a pure re-layout) bit-exact."""
    from oracle import make_golden as MG
    from oracle import prep_oracle as P
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    src, wq, wk, cos, sin, n_rope = MG.prep_inputs()
    rows = src[0].shape[1]
    q, k, v = (torch.zeros(1, 2, rows, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
    plan = ops.Plan(q, k, v, G.wan(rows), 1, 0.3, None)
    plan.qkv_prep(*(t.to(dev) for t in src), q_weight=wq, k_weight=wk, eps=1e-6, rope=(cos, sin), pool=False)
    torch.cuda.synchronize()
    for name, got, x, w, nr in (("q", q, src[0], wq, n_rope), ("k", k, src[1], wk, n_rope), ("v", v, src[2], None, 0)):
        want = P.prep(x.float().numpy(), 2, None if w is None else w.float().numpy(), 1e-6, cos.numpy(), sin.numpy(), nr)
        want = torch.from_numpy(want).to(torch.bfloat16)
        tref = MG.torch_prep(x.to(dev), 2, None if w is None else w.to(dev), 1e-6, cos.to(dev), sin.to(dev), nr).cpu()
        for ref, what in ((want, "oracle"), (tref, "torch on GPU")):
            assert_prep_close(got.float().cpu().numpy(), ref.float().numpy(), f"{name} vs {what}")
        if name == "v":
            assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16))


def test_qkv_prep_compact_rotary_table_is_bit_identical(dev):
    """diffusers' rotary tables repeat every cos / sin for the two elements of a pair; the host layer then hands kernel
    0 ONE [rows, 64] table of (cos_i, sin_i) pairs (rsa_prep_desc.rope_compact) -- the same numbers, half the bytes.
    Both layouts must give the same bits; a table that is not pair-repeated must take the two-table path."""
    from oracle import make_golden as MG
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    src, wq, wk, cos, sin, n_rope = MG.prep_inputs()
    rows = src[0].shape[1]
    cos, sin = cos.to(dev), sin.to(dev)
    assert bool((cos[:, 0::2] == cos[:, 1::2]).all()) and ops._compact_rope(cos, sin, cos, sin) is not None
    outs = []
    for compact in (True, False):
        ops.COMPACT_ROPE = compact
        try:
            q, k, v = (torch.zeros(1, 2, rows, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
            plan = ops.Plan(q, k, v, G.wan(rows), 1, 0.3, None)
            plan.qkv_prep(*(t.to(dev) for t in src), q_weight=wq, k_weight=wk, eps=1e-6, rope=(cos, sin), pool=True)
            torch.cuda.synchronize()
            outs.append((q, k, plan.view()["q_pool"].clone(), plan.view()["k_cat"].clone()))
        finally:
            ops.COMPACT_ROPE = True
    for a, b in zip(*outs):
        assert torch.equal(a.view(torch.int16) if a.dtype == torch.bfloat16 else a,
                           b.view(torch.int16) if b.dtype == torch.bfloat16 else b)
    odd = cos.clone()
    odd[:, 1] += 0.25                                   # no longer pair-repeated
    assert ops._compact_rope(odd, sin, odd, sin) is None


@pytest.mark.parametrize("name,dual", [("hunyuan_small", False), ("hunyuan_small", True), ("hunyuan_ragged", True),
                                       ("wan_ragged", False)])
def test_qkv_prep_pooled_path_equals_separate_kernels(dev, name, dual):
    """qkv_prep(pool=True) + run_pooled() must give exactly what kernel 2 computes from the rows kernel 0 stored, and the
    same attention output: pooled statistics bit-identical, output bit-identical.  `dual` = latent and encoder streams
    come from two source tensors (dual-stream block); otherwise one concatenated source (single-stream block)."""
    from rsa_b200 import ops
    case = load_case(name)
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    nbr = torch.from_numpy(case["nbr"])
    heads, s, nv = case["heads"], case["s"], case["nv"]
    g = torch.Generator().manual_seed(17)
    # projection outputs whose head split reproduces the case's Q, K, V: [1, S, H*128]
    src = [torch.from_numpy(case[n]).to(torch.bfloat16)[0].permute(1, 0, 2).reshape(1, s, heads * 128).contiguous().to(dev)
           for n in ("q", "k", "v")]
    wq = (1 + 0.1 * torch.randn(128, generator=g)).to(torch.bfloat16)
    wk = (1 + 0.1 * torch.randn(128, generator=g)).to(torch.bfloat16)
    ang = torch.outer(torch.arange(nv, dtype=torch.float32), 1.0 / (64.0 ** (torch.arange(0, 128, 2) / 128)))
    rope = (ang.cos().repeat_interleave(2, 1), ang.sin().repeat_interleave(2, 1))
    q, k, v = (torch.zeros(1, heads, s, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
    plan = ops.Plan(q, k, v, geo, case["top_k"], case["p"], nbr)
    if dual:
        plan.qkv_prep(*(t[:, :nv] for t in src), dst_row=0, q_weight=wq, k_weight=wk, rope=rope)
        plan.qkv_prep(*(t[:, nv:] for t in src), dst_row=nv, q_weight=wq, k_weight=wk)
    else:
        plan.qkv_prep(*src, dst_row=0, q_weight=wq, k_weight=wk, rope=rope, rope_rows=nv)
    torch.cuda.synchronize()
    vw = plan.view()
    fused = {n: vw[n].clone() for n in ("q_pool", "q_mad", "k_cat", "k_mad", "v_pool")}
    out_fused = plan.run_pooled().clone()
    torch.cuda.synchronize()
    # the rows kernel 0 stored, through the separate kernels
    plan2 = ops.Plan(q, k, v, geo, case["top_k"], case["p"], nbr)
    plan2.pool_stats()
    torch.cuda.synchronize()
    vw2 = plan2.view()
    for n in fused:
        assert torch.equal(fused[n].view(torch.int32), vw2[n].view(torch.int32)), n
    out_sep = plan2.run()
    torch.cuda.synchronize()
    assert torch.equal(out_fused.view(torch.int16), out_sep.view(torch.int16))
    # and the stored rows are the PyTorch sequence's (within a bf16 ulp)
    from oracle import make_golden as MG
    tq = MG.torch_prep(src[0], heads, wq.to(dev), 1e-6, rope[0].to(dev), rope[1].to(dev), nv)
    assert_prep_close(q.float().cpu().numpy(), tq.float().cpu().numpy(), "q vs torch on GPU")


def test_qkv_prep_wan_form(dev):
    """Kernel 0, Wan form: RMSNorm across heads (row statistics from row_rms_kernel) + rotary embedding on every token,
    against the oracle and against the literal Wan2.1 sequence (complex multiply in float64) on the GPU; then the pooled
    path against the separate kernels, bit for bit."""
    from oracle import make_golden as MG
    from oracle import prep_oracle as P
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    from test_oracle_golden import wan_prep_inputs
    src, wq, wk, cos, sin = wan_prep_inputs()
    rows = src[0].shape[1]
    q, k, v = (torch.zeros(1, 2, rows, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
    plan = ops.Plan(q, k, v, G.wan(rows, 1), 1, 0.3, None)
    plan.qkv_prep(*(t.to(dev) for t in src), q_weight=wq, k_weight=wk, eps=1e-6, rope=(cos, sin), pool=True)
    torch.cuda.synchronize()
    ang = torch.outer(torch.arange(rows, dtype=torch.float64),
                      1.0 / (256.0 ** (torch.arange(0, 128, 2, dtype=torch.float64) / 128)))
    freqs = torch.polar(torch.ones_like(ang), ang)[None, None].to(dev)
    for name, got, x, w in (("q", q, src[0], wq), ("k", k, src[1], wk)):
        want = P.prep(x.float().numpy(), 2, w.float().numpy(), 1e-6, cos.numpy(), sin.numpy(), rows)
        tref = MG.torch_prep_wan(x.to(dev), 2, w.to(dev), 1e-6, freqs).float().cpu().numpy()
        assert_prep_close(got.float().cpu().numpy(), want, f"{name} vs oracle")
        assert_prep_close(got.float().cpu().numpy(), tref, f"{name} vs Wan2.1 sequence on GPU")
    assert torch.equal(v.cpu().view(torch.int16),
                       src[2].unflatten(2, (2, -1)).transpose(1, 2).contiguous().view(torch.int16))
    vw = plan.view()
    fused = {n: vw[n].clone() for n in ("q_pool", "q_mad", "k_cat", "k_mad", "v_pool")}
    out_fused = plan.run_pooled().clone()
    plan2 = ops.Plan(q, k, v, G.wan(rows, 1), 1, 0.3, None)
    out_sep = plan2.run()
    torch.cuda.synchronize()
    vw2 = plan2.view()
    for n in fused:
        assert torch.equal(fused[n].view(torch.int32), vw2[n].view(torch.int32)), n
    assert torch.equal(out_fused.view(torch.int16), out_sep.view(torch.int16))


def test_fused_ulysses_two_gpus(dev):
    """The fused Ulysses exchange (kernel 0 gathers over peer memory, kernel 4's epilogue scatters) on 2 GPUs of one box,
    through torchrun: bit-identical to the NCCL all-to-all form on every rank and to the single-GPU call.  Skipped on a
    single-GPU box (the driver's 1-GPU tier); tools/check_fused_ulysses.py is the same check run by hand on 2-8 GPUs."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(repo, "tools", "check_fused_ulysses.py"), "c2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    # c2 = Wan: RMSNorm across heads (each rank computes the statistic of its own tokens, the gathering ranks read it over
    # NVLink); the bit reference is the single-GPU call (the NCCL form's norm is PyTorch's, other reduction order)
    assert line["fused_equals_single_gpu_bitwise_rank0"] and line["fraction_of_elements_equal_to_nccl_form_rank0"] > 0.9
    # HunyuanVideo form (per-head norm) on a small RAGGED visual segment: two gather calls, all three forms bit-identical
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29578",
                        os.path.join(repo, "tools", "check_fused_ulysses.py"), "c3b"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["fused_equals_nccl_form_bitwise_all_ranks"] and line["fused_equals_single_gpu_bitwise_rank0"]
    # tokens and heads that do not divide by the ranks (1035 tokens on 2 ranks: 518 + 517; 3 heads: 2 + 1), Wan form:
    # every rank's rows bit-identical to the single-GPU call
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29579",
                        os.path.join(repo, "tools", "check_fused_ulysses.py"), "odd"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["rows_last_rank"] < line["rows_per_rank"] and line["heads_rank0"] == 2
    assert line["fused_equals_single_gpu_bitwise_checked_ranks"] == [True, True]


def test_qkv_prep_cogvideo_form(dev):
    """Kernel 0, CogVideoX form: LayerNorm(head_dim) with weight and bias + rotary embedding on the video tokens, against
    the oracle and torch.nn.functional.layer_norm on the GPU."""
    from oracle import make_golden as MG
    from oracle import prep_oracle as P
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    src, _, _, cos, sin, n_rope = MG.prep_inputs()
    cw, cb = MG.cog_prep_params()
    rows = src[0].shape[1]
    q, k, v = (torch.zeros(1, 2, rows, 128, dtype=torch.bfloat16, device=dev) for _ in range(3))
    plan = ops.Plan(q, k, v, G.wan(rows), 1, 0.3, None)
    plan.qkv_prep(*(t.to(dev) for t in src), q_weight=cw[0], k_weight=cw[1], q_bias=cb[0], k_bias=cb[1], eps=1e-6,
                  rope=(cos, sin), pool=False)
    torch.cuda.synchronize()
    for i, (name, got) in enumerate((("q", q), ("k", k))):
        want = P.prep(src[i].float().numpy(), 2, cw[i].float().numpy(), 1e-6, cos.numpy(), sin.numpy(), n_rope,
                      bias=cb[i].float().numpy())
        tref = MG.torch_prep_cog(src[i].to(dev), 2, cw[i].to(dev), cb[i].to(dev), 1e-6, cos.to(dev), sin.to(dev), n_rope)
        assert_prep_close(got.float().cpu().numpy(), want, f"{name} vs oracle")
        assert_prep_close(got.float().cpu().numpy(), tref.float().cpu().numpy(), f"{name} vs torch on GPU")


def test_peer_buffers_single_process(dev):
    """CUDA IPC plumbing that does not need a second process: allocate, export a handle, wrap as a tensor, free."""
    import ctypes as CT
    from rsa_b200 import native as N
    from rsa_b200 import parallel
    L = N.lib()
    ptr = CT.c_void_p()
    N.check(L.rsa_peer_alloc(4096, CT.byref(ptr)), "rsa_peer_alloc")
    h = (CT.c_char * 64)()
    N.check(L.rsa_peer_export(ptr, h), "rsa_peer_export")
    assert any(b != 0 for b in bytes(h))
    t = parallel._peer_tensor(ptr.value, (2, 8, 128))
    assert t.dtype == torch.bfloat16 and t.data_ptr() == ptr.value and float(t.float().abs().max()) == 0.0   # zero-filled
    t.fill_(1.5)
    torch.cuda.synchronize()
    assert float(t.float().sum()) == 1.5 * 2048
    del t
    N.check(L.rsa_peer_free(ptr), "rsa_peer_free")
    assert L.rsa_peer_alloc(0, CT.byref(ptr)) == -1


@pytest.mark.parametrize("s,top_k,p", [(100, 1, 0.3), (128, 0, 0.0), (129, 1, 0.3), (700, 0, 0.0), (700, 2, 1.0),
                                        (700, 99, 0.3)])
def test_edge_geometries_wan(dev, s, top_k, p):
    """Shapes and parameters at the edges of the Wan path against the oracle: a single partial block, exactly one block,
    one token into the second block, top_k = 0 with p = 0 (one block per row + nothing else), p = 1 (every block by
    threshold), top_k >= NB; with an all-False neighbour matrix, which must equal passing None (the scripts' commented
    "linear settings")."""
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    q, k, v = O.synth_qkv(2, s, 128, "walk", 40 + s)
    nb = (s + 127) // 128
    tq, tk, tv = (torch.from_numpy(x).to(dev).to(torch.bfloat16) for x in (q, k, v))
    out = ops.rectified_attention(tq, tk, tv, G.wan(s), top_k, p, None).float().cpu().numpy()
    out2 = ops.rectified_attention(tq, tk, tv, G.wan(s), top_k, p, torch.zeros(nb, nb, dtype=torch.bool)).float().cpu().numpy()
    assert np.array_equal(out, out2)
    ref = O.forward(q, k, v, O.geometry_wan(s, top_k, p, 0), None, q_dtype="bf16")
    assert np.abs(out - ref).max() <= ATOL_OUT and cos_sim(out, ref) >= COS_OUT


def test_edge_geometries_joint(dev):
    """Joint family edges: a single visual block, one valid text token (a = 1), every text token valid, H = 1."""
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    for nv, ntrue_d, heads in ((128, 1, 1), (128, 256, 2), (384, 1, 2)):
        s = nv + 256
        q, k, v = O.synth_qkv(heads, s, 128, "cluster", 60 + ntrue_d)
        tq, tk, tv = (torch.from_numpy(x).to(dev).to(torch.bfloat16) for x in (q, k, v))
        out = ops.rectified_attention(tq, tk, tv, G.hunyuan(s, nv + ntrue_d), 1, 0.3, None).float().cpu().numpy()
        ref = O.forward(q, k, v, O.geometry_hunyuan(s, nv + ntrue_d, 1, 0.3), None, q_dtype="bf16")
        assert np.abs(out - ref).max() <= ATOL_OUT and cos_sim(out, ref) >= COS_OUT, (nv, ntrue_d, heads)
        assert np.all(out[0, nv + ntrue_d:] == 0)


# ------------------------------------------------ kernel 4: scores far above the running maximum
@pytest.mark.parametrize("gain", [3.0, 12.0], ids=["2^49", "beyond_2^128"])
def test_scores_far_above_running_maximum(dev, gain):
    """Kernel 4 rescales O and l lazily (only when a block's row maximum exceeds the reference maximum by more than
    2^8) and per warp.  Keys that match their query `gain` times over in a LATER block put that block's scores 2^49
    (gain 3) or more than 2^128 (gain 12: the rescale factor underflows to zero) above everything accumulated so far,
    for half of the rows of a warp's tile only -- the result must still be exact."""
    from rsa_b200 import ops
    g = torch.Generator().manual_seed(21)
    s = 640
    q = torch.randn(1, 2, s, 128, generator=g)
    k = 0.05 * torch.randn(1, 2, s, 128, generator=g)
    v = torch.randn(1, 2, s, 128, generator=g)
    k[:, :, 256:384] = 3.0 * q[:, :, 0:128]          # block 2 matches query tile 0 ...
    k[:, :, 384:512] = gain * q[:, :, 0:128]         # ... block 3 matches it `gain` times over
    k[:, :, 512:640] = gain * q[:, :, 128:256]       # and block 4 matches query tile 1
    q, k, v = (t.to(torch.bfloat16).to(dev) for t in (q, k, v))
    mask = torch.ones(1, 2, 5, 5, dtype=torch.bool, device=dev)
    out = ops.masked_attention(q, k, v, mask, s).float()
    # the reference kernel's arithmetic (wan21 :61-62, :90-102): q~ = bf16(q * sm_scale * log2 e), exp2 -- with logits of
    # this size the rounding of q~ alone moves the result by 6e-2 against exact attention
    qt = (q.float() * (128 ** -0.5 * 1.44269504)).to(torch.bfloat16).float()
    sc = qt @ k.float().transpose(-1, -2)
    pe = torch.exp2(sc - sc.max(dim=-1, keepdim=True).values)
    ref = (pe @ v.float()) / pe.sum(dim=-1, keepdim=True)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= ATOL_OUT
    assert cos_sim(out.cpu().numpy(), ref.cpu().numpy()) >= COS_OUT


# ------------------------------------------------------------------------------------------ fp16 tensors
def test_kernel4_fp16_against_the_literal_reference_kernel(dev, gold_dir):
    """tests/golden/kernel_fp16.npz holds the output of the reference's OWN Triton kernel
    (_triton_block_sparse_attention_onehot, rectified_wan21_attn.py:108-168) run under TRITON_INTERPRET=1 in fp16 --
    the one dtype the interpreter and this library share.  Kernel 4 on the same q, k, v, block mask and seqlen must
    reproduce it to fp16 rounding of P and of the output (the accumulation order inside a block differs)."""
    import os
    from rsa_b200 import ops
    g = np.load(os.path.join(gold_dir, "kernel_fp16.npz"))
    seqlen = int(g["seqlen"])
    q, k, v = (torch.from_numpy(g[n]).to(dev) for n in ("q", "k", "v"))
    assert q.dtype == torch.float16
    out = ops.masked_attention(q, k, v, torch.from_numpy(g["mask"]).to(dev), seqlen)
    assert out.dtype == torch.float16
    got, ref = out.float().cpu().numpy(), g["out"].astype(np.float32)
    assert np.abs(got[:, :, :seqlen] - ref[:, :, :seqlen]).max() <= 3e-3
    assert cos_sim(got[:, :, :seqlen], ref[:, :, :seqlen]) >= 0.99999


@pytest.mark.parametrize("name", ["wan_ragged", "hunyuan_small", "flux_small", "hunyuan_ragged"])
def test_end_to_end_fp16_vs_oracle(dev, name):
    """The whole call on fp16 tensors (rsa_attn_desc.dtype = RSA_DTYPE_F16; kernels 2 and 4 read and write fp16, the
    pooled statistics / scores / selection are fp32 either way) against the oracle on the same fp16-rounded values."""
    from rsa_b200 import ops
    case = load_case(name)
    q, k, v = (torch.from_numpy(case[n]).to(torch.float16) for n in ("q", "k", "v"))
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    plan = ops.Plan(q.to(dev), k.to(dev), v.to(dev), geo, case["top_k"], case["p"], torch.from_numpy(case["nbr"]))
    out = plan.run()
    assert out.dtype == torch.float16
    out = out.float().cpu().numpy()
    ref = O.forward(q.float().numpy(), k.float().numpy(), v.float().numpy(), case["ogeo"], case["nbr"],
                    q_dtype="fp16").reshape(out.shape)
    assert np.abs(out - ref).max() <= ATOL_OUT and cos_sim(out, ref) >= COS_OUT
    # pooled statistics of the fp16 rows: bit-exact like the bf16 ones
    vw = plan.view()
    ogeo = case["ogeo"]
    qh, kh, vh = O.padded_inputs(q[0, 0].float().numpy(), k[0, 0].float().numpy(), v[0, 0].float().numpy(), ogeo)
    qp, dq = O.pool_stats(qh, ogeo.seq + ogeo.gap, ogeo.nq_blocks)
    assert np.array_equal(vw["q_pool"][0].cpu().numpy(), qp) and np.array_equal(vw["q_mad"][0].cpu().numpy(), dq)
    with pytest.raises(RuntimeError):                     # kernel 0 is bf16 only
        plan.qkv_prep(*(torch.zeros(1, case["s"], case["heads"] * 128, dtype=torch.float16, device=dev) for _ in range(3)))
    with pytest.raises(RuntimeError):                     # mixed dtypes
        ops.Plan(q.to(dev), k.to(dev).to(torch.bfloat16), v.to(dev), geo, case["top_k"], case["p"], None)


# ------------------------------------------------------------------------------- head_dim 64 (CogVideoX)
def test_head_dim_64(dev):
    """CogVideoX has 64-dimensional heads (the reference kernel takes Lk in {16, 32, 64, 128}, wan21 :121).  The
    kernels work on 128 columns and read the missing 64 as zeros (kernel 2 by predication, kernel 4 through TMA's
    out-of-bounds fill), with the scale of the real head_dim; no copy.  Whole call against the oracle on the 64-column
    tensors; pooled statistics and scores bit-exact (the extra fmaf steps add exact zeros); kernel 4 alone against
    SDPA; the host-buffer call bit-identical to the device call."""
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    heads, t, h, w, text = 2, 4, 16, 16, 226
    nv = t * h * w
    s = nv + text
    q, k, v = O.synth_qkv(heads, s, 64, "walk", 31)
    nbr = GO.gilbert_block_neighbors(t, h, w)
    tq, tk, tv = (torch.from_numpy(a).to(dev).to(torch.bfloat16) for a in (q, k, v))
    geo = G.cogvideo(s, text)
    out = ops.rectified_attention(tq, tk, tv, geo, 2, 0.3, torch.from_numpy(nbr))
    assert out.shape == (1, s, heads * 64)
    ogeo = O.geometry_cogvideo(s, text, 2, 0.3)
    ref = O.forward(q, k, v, ogeo, nbr, q_dtype="bf16")
    got = out.float().cpu().numpy()
    assert np.abs(got - ref).max() <= ATOL_OUT and cos_sim(got, ref) >= COS_OUT
    plan = ops.Plan(tq, tk, tv, geo, 2, 0.3, torch.from_numpy(nbr), debug_dump_probs=True)
    plan.pool_stats()
    plan.block_scores()
    plan.block_select()
    torch.cuda.synchronize()
    vw = plan.view()
    _, st = O.head_forward(q[0, 0], k[0, 0], v[0, 0], ogeo, nbr, return_stages=True)
    assert np.array_equal(vw["q_pool"][0, :, :64].cpu().numpy(), st["qp"]) and not vw["q_pool"][0, :, 64:].any()
    assert np.array_equal(vw["scores"][0].cpu().numpy(), st["scores"])
    np.testing.assert_allclose(vw["probs"][0].cpu().numpy(), st["probs"], rtol=3e-5, atol=1e-9)
    # kernel 4 alone (dense): the surface fullattn(mode="flash") uses on CogVideoX's warm-up steps
    mask = torch.ones(1, heads, (s + 127) // 128, (s + 127) // 128, dtype=torch.bool, device=dev)
    dense = ops.masked_attention(tq, tk, tv, mask, s)
    sd = torch.nn.functional.scaled_dot_product_attention(tq.float(), tk.float(), tv.float())
    assert dense.shape == tq.shape and (dense.float() - sd).abs().max().item() <= ATOL_OUT
    # the 64-column instantiation of kernel 4 (4 k-steps of Q K^T, N = 64 P V) against the 128-column one reading the
    # same tensors with the second granule zero-filled by TMA: the extra MMAs add exact zeros
    ops.set_attention_flags(4)
    try:
        wide = ops.rectified_attention(tq, tk, tv, geo, 2, 0.3, torch.from_numpy(nbr))
        dense_wide = ops.masked_attention(tq, tk, tv, mask, s)
    finally:
        ops.set_attention_flags(0)
    # (bit-identical when both evaluate the exponentials the same way; the 64-column form runs a quarter of them as an
    # FMA-pipe polynomial of relative error 7.5e-5 by default, far below P's bf16 rounding: at most one output ulp)
    for x, y in ((wide, out), (dense_wide, dense)):
        d = (x.float() - y.float()).abs()
        assert bool((d <= 2.0 ** -7 * y.float().abs() + 1e-3).all()) and d.mean().item() <= 2e-4
    # fp16 tensors through the 64-column instantiation
    hq16, hk16, hv16 = (x.to(torch.float16) for x in (tq, tk, tv))
    dense16 = ops.masked_attention(hq16, hk16, hv16, mask, s)
    sd16 = torch.nn.functional.scaled_dot_product_attention(hq16.float(), hk16.float(), hv16.float())
    assert (dense16.float() - sd16).abs().max().item() <= ATOL_OUT
    # host-buffer call
    hq, hk, hv = (x.cpu().pin_memory() for x in (tq, tk, tv))
    hout = ops.rectified_attention_host(hq, hk, hv, geo, 2, 0.3, torch.from_numpy(nbr))
    torch.cuda.synchronize()
    assert torch.equal(hout.view(torch.int16), out.cpu().view(torch.int16))
    with pytest.raises(RuntimeError):                     # kernel 0 is built for 128 columns
        plan.qkv_prep(*(torch.zeros(1, s, heads * 64, dtype=torch.bfloat16, device=dev) for _ in range(3)))


def test_triton_mirror_per_batch_kv_len_and_head_dim_64(dev):
    """_triton_block_sparse_attention_onehot(q, k, v, seqlens, block_mask, sm_scale): the reference kernel loads
    seqlens[off_hz // H] per batch element (rectified_wan21_attn.py:36-37, :86) and takes head_dim 64 (:121) -- the
    mirror does both, on strided [B, H, S, D] views without copies."""
    from rectified_spaattn.rectified_cogvideo_attn import _triton_block_sparse_attention_onehot as cog_kernel
    from rectified_spaattn.rectified_wan21_attn import _triton_block_sparse_attention_onehot as kernel
    g = torch.Generator().manual_seed(17)
    for d, fn in ((128, kernel), (64, cog_kernel)):
        b, h, s = 2, 2, 512
        # [B, S, H, D] storage viewed as [B, H, S, D]: token stride H*D, head stride D (what a processor holds)
        q, k, v = (torch.randn(b, s, h, d, generator=g).to(torch.bfloat16).to(dev).transpose(1, 2) for _ in range(3))
        mask = torch.rand(b, h, 4, 4, generator=g) < 0.5
        mask |= torch.eye(4, dtype=torch.bool)
        lens = torch.tensor([512, 300], dtype=torch.int32, device=dev)
        out = fn(q, k, v, lens, mask.to(dev), d ** -0.5, 128, 128).float().cpu().numpy()
        for bi in range(b):
            for hi in range(h):
                ref = O.masked_attention(q[bi, hi].float().cpu().numpy(), k[bi, hi].float().cpu().numpy(),
                                         v[bi, hi].float().cpu().numpy(), mask[bi, hi].numpy(), int(lens[bi]), s,
                                         q_dtype="bf16")
                assert np.isfinite(out[bi, hi]).all()
                assert np.abs(out[bi, hi] - ref).max() <= ATOL_OUT and cos_sim(out[bi, hi], ref) >= COS_OUT
