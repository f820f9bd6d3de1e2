"""Kernel 4 with many heads: the grid order (in-order branch of the decode, re-paired tail) and the 2-CTA clusters with
their multicast prefix, each against the form it replaced and against the independent mma.sync kernel."""
import os
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

pytestmark = pytest.mark.gpu

ATOL_OUT = 2e-2


def _bench():
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        import bench
    finally:
        sys.argv = argv
    return bench


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("grid,text", [((8, 32, 32), 256), ((5, 34, 47), 256)])
def test_grid_order_with_many_heads(dev, grid, text):
    """Kernel 4's grid puts the text pairs of the last few heads first and walks the other heads in order (and re-pairs
    the tail when the number of visual tiles is odd).  With the 2 heads of the other tests every head is in the front
    set; here 24 heads x 64 (63) visual blocks make the host's estimate 21, so heads 0..2 take the in-order branch.  The
    output must agree with the former order (attention flag 16: head by head, (2p, 2p+1) pairing -- the order the other
    tests validated for five sessions) and, where the visual segment is block-aligned, with the independent mma.sync
    kernel, which has its own indexing."""
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    bench = _bench()
    t, h, w = grid
    nv, heads, top_k = t * h * w, 24, 16
    s = nv + text
    geo = G.hunyuan(s, nv + 200, text)
    nbr = ops.gilbert_block_neighbors(t, h, w)
    q, k, v = bench.synth_heads_device(heads, 0, s, "walk", dev, seed=3)
    plan = ops.Plan(q, k, v, geo, top_k, 0.3, nbr, private_workspace=True)
    out = plan.run().clone()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out.float()).all())
    ops.set_attention_flags(16)
    try:
        former = plan.sparse_attention().clone()
        torch.cuda.synchronize()
    finally:
        ops.set_attention_flags(0)
    # the same arithmetic per tile where the pairing is unchanged; the re-paired text tiles walk their lists in another
    # order, which can flip the bf16 rounding of an output element: the absolute bar plus two output ulps
    d = (out.float() - former.float()).abs()
    assert bool((d <= ATOL_OUT + 2.0 ** -6 * former.float().abs()).all()), f"new vs former grid order: max-abs {float(d.max()):.4f}"
    cos = float(torch.nn.functional.cosine_similarity(out.float().flatten(), former.float().flatten(), dim=0))
    assert cos >= 0.9999
    if geo.gap == 0:
        import xcheck
        ref_all = xcheck.sparse_attention(plan)
        torch.cuda.synchronize()
        # two independent bf16 kernels: the absolute bar, plus one output ulp where |o| >= 2 (the first run of this test
        # measured exactly one ulp, 2^-5, on an element of magnitude 4..8)
        d = (out.float() - ref_all.float()).abs()
        assert bool((d <= ATOL_OUT + 2.0 ** -7 * ref_all.float().abs()).all()), f"tcgen05 vs mma.sync max-abs {float(d.max()):.4f}"
        cos = float(torch.nn.functional.cosine_similarity(out.float().flatten(), ref_all.float().flatten(), dim=0))
        assert cos >= 0.9999


def test_clusters_against_the_kernel_without_them(dev):
    """Two CTAs with consecutive grid ids form a cluster and fetch the blocks four adjacent query tiles share once (TMA
    multicast, releases multicast onto both rings).  Attention flag 32 launches the same kernel without clusters and
    schedules the lists without the quad prefix (the round-1 skeleton): the same kept sets walked in another order, so
    the outputs agree within the output bar plus an ulp, and the cluster form must also agree with the mma.sync kernel.
    24 heads x 64 visual blocks, top_k 32: long common prefixes, partner CTAs of different list lengths, the odd pair at
    the end of a head whose partner belongs to the next head (no sharing there)."""
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    import xcheck
    bench = _bench()
    t, h, w, text = 8, 32, 32, 256
    nv, heads, top_k = t * h * w, 24, 32
    s = nv + text
    geo = G.hunyuan(s, nv + 200, text)
    nbr = ops.gilbert_block_neighbors(t, h, w)
    q, k, v = bench.synth_heads_device(heads, 0, s, "walk", dev, seed=5)
    plan = ops.Plan(q, k, v, geo, top_k, 0.3, nbr, private_workspace=True)
    out = plan.run().clone()
    torch.cuda.synchronize()
    vw = plan.view()
    assert int(vw["quad_shared"].max().item()) >= 8, "the case must exercise the multicast prefix"
    ops.set_attention_flags(32)
    try:
        plain = plan.run().clone()
        torch.cuda.synchronize()
    finally:
        ops.set_attention_flags(0)
    d = (out.float() - plain.float()).abs()
    assert bool((d <= ATOL_OUT + 2.0 ** -6 * plain.float().abs()).all()), f"clusters vs none: max-abs {float(d.max()):.4f}"
    assert float(torch.nn.functional.cosine_similarity(out.float().flatten(), plain.float().flatten(), dim=0)) >= 0.9999
    plan.run()
    ref_all = xcheck.sparse_attention(plan)
    torch.cuda.synchronize()
    d = (out.float() - ref_all.float()).abs()
    assert bool((d <= ATOL_OUT + 2.0 ** -7 * ref_all.float().abs()).all()), f"tcgen05 vs mma.sync max-abs {float(d.max()):.4f}"


def test_selection_with_power_of_two_visual_blocks(dev):
    """Kernel 3b leaves the joint family's text aggregate out of the sorting network and inserts it afterwards exactly
    when that halves the network: the visual blocks alone fill a power of two >= 256 (Flux at 4096^2: 512).  No case of
    oracle/cases.py has that shape, so here is one: Flux geometry with 256 visual blocks (1 x 128 x 256 latent + 512 text
    tokens), one head, stage level only (the CPU oracle's attention would take minutes).  Probabilities against the
    oracle; selection bit-exact from the kernel's own probabilities (n_needed, mask), as in test_select_and_rectify."""
    import numpy as np

    from oracle import rsa_oracle as O
    from rsa_b200 import geometry as G
    from rsa_b200 import ops
    t, h, w, text, top_k = 1, 128, 256, 512, 25
    nv = t * h * w
    s = nv + text
    q, k, v = O.synth_qkv(1, s, 128, "walk", 11)
    nbr = ops.gilbert_block_neighbors(t, h, w).numpy()
    geo = O.geometry_flux(s, text, top_k, 0.3)
    nq = geo.nq_blocks
    assert nq == 256 and geo.n_blocks == 260
    # oracle stages (rsa_oracle.head_forward without the attention)
    q0, k0, v0 = O.padded_inputs(q[0, 0], k[0, 0], v[0, 0], geo)
    qp, dq = O.pool_stats(q0, geo.seq + geo.gap, nq)
    kp, dk = O.pool_stats(k0, geo.kv_zero_from, nq)
    a, _ = O.block_scores(qp, dq, kp, dk, k0[nq * 128: nq * 128 + geo.text_keys])
    p_ref = O.probs_from_scores(a, nq, 128, True)
    # kernels 2, 3a, 3b
    tq, tk, tv = (torch.from_numpy(x).to(dev).to(torch.bfloat16) for x in (q, k, v))
    plan = ops.Plan(tq, tk, tv, G.flux(s, text), top_k, 0.3, torch.from_numpy(nbr), debug_dump_probs=True,
                    private_workspace=True)
    plan.pool_stats()
    plan.block_scores()
    plan.block_select()
    torch.cuda.synchronize()
    vw = plan.view()
    assert np.array_equal(vw["scores"][0].cpu().numpy(), a)
    p_gpu = vw["probs"][0].cpu().numpy()
    np.testing.assert_allclose(p_gpu, p_ref, rtol=3e-5, atol=1e-9)
    m, n = O.select_blocks(p_gpu, geo, nbr)
    assert np.array_equal(vw["n_needed"][0].cpu().numpy(), n)
    assert np.array_equal(plan.dense_mask()[0, :nq].cpu().numpy(), m)
