"""GPU parity at BASELINE.json's FULL sizes (2 heads each, so a case takes about a second): the CPU oracle cannot reach
these shapes, so the checks are the size-independent ones --

  * the invariants of SURVEY.md Appendix C on the kernels' own outputs (kept counts >= top_k, mask contains the Gilbert
    neighbours / the first-frame square / the text blocks, 0 < R <= 1, text query tiles dense with R = 1 and C = 0,
    kept lists strictly ascending);
  * sampled query tiles (first, last visual -- the ragged one at 129 frames --, random ones, a text tile) recomputed in
    fp32 PyTorch from the kernel's kept list, R and C, with the reference kernel's rounding of the pre-scaled query on
    visual tiles: output max-abs-err <= 2e-2 and cosine >= 0.999;
  * where the visual segment is block-aligned, the whole output against the tests' independent mma.sync kernel
    (tests/xcheck: its own library, not part of the product).
"""
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

pytestmark = pytest.mark.gpu

ATOL_OUT, COS_OUT = 2e-2, 0.999
HEADS = 2


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


def _rows_of_block(j, nq, nv, seq):
    """Memory rows of block j of the padded layout (rsa_common.cuh RowMap)."""
    if j < nq:
        return 128 * j, min(128 * j + 128, nv)
    first = nv + 128 * (j - nq)
    return first, min(first + 128, seq)


@pytest.mark.parametrize("name", ["c3b", "c3a", "c4", "c5", "c2"])
def test_full_size_invariants_and_sampled_tiles(dev, name):
    from rsa_b200 import ops
    bench = _bench()
    wp = bench.workload_params(name)
    s, nv = wp["s"], wp["nv"]
    t, h, w = wp["grid"]
    nbr = ops.gilbert_block_neighbors(t, h, w)
    q, k, v = bench.synth_heads_device(HEADS, 0, s, "walk", dev)
    geo = bench.product_geometry(wp)
    plan = ops.Plan(q, k, v, geo, wp["top_k"], bench.P_REMAIN, nbr)
    out = plan.run().clone()                                  # [1, S, H, D]
    torch.cuda.synchronize()
    vw = plan.view()
    nq, nb, gap = geo.nq_blocks, geo.n_blocks, geo.gap
    joint = geo.family == 1
    nv_mem = nv if joint else s                               # rows of the visual segment in memory
    cnt = vw["kept_cnt"].cpu().numpy()
    idx = vw["kept_idx"].cpu().numpy().astype(np.int64) & 0xFFFF
    R = vw["R"].cpu().numpy()
    C = vw["C"]
    mask = plan.dense_mask()[:, :nq].cpu().numpy()            # [H, NQ, NB]
    kvb = (geo.kv_len + 127) // 128

    # ---- invariants
    n_ent = nq + (1 if joint else 0)
    assert np.all(cnt[:, :nq] >= min(wp["top_k"], n_ent) - 1)     # the text aggregate may take one of the top_k slots
    assert np.all(mask.sum(2) >= min(wp["top_k"], n_ent))
    nb_np = nbr.numpy()
    assert np.all(mask[:, : nb_np.shape[0], : nb_np.shape[1]][:, nb_np[:nq, :nq]])
    if joint:
        assert np.all(mask[:, :, nq: geo.text_end_block])
        assert np.all(cnt[:, nq:] == kvb) and np.all(R[:, nq:] == 1.0) and float(C[:, nq:].abs().max()) == 0.0
    else:
        f = wp["ffb_blocks"]
        assert np.all(mask[:, :f, :f])
    assert np.all(R[:, :nq] > 0) and np.all(R[:, :nq] <= 1 + 1e-5)
    for hi in range(HEADS):
        for i in (0, nq // 3, nq - 1):
            li = idx[hi, i, : cnt[hi, i]]
            assert np.all(np.diff(li) > 0) and li.max() < kvb
            assert np.array_equal(li, np.nonzero(mask[hi, i, :kvb])[0])

    # ---- sampled query tiles against fp32 PyTorch on the kernel's own list, R, C
    rng = np.random.default_rng(5)
    tiles = sorted({0, nq - 1, *rng.integers(1, nq - 1, size=4).tolist(), *([nq] if joint else [])})
    scale = 128 ** -0.5
    for hi in range(HEADS):
        for i in tiles:
            r0, r1 = _rows_of_block(i, nq, nv_mem, s)
            if i >= nq:                                       # text tile: rows beyond the valid text are zero-filled
                r1 = min(r1, nv_mem + geo.text_q_valid)
            qi = q[0, hi, r0:r1].float()
            keys = []
            for j in idx[hi, i, : cnt[hi, i]]:
                a0, a1 = _rows_of_block(int(j), nq, nv_mem, s)
                v0 = 128 * int(j)                             # padded-layout position of the block's first key
                a1 = min(a1, a0 + max(0, geo.kv_len - v0))    # keys >= kv_len are never attended
                keys.append(torch.arange(a0, a1, device=dev))
            keys = torch.cat(keys)
            if i < nq:      # visual tile: the reference kernel's q~ = bf16(q * sm_scale * log2 e) and exp2 (wan21 :61-62)
                sc = (qi * (scale * 1.44269504)).to(torch.bfloat16).float() @ k[0, hi, keys].float().T
                p = torch.exp2(sc - sc.max(dim=-1, keepdim=True).values)
                p = p / p.sum(dim=-1, keepdim=True)
            else:           # text tile: flash-attn's arithmetic (fp32 scores scaled)
                p = torch.softmax((qi @ k[0, hi, keys].float().T) * scale, dim=-1)
            ref = p @ v[0, hi, keys].float()
            ref = ref * float(R[hi, i]) + C[hi, i][None, :]
            got = out[0, r0:r1, hi].float()
            err = float((got - ref).abs().max())
            cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
            assert err <= ATOL_OUT and cos >= COS_OUT, f"{name} head {hi} tile {i}: max-abs {err:.4f}, cos {cos:.5f}"
    if joint and geo.text_q_valid < s - nv_mem:               # padded text query rows are written as zeros
        assert float(out[0, nv_mem + geo.text_q_valid:].abs().max()) == 0.0

    # ---- whole output against the mma.sync kernel (block-aligned visual segments only)
    if gap == 0:
        import xcheck
        ref_all = xcheck.sparse_attention(plan)
        torch.cuda.synchronize()
        d = (out.float() - ref_all.float()).abs()
        assert float(d.max()) <= ATOL_OUT, f"{name}: tcgen05 vs mma.sync max-abs {float(d.max()):.4f}"
        cos = float(torch.nn.functional.cosine_similarity(out.float().flatten(), ref_all.float().flatten(), dim=0))
        assert cos >= 0.9999

