"""CPU tests: the oracle (oracle/) against fixtures produced by the reference itself (tests/golden/, made by
oracle/make_golden.py).  These pin the oracle; the GPU parity tests then compare the CUDA path with it."""
import json
import os

import numpy as np
import pytest

from oracle import cases as C
from oracle import gilbert_oracle as G
from oracle import rsa_oracle as O

from helpers import assert_prep_close


def _unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(bool)


def _geo(fam, nv, s, text_len, ntrue_d, top_k, p, t):
    if fam == "wan":
        return O.geometry_wan(s, top_k, p, (nv + 127) // 128 // t)
    if fam == "hunyuan":
        return O.geometry_hunyuan(s, nv + ntrue_d, top_k, p)
    if fam == "flux":
        return O.geometry_flux(s, text_len, top_k, p)
    return O.geometry_cogvideo(s, text_len, top_k, p)


@pytest.fixture(scope="module")
def gilbert_gold(gold_dir):
    with open(os.path.join(gold_dir, "gilbert.json")) as f:
        return json.load(f)


def test_gilbert_small_grids(gilbert_gold):
    for e in gilbert_gold["small"]:
        t, h, w = e["grid"]
        l2h, h2l = G.gilbert_mapping(t, h, w)
        assert l2h.tolist() == e["l2h"], e["grid"]
        assert h2l.tolist() == e["h2l"], e["grid"]
        nbr = G.gilbert_block_neighbors(t, h, w, l2h=l2h)
        assert np.array_equal(nbr, _unpack(np.array(e["nbr"], dtype=np.uint8), e["nbr_shape"])), e["grid"]
        _, h2l2 = G.gilbert_mapping(t, h, w, axis_order=("t", "h", "w"))
        assert h2l2.tolist() == e["h2l_thw"], e["grid"]
        # Appendix C invariants
        assert np.array_equal(h2l[l2h], np.arange(t * h * w))
        assert np.array_equal(nbr, nbr.T) and nbr.diagonal().all()


@pytest.mark.parametrize("name", list(C.CASES))
def test_mask_builder_against_reference(name, gold_dir):
    fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v = C.case_inputs(name)
    g = np.load(os.path.join(gold_dir, f"mask_{name}.npz"))
    mask_ref = _unpack(g["mask"], g["mask_shape"])
    nogapr_ref = _unpack(g["nogapr"], g["nogapr_shape"])
    nbr = _unpack(g["nbr"], g["nbr_shape"])
    # the oracle's own neighbour matrix equals the reference's
    assert np.array_equal(G.gilbert_block_neighbors(t, h, w), nbr)
    geo = _geo(fam, nv, s, text_len, ntrue_d, top_k, p, t)
    out_ref = g["out"].reshape(s, heads, 128)
    for hi in range(heads):
        out, st = O.head_forward(q[0, hi], k[0, hi], v[0, hi], geo, nbr, return_stages=True)
        assert st["mask"].shape == mask_ref[hi].shape
        np.testing.assert_allclose(st["probs"], g["probs"][hi], rtol=2e-5, atol=1e-7)
        # discrete decisions: identical unless the reference sits on a tie (none in these seeds)
        assert np.array_equal(st["mask"], mask_ref[hi]), f"{name} head {hi}: mask differs"
        assert np.array_equal(st["nogapr"], nogapr_ref[hi]), f"{name} head {hi}: nogapr differs"
        np.testing.assert_allclose(st["R"], g["R"][hi], rtol=0, atol=2e-6)
        np.testing.assert_allclose(st["C"], g["C"][hi], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(out, out_ref[:, hi], rtol=1e-4, atol=2e-5)
        # invariants
        assert np.all(st["mask"].sum(1) >= min(top_k, st["probs"].shape[1]))
        np.testing.assert_allclose(st["probs"].sum(1), 1.0, atol=1e-5)
        assert np.all(st["R"] > 0) and np.all(st["R"] <= 1 + 1e-6)


def test_mask_builder_head_dim_64_against_reference(gold_dir):
    """CogVideoX's real head dimension.  `mask_cog_d64.npz` is the unmodified reference (rectified_cogvideo_attn.py) run
    on CPU on 64-column fp32 tensors (oracle/make_golden.py, EXTRA_CASES; the generator also holds the reference to
    passing sm_scale = 64^-1/2 to its kernel): probabilities, mask, GAPR bytes, R, C and the output of the oracle on the
    same tensors -- the checker the GPU's head_dim 64 path is held to in test_head_dim_64."""
    name = "cog_d64"
    fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v = C.case_inputs(name)
    assert q.shape[-1] == 64
    g = np.load(os.path.join(gold_dir, f"mask_{name}.npz"))
    mask_ref = _unpack(g["mask"], g["mask_shape"])
    nogapr_ref = _unpack(g["nogapr"], g["nogapr_shape"])
    nbr = _unpack(g["nbr"], g["nbr_shape"])
    assert np.array_equal(G.gilbert_block_neighbors(t, h, w), nbr)
    geo = _geo(fam, nv, s, text_len, ntrue_d, top_k, p, t)
    out_ref = g["out"].reshape(s, heads, 64)
    for hi in range(heads):
        out, st = O.head_forward(q[0, hi], k[0, hi], v[0, hi], geo, nbr, return_stages=True)
        np.testing.assert_allclose(st["probs"], g["probs"][hi], rtol=2e-5, atol=1e-7)
        assert np.array_equal(st["mask"], mask_ref[hi]), f"head {hi}: mask differs"
        assert np.array_equal(st["nogapr"], nogapr_ref[hi]), f"head {hi}: nogapr differs"
        np.testing.assert_allclose(st["R"], g["R"][hi], rtol=0, atol=2e-6)
        np.testing.assert_allclose(st["C"], g["C"][hi], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(out, out_ref[:, hi], rtol=1e-4, atol=2e-5)
        assert float(st["R"].min()) < 0.9                       # a case in which the rectification matters


def _mid_geo(name):
    fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v = C.mid_case_inputs(name)
    return fam, (t, h, w), nv, heads, _geo(fam, nv, s, text_len, ntrue_d, top_k, p, t), q, k, v


@pytest.mark.parametrize("name", list(C.MID_CASES))
def test_mask_builder_mid_sizes_against_reference(name, gold_dir):
    """300 ... 1101 sortable entries per query block (the selection kernel's 2-, 4- and 8-entries-per-thread sorting
    networks, the text aggregate inserted beside exactly 512 visual blocks, the cumulative threshold beyond top_k on iid
    inputs, exact four-way ties across the cut).  `mask_<name>.npz` = the unmodified reference's
    _build_block_index_with_importance_optimized + R + C on fp32 CPU tensors (oracle/make_golden.py `mid`)."""
    fam, (t, h, w), nv, heads, geo, q, k, v = _mid_geo(name)
    g = np.load(os.path.join(gold_dir, f"mask_{name}.npz"))
    mask_ref = _unpack(g["mask"], g["mask_shape"])
    nogapr_ref = _unpack(g["nogapr"], g["nogapr_shape"])
    nbr = _unpack(g["nbr"], g["nbr_shape"])
    assert np.array_equal(G.gilbert_block_neighbors(t, h, w), nbr)
    ties = C.MID_CASES[name][7] == "ties"
    for hi in range(heads):
        st, _, _, _ = O.mask_stages(q[0, hi], k[0, hi], v[0, hi], geo, nbr)
        assert st["mask"].shape == mask_ref[hi].shape
        np.testing.assert_allclose(st["probs"][g["prob_rows"]], g["probs"][hi], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(st["probs"].sum(1, dtype=np.float64), g["prob_sum"][hi], atol=2e-6)
        assert np.array_equal(st["nogapr"], nogapr_ref[hi]), f"{name} head {hi}: nogapr differs"
        if not ties:
            assert np.array_equal(st["mask"], mask_ref[hi]), f"{name} head {hi}: mask differs"
            np.testing.assert_allclose(st["R"], g["R"][hi], rtol=0, atol=2e-6)
            np.testing.assert_allclose(st["C"], g["C"][hi], rtol=1e-4, atol=2e-6)
        else:
            # torch.sort without stable=True (wan21 :220) leaves the order inside a tie group open, so the reference may
            # keep other members of the group the cut falls into: same count per row, same selected probabilities --
            # and this oracle keeps the LOWEST indices of that group (the documented tie-break)
            p = st["probs"]
            forced = np.zeros_like(st["mask"])
            forced[:, : nbr.shape[1]] |= nbr[: forced.shape[0]]
            forced[: geo.first_frame_blocks, : geo.first_frame_blocks] = True
            order = np.argsort(-p, axis=1, kind="stable")
            for i in range(p.shape[0]):
                n = st["n_needed"][i]
                cut = p[i, order[i, n - 1]]
                above, grp = p[i] > cut, np.nonzero(p[i] == cut)[0]
                need = n - int(above.sum())
                ref_row = mask_ref[hi][i]
                assert ref_row[above].all(), (name, i)                                # everything above the cut is kept
                assert not (ref_row & (p[i] < cut) & ~forced[i]).any(), (name, i)     # nothing below it, unions aside
                took = int(ref_row[grp].sum())
                assert need <= took <= need + int(forced[i, grp].sum()), (name, i)    # `need` members of the tie group
                mine = st["mask"][i] & ~forced[i]
                kept = grp[mine[grp] | forced[i, grp]]
                # this oracle's ranking keeps the LOWEST indices of the group (the documented tie-break)
                pure = np.zeros(p.shape[1], dtype=bool)
                pure[order[i, :n]] = True
                assert np.array_equal(grp[pure[grp]], grp[:need]), (name, i)
                assert np.array_equal(st["mask"][i], pure | forced[i]), (name, i)
            assert (p[:, 0] == p[:, 1]).all() and (st["n_needed"] % 4 != 0).any()   # ties exist and a cut splits one
            # R sums `part = mask | nogapr`: another member of the tie group may or may not be a GAPR entry already, so R
            # can differ from the reference's by the group's probabilities
            cutv = np.take_along_axis(p, order, 1)[np.arange(p.shape[0]), st["n_needed"] - 1]
            assert np.all(np.abs(st["R"] - g["R"][hi]) <= 4 * cutv + 2e-6)
        assert np.all(st["mask"].sum(1) >= min(geo.top_k, st["probs"].shape[1]))
        assert np.all(st["R"] > 0) and np.all(st["R"] <= 1 + 1e-6)


def test_triton_kernel_semantics_fp16(gold_dir):
    """The literal reference kernel (TRITON_INTERPRET, fp16) vs the oracle's dense masked restatement."""
    import torch

    g = np.load(os.path.join(gold_dir, "kernel_fp16.npz"))
    rnd = lambda x: x.half().float()
    seqlen = int(g["seqlen"])
    for hi in range(g["q"].shape[1]):
        o = O.masked_attention(g["q"][0, hi].astype(np.float32), g["k"][0, hi].astype(np.float32),
                               g["v"][0, hi].astype(np.float32), g["mask"][0, hi], seqlen, seqlen, emulate=rnd)
        ref = g["out"][0, hi].astype(np.float32)
        assert np.abs(o - ref[:seqlen]).max() <= 1e-3      # 1-2 fp16 ulp at |o| <= 1
        assert np.all(ref[seqlen:] == 0)                      # padded query rows are never stored (wan21 :105)
        # and the exact (unrounded) restatement is within the same distance
        o32 = O.masked_attention(g["q"][0, hi].astype(np.float32), g["k"][0, hi].astype(np.float32),
                                 g["v"][0, hi].astype(np.float32), g["mask"][0, hi], seqlen, seqlen)
        assert np.abs(o32 - ref[:seqlen]).max() <= 3e-3


def test_dense_limit():
    """top_k >= NB -> mask all true -> R = 1, C = 0 -> output equals dense attention (Appendix C)."""
    import torch

    q, k, v = O.synth_qkv(1, 512, 128, "walk", 21)
    geo = O.geometry_wan(512, 99, 0.3, 0)
    out, st = O.head_forward(q[0, 0], k[0, 0], v[0, 0], geo, None, return_stages=True)
    assert st["mask"].all() and np.allclose(st["R"], 1.0, atol=1e-6) and np.abs(st["C"]).max() == 0
    ref = torch.nn.functional.scaled_dot_product_attention(
        torch.from_numpy(q[0, 0])[None], torch.from_numpy(k[0, 0])[None], torch.from_numpy(v[0, 0])[None])[0]
    np.testing.assert_allclose(out, ref.numpy(), atol=2e-5)


def test_select_block_num_truncation():
    # int((1-0.8)*900) = 179, int((1-0.9)*512) = 51, int(0.25*591) = 147
    assert O.select_block_num(0.8, 900) == 179
    assert O.select_block_num(0.9, 512) == 51
    assert O.select_block_num(0.75, 591) == 147


def test_selection_tie_break_and_threshold():
    geo = O.geometry_wan(4 * 128, 1, 0.5, 0)
    p = np.array([[0.25, 0.25, 0.25, 0.25], [0.1, 0.6, 0.2, 0.1], [0.5, 0.5, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]],
                 dtype=np.float32)
    m, n = O.select_blocks(p, geo, None)
    # row0: cumsum .25,.5,.75,1 -> two entries <= .5 -> n = 3, ties resolved towards the lower index
    assert n.tolist() == [3, 1, 2, 1]
    assert m.tolist() == [[True, True, True, False], [False, True, False, False], [True, True, False, False],
                          [False, False, False, True]]


def test_prep_oracle_against_torch_sequence(gold_dir):
    """oracle/prep_oracle.py (head split, diffusers RMSNorm, diffusers apply_rotary_emb, restated) against the literal
    PyTorch op sequence run in bf16 (tests/golden/prep_hunyuan.npz).  The reduction order of mean(x^2) and rsqrt differ
    by fp32 ulps, so a bf16 rounding may flip: every element within 1 bf16 ulp, >= 99.9 % identical."""
    import torch

    from oracle import make_golden as MG
    from oracle import prep_oracle as P
    src, wq, wk, cos, sin, n_rope = MG.prep_inputs()
    g = np.load(os.path.join(gold_dir, "prep_hunyuan.npz"))
    for name, x, w, nr in (("q", src[0], wq, n_rope), ("k", src[1], wk, n_rope), ("v", src[2], None, 0)):
        got = P.prep(x.float().numpy(), 2, None if w is None else w.float().numpy(), 1e-6, cos.numpy(), sin.numpy(), nr)
        ref = torch.from_numpy(g[name]).view(torch.bfloat16).float().numpy()
        assert_prep_close(got, ref, name)
        if name == "v":
            assert np.array_equal(got, ref)


def wan_prep_inputs():
    """Inputs of tests/golden/prep_wan.npz (oracle/make_golden.py::make_prep): the Hunyuan sources, inner-dim weights,
    rotary angles for every row (as fp32 cos/sin tables, the form kernel 0 takes)."""
    import torch

    from oracle import make_golden as MG
    src, *_ = MG.prep_inputs()
    g = torch.Generator().manual_seed(32)
    rows = src[0].shape[1]
    wq = (1 + 0.1 * torch.randn(256, generator=g)).to(torch.bfloat16)
    wk = (1 + 0.1 * torch.randn(256, generator=g)).to(torch.bfloat16)
    ang = torch.outer(torch.arange(rows, dtype=torch.float64),
                      1.0 / (256.0 ** (torch.arange(0, 128, 2, dtype=torch.float64) / 128)))
    cos = ang.cos().float().repeat_interleave(2, dim=1)
    sin = ang.sin().float().repeat_interleave(2, dim=1)
    return src, wq, wk, cos, sin


def test_prep_oracle_wan_form(gold_dir):
    """The Wan form (RMSNorm across heads, complex rotary embedding in float64) against the literal PyTorch sequence:
    the oracle does the rotation in fp32 like the kernel; same bar as the per-head form."""
    import torch

    from oracle import prep_oracle as P
    src, wq, wk, cos, sin = wan_prep_inputs()
    g = np.load(os.path.join(gold_dir, "prep_wan.npz"))
    for name, x, w in (("q", src[0], wq), ("k", src[1], wk)):
        got = P.prep(x.float().numpy(), 2, w.float().numpy(), 1e-6, cos.numpy(), sin.numpy(), x.shape[1])
        ref = torch.from_numpy(g[name]).view(torch.bfloat16).float().numpy()
        assert_prep_close(got, ref, name)


def test_prep_oracle_cogvideo_form(gold_dir):
    """The CogVideoX form (LayerNorm over head_dim with bias, rotary embedding on the video tokens) against
    torch.nn.functional.layer_norm + the apply_rotary_emb expression in bf16."""
    import torch

    from oracle import make_golden as MG
    from oracle import prep_oracle as P
    src, _, _, cos, sin, n_rope = MG.prep_inputs()
    cw, cb = MG.cog_prep_params()
    g = np.load(os.path.join(gold_dir, "prep_cog.npz"))
    for i, name in enumerate(("q", "k")):
        got = P.prep(src[i].float().numpy(), 2, cw[i].float().numpy(), 1e-6, cos.numpy(), sin.numpy(), n_rope,
                     bias=cb[i].float().numpy())
        ref = torch.from_numpy(g[name]).view(torch.bfloat16).float().numpy()
        assert_prep_close(got, ref, name)
