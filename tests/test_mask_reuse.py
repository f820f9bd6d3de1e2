"""Mask re-use across calls: rsa_rectified_attention_reuse / ops.MaskCache / the processors'
`mask_refresh_interval`.  Not a reference feature (the reference rebuilds its mask on every call, hunyuan :334-346),
so the checker is the oracle with the earlier call's selection carried over (`head_forward(keep_mask=, keep_rc=)`)
and the bars are the same as for the plain call: bit-identical where nothing changed, R within 2e-6, C within fp32
tolerance, output max-abs-err <= 2e-2 and cosine >= 0.999."""
import numpy as np
import pytest
import torch

from helpers import cos_sim, load_case, product_geometry
from oracle import rsa_oracle as O

ATOL_OUT, COS_OUT = 2e-2, 0.999


# ------------------------------------------------------------------------------------------------- host logic (CPU)
def test_mask_cache_schedule():
    from rsa_b200 import native as N
    from rsa_b200 import ops
    dev = torch.device("cpu")

    def call(c, key="k", nbytes=64):
        mode, ws = c.next_mode(key, dev, nbytes)
        c.ran(mode)
        return mode, ws

    c = ops.MaskCache(refresh_every=3, keep="lists")
    assert [call(c)[0] for _ in range(7)] == [N.MASK_BUILD, N.MASK_KEEP_LISTS, N.MASK_KEEP_LISTS] * 2 + [N.MASK_BUILD]
    ws = call(c)[1]
    assert call(c)[1] is ws                                          # the workspace IS the cache: it must persist
    # a plan that is built but never run must not advance the schedule or validate the workspace
    d = ops.MaskCache(refresh_every=2)
    assert [d.next_mode("k", dev, 64)[0] for _ in range(3)] == [N.MASK_BUILD] * 3
    assert [call(d)[0] for _ in range(3)] == [N.MASK_BUILD, N.MASK_KEEP_LISTS, N.MASK_BUILD]
    assert call(c, "other geometry")[0] == N.MASK_BUILD              # a different call shape invalidates it
    assert call(c, "other geometry")[0] == N.MASK_KEEP_LISTS
    assert call(c, "other geometry", 4096)[0] == N.MASK_BUILD        # a workspace that has to grow is empty
    c.reset()
    assert call(c, "other geometry")[0] == N.MASK_BUILD
    a = ops.MaskCache(refresh_every=2, keep="all")
    assert [call(a, 1, 8)[0] for _ in range(4)] == [N.MASK_BUILD, N.MASK_KEEP_ALL] * 2
    one = ops.MaskCache(refresh_every=1)
    assert [call(one, 1, 8)[0] for _ in range(3)] == [N.MASK_BUILD] * 3
    with pytest.raises(ValueError):
        ops.MaskCache(keep="nothing")


def test_processor_cache_follows_its_attributes():
    from rectified_spaattn.rectified_wan21_attn import RectifiedWanT2VSpaAttnProcessor2_0 as P
    p = P("sparse", 4, None, 0.3)
    assert p._mask_cache() is None                                   # default: the reference's behaviour
    p.mask_refresh_interval = 4
    c = p._mask_cache()
    assert c is p._mask_cache() and c.refresh_every == 4 and c.keep == "lists"
    p.mask_keep = "all"
    assert p._mask_cache() is not c and p._mask_cache().keep == "all"
    c2 = p._mask_cache()
    c2.calls, c2._key, c2.valid = 3, "x", True
    p.current_step = p.steps_per_cycle - 1
    p._tick()                                                        # a new generation starts: selection forgotten
    assert p.current_step == 0 and c2.calls == 0 and c2._key is None and not c2.valid


def test_processor_cache_is_per_guidance_branch():
    """The Wan pipelines call a processor twice per denoising step (conditional, unconditional: the reference's counter
    wraps at 50 * 2, rectified_wan21_attn.py:501): each branch has its own cache, so a refresh interval counts denoising
    steps and the unconditional pass never runs on a selection built from the conditional pass.  HunyuanVideo (one call
    per step, wrap at 50, rectified_hunyuan_attn.py:542) has one."""
    from rectified_spaattn.rectified_hunyuan_attn import RectifiedHunyuanVideoSpaAttnProcessor2_0 as H
    from rectified_spaattn.rectified_wan21_attn import RectifiedWanT2VSpaAttnProcessor2_0 as W
    w = W("sparse", 4, None, 0.3)
    w.mask_refresh_interval = 2
    seen = []
    for _ in range(6):
        seen.append(w._mask_cache())
        w._tick()
    assert seen[0] is seen[2] is seen[4] and seen[1] is seen[3] is seen[5] and seen[0] is not seen[1]
    h = H("sparse", 4, None, 0.3)
    h.mask_refresh_interval = 2
    a = h._mask_cache()
    h._tick()
    assert h._mask_cache() is a


# ----------------------------------------------------------------------------------------------------- GPU parity
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _second_inputs(case, seed):
    """A later call of the same layer: the first call's tensors plus a perturbation of half their spread."""
    g = np.random.default_rng(seed)
    out = []
    for n in ("q", "k", "v"):
        x = case[n] + 0.5 * g.standard_normal(case[n].shape).astype(np.float32)
        out.append(torch.from_numpy(x).to(torch.bfloat16).float().numpy())
    return out


def _plan(case, dev, tensors, cache):
    from rsa_b200 import ops
    q, k, v = (torch.from_numpy(x).to(dev).to(torch.bfloat16) for x in tensors)
    geo = product_geometry(case["fam"], case["nv"], case["s"], case["text_len"], case["ntrue_d"], case["grid"][0])
    return ops.Plan(q, k, v, geo, case["top_k"], case["p"], torch.from_numpy(case["nbr"]), mask_cache=cache)


@pytest.mark.gpu
@pytest.mark.parametrize("keep", ["lists", "all"])
@pytest.mark.parametrize("name", ["wan_c1", "wan_ragged", "hunyuan_small", "flux_small", "cog_small", "hunyuan_ragged"])
def test_reuse_on_unchanged_inputs_is_bit_identical(dev, name, keep):
    from rsa_b200 import native as N
    from rsa_b200 import ops
    case = load_case(name)
    first = (case["q"], case["k"], case["v"])
    cache = ops.MaskCache(refresh_every=2, keep=keep)
    p0 = _plan(case, dev, first, cache)
    assert p0.mask_mode == N.MASK_BUILD
    a = p0.run().clone()
    p1 = _plan(case, dev, first, cache)
    assert p1.mask_mode == (N.MASK_KEEP_LISTS if keep == "lists" else N.MASK_KEEP_ALL) and p1.ws is p0.ws
    b = p1.run()
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    ref = O.forward(*first, case["ogeo"], case["nbr"], q_dtype="bf16").reshape(a.shape)
    assert np.abs(b.float().cpu().numpy() - ref).max() <= ATOL_OUT


@pytest.mark.gpu
@pytest.mark.parametrize("keep", ["lists", "all"])
@pytest.mark.parametrize("name", ["wan_c1", "hunyuan_small", "flux_small", "cog_small", "hunyuan_mid", "hunyuan_ragged"])
def test_reuse_on_new_inputs_matches_oracle_with_carried_selection(dev, name, keep):
    from rsa_b200 import ops
    case = load_case(name)
    geo, nbr = case["ogeo"], case["nbr"]
    nq = geo.nq_blocks
    first = (case["q"], case["k"], case["v"])
    second = _second_inputs(case, 17)
    cache = ops.MaskCache(refresh_every=2, keep=keep)
    p0 = _plan(case, dev, first, cache)
    p0.run()
    torch.cuda.synchronize()
    vw = p0.view()
    lists0 = (vw["kept_idx"].clone(), vw["kept_cnt"].clone(), vw["mask_bits"].clone())
    r0, c0 = vw["R"].cpu().numpy().copy(), vw["C"].cpu().numpy().copy()
    mask0 = p0.dense_mask().cpu().numpy()
    p1 = _plan(case, dev, second, cache)
    out = p1.run().float().cpu().numpy()[0]                           # [S, H, D]
    vw = p1.view()
    # the selection of the first call stands
    assert all(torch.equal(a, b) for a, b in zip(lists0, (vw["kept_idx"], vw["kept_cnt"], vw["mask_bits"])))
    for hi in range(case["heads"]):
        q, k, v = (x[0, hi] for x in second)
        keep_rc = None if keep == "lists" else (r0[hi, :nq], c0[hi, :nq])
        ref, st = O.head_forward(q, k, v, geo, nbr, return_stages=True, keep_mask=mask0[hi, :nq], keep_rc=keep_rc,
                                 q_dtype="bf16")
        if keep == "lists":     # R, C recomputed from the second call's P, GAPR bytes and pooled V
            np.testing.assert_allclose(vw["R"][hi, :nq].cpu().numpy(), st["R"], rtol=0, atol=3e-6)
            np.testing.assert_allclose(vw["C"][hi, :nq].cpu().numpy(), st["C"], rtol=1e-4, atol=3e-6)
            assert not np.array_equal(st["R"], r0[hi, :nq])           # and they did change
        else:
            assert np.array_equal(vw["R"][hi].cpu().numpy(), r0[hi]) and np.array_equal(vw["C"][hi].cpu().numpy(), c0[hi])
        got = out[:, hi]
        assert np.abs(got - ref).max() <= ATOL_OUT, f"{name} head {hi}: {np.abs(got - ref).max()}"
        assert cos_sim(got, ref) >= COS_OUT
    # the third call rebuilds: identical to a call without any cache
    p2 = _plan(case, dev, second, cache)
    from rsa_b200 import native as N
    assert p2.mask_mode == N.MASK_BUILD
    a = p2.run().clone()
    b = _plan(case, dev, second, None).run()
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))


@pytest.mark.gpu
def test_reuse_through_the_family_entry_point_and_bad_mode(dev):
    import ctypes as C
    from rectified_spaattn.rectified_wan21_attn import rectified_block_sparse_attention
    from rsa_b200 import native as N
    from rsa_b200 import ops
    case = load_case("wan_c1")
    q, k, v = (torch.from_numpy(case[n]).to(dev).to(torch.bfloat16) for n in ("q", "k", "v"))
    nbr = torch.from_numpy(case["nbr"])
    ffb = (case["nv"] + 127) // 128 // case["grid"][0]
    kw = dict(attn_mask=None, top_k=case["top_k"], block_neighbor_list=nbr, p_remain_rates=case["p"],
              first_frame_blocks=ffb)
    plain = rectified_block_sparse_attention(q, k, v, **kw)
    cache = ops.MaskCache(refresh_every=3)
    outs = [rectified_block_sparse_attention(q, k, v, mask_cache=cache, **kw).clone() for _ in range(3)]
    torch.cuda.synchronize()
    assert cache.calls == 3 and cache.valid and all(torch.equal(o, plain) for o in outs)
    plan = ops.Plan(q, k, v, product_geometry("wan", case["nv"], case["s"], 0, 0, case["grid"][0]), case["top_k"],
                    case["p"], nbr, private_workspace=True)
    rc = N.lib().rsa_rectified_attention_reuse(C.byref(plan.desc), q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                               plan.out.data_ptr(), plan.ws.data_ptr(), plan.ws_bytes, 7, 0, None)
    assert rc == -1 and b"mask_mode" in N.lib().rsa_last_error_string()
    with pytest.raises(RuntimeError):                                # host-buffer call: no lasting workspace
        ops.rectified_attention(q.cpu().pin_memory(), k.cpu().pin_memory(), v.cpu().pin_memory(), plan_geo(case), 2, 0.3,
                                None, mask_cache=cache)


def plan_geo(case):
    return product_geometry("wan", case["nv"], case["s"], 0, 0, case["grid"][0])
