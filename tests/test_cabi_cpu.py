"""CPU tests of the C ABI: the library loads, exports every symbol include/rsa.h declares, the pure-CPU host
geometry matches the reference fixtures, and descriptor validation fails loudly.  No GPU compute is called."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest
import torch

from rsa_b200 import geometry as G
from rsa_b200 import native as N
from rsa_b200 import ops

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "rsa.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rsa_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = N.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in rsa.h but not exported"
    assert declared == set(N.EXPORTS)
    assert lib.rsa_version() == N.ABI_VERSION == int(re.search(r"#define RSA_VERSION (\d+)", hdr).group(1))


def test_struct_layout_matches_header():
    # sizeof via a tiny C program compiled against the header
    import subprocess
    import tempfile
    src = '#include "rsa.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n", sizeof(rsa_attn_desc), sizeof(rsa_ws_view));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "s.c")
        open(p, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), p, "-o", exe])
        a, b = map(int, subprocess.check_output([exe]).split())
    assert a == C.sizeof(N.AttnDesc) and b == C.sizeof(N.WsView)


def test_integration_md_stub_matches_the_library():
    """INTEGRATION.md section 2 shows the ctypes struct a maintainer of the reference would write.  A stale copy (a
    field missing at the end) makes the library read past the caller's buffer: the documented class is extracted from
    the document, executed, and held to the library's own sizeof and to every field offset of rsa_b200/native.py."""
    import re
    doc = open(os.path.join(REPO, "INTEGRATION.md"), encoding="utf-8").read()
    m = re.search(r"^class _Desc\(C\.Structure\):.*?\n(?=\n)", doc, re.S | re.M)
    assert m, "INTEGRATION.md no longer shows class _Desc"
    ns = {"C": C}
    exec(m.group(0), ns)
    doc_desc = ns["_Desc"]
    L = N.lib()
    assert C.sizeof(doc_desc) == L.rsa_attn_desc_size() == C.sizeof(N.AttnDesc)
    assert [(n, getattr(doc_desc, n).offset, getattr(doc_desc, n).size) for n, _ in doc_desc._fields_] == \
           [(n, getattr(N.AttnDesc, n).offset, getattr(N.AttnDesc, n).size) for n, _ in N.AttnDesc._fields_]
    assert "rsa_attn_desc_size() == C.sizeof(_Desc)" in doc          # and the stub checks itself at import
    assert L.rsa_prep_desc_size() == C.sizeof(N.PrepDesc) and L.rsa_peer_route_size() == C.sizeof(N.PeerRoute)


def test_too_many_kv_blocks_is_refused():
    """Kernel 3b keeps a row's kept-block bitmask in RSA_MAX_ENTRIES/32 + 1 shared-memory words: a JOINT descriptor whose
    visual blocks pass the entry bound but whose text blocks push n_blocks beyond it must fail validation."""
    from rsa_b200 import geometry as G
    d = N.AttnDesc()
    nq, text = 2000, 200 * 128
    s = nq * 128 + text
    geo = G.BlockGeometry(1, s, nq + 200, nq, 128, s, s, nq + 200, text)
    ops._fill_desc(d, (1, 1, s, 128), [(s * 128, s * 128, 128)] * 4, geo, 10, 0.3, None)
    assert N.lib().rsa_attn_workspace_bytes(C.byref(d)) == 0
    assert b"KV blocks" in N.lib().rsa_last_error_string()
    geo_ok = G.BlockGeometry(1, nq * 128 + 1024, nq + 8, nq, 128, nq * 128 + 1024, nq * 128 + 1024, nq + 8, 1024)
    ops._fill_desc(d, (1, 1, geo_ok.seq, 128), [(geo_ok.seq * 128, geo_ok.seq * 128, 128)] * 4, geo_ok, 10, 0.3, None)
    assert N.lib().rsa_attn_workspace_bytes(C.byref(d)) > 0


@pytest.fixture(scope="module")
def gilbert_gold(gold_dir):
    with open(os.path.join(gold_dir, "gilbert.json")) as f:
        return json.load(f)


def test_gilbert_cabi_small(gilbert_gold):
    for e in gilbert_gold["small"]:
        t, h, w = e["grid"]
        l2h, h2l = ops.gilbert_mapping(t, h, w)
        assert l2h.tolist() == e["l2h"] and h2l.tolist() == e["h2l"], e["grid"]
        nbr = ops.gilbert_block_neighbors(t, h, w)
        n = int(np.prod(e["nbr_shape"]))
        ref = np.unpackbits(np.array(e["nbr"], dtype=np.uint8))[:n].reshape(e["nbr_shape"]).astype(bool)
        assert np.array_equal(nbr.numpy(), ref), e["grid"]
        _, h2l2 = ops.gilbert_mapping(t, h, w, ("t", "h", "w"))
        assert h2l2.tolist() == e["h2l_thw"]


def test_gilbert_cabi_baseline_grids(gilbert_gold):
    sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    for e in gilbert_gold["big"]:
        t, h, w = e["grid"]
        l2h, h2l = ops.gilbert_mapping(t, h, w)
        nbr = ops.gilbert_block_neighbors(t, h, w)
        assert sha(h2l.numpy()) == e["sha_h2l"] and sha(l2h.numpy()) == e["sha_l2h"], e["grid"]
        assert sha(nbr.numpy().astype(np.uint8)) == e["sha_nbr"] and int(nbr.sum()) == e["nbr_nnz"], e["grid"]
        assert h2l[:6].tolist() == e["h2l_head"]


def test_jenga_gilbert_mirror_signatures(gilbert_gold):
    from utils import jenga_gilbert as J
    e = gilbert_gold["small"][0]
    t, h, w = e["grid"]
    l2h, h2l = J.gilbert_mapping(t, h, w)
    assert isinstance(l2h, list) and l2h == e["l2h"] and h2l == e["h2l"]
    nbr = J.gilbert_block_neighbor_mapping(t, h, w)
    assert nbr.dtype == torch.bool and tuple(nbr.shape) == tuple(e["nbr_shape"])
    assert J.gilbert_xyz2d(3, 2, 1, w, h, t, ("w", "h", "t")) == e["l2h"][(1 * h + 2) * w + 3]


def test_gilbert_bad_arguments():
    lib = N.lib()
    buf = (C.c_int64 * 8)()
    assert lib.rsa_gilbert_map(0, 2, 2, b"wht", buf, buf) == -1
    assert lib.rsa_gilbert_map(2, 2, 2, b"wxt", buf, buf) == -1
    assert b"axis_order" in lib.rsa_last_error_string()


def _desc(geo, b=1, h=2, top_k=2, p=0.3):
    d = N.AttnDesc()
    d.batch, d.heads, d.seq, d.head_dim = b, h, geo.seq, 128
    for name in ("q_stride", "k_stride", "v_stride", "o_stride"):
        arr = getattr(d, name)
        arr[0], arr[1], arr[2] = h * geo.seq * 128, geo.seq * 128, 128
    d.family, d.n_blocks, d.nq_blocks, d.text_keys = geo.family, geo.n_blocks, geo.nq_blocks, geo.text_keys
    d.kv_len, d.kv_zero_from, d.text_end_block = geo.kv_len, geo.kv_zero_from, geo.text_end_block
    d.text_q_valid, d.top_k, d.p_remain, d.first_frame_blocks = geo.text_q_valid, top_k, p, geo.first_frame_blocks
    d.vis_len = geo.vis_len
    return d


def test_descriptor_validation_and_workspace_size():
    lib = N.lib()
    d = _desc(G.hunyuan(1280, 1224))
    n = lib.rsa_attn_workspace_bytes(C.byref(d))
    assert n > 0 and n % 256 == 0
    d.head_dim = 64                                         # CogVideoX: same workspace (statistics are 128 wide)
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) == n
    d.head_dim = 96
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) == 0
    assert b"head_dim" in lib.rsa_last_error_string()
    d = _desc(G.wan(1000))
    d.n_blocks = 7
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) == 0
    d = _desc(G.wan(1000))
    d.dtype = N.DTYPE_F16                                   # fp16 tensors: same workspace (fp32 statistics)
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) == lib.rsa_attn_workspace_bytes(C.byref(_desc(G.wan(1000))))
    d.dtype = 7
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) == 0 and b"dtype" in lib.rsa_last_error_string()
    d = _desc(G.wan(1000))
    d.head_dim = 32                                         # 128 and 64 (CogVideoX) are built, nothing else
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) == 0 and b"head_dim" in lib.rsa_last_error_string()
    # stage calls refuse a null / short workspace before touching the device
    d = _desc(G.wan(1024))
    assert lib.rsa_block_scores(C.byref(d), None, 0, None) == -4


def test_product_geometry_matches_oracle():
    from oracle import rsa_oracle as O
    pairs = [(G.wan(32760, 12), O.geometry_wan(32760, 64, 0.3, 12)),
             (G.hunyuan(115456, 115400), O.geometry_hunyuan(115456, 115400, 179, 0.3)),
             (G.flux(66048, 512), O.geometry_flux(66048, 512, 51, 0.3)),
             (G.cogvideo(42466, 226), O.geometry_cogvideo(42466, 226, 49, 0.3)),
             # HunyuanVideo 129 frames: 118 800 visual tokens = 928 blocks + 16 -> 929 visual blocks, 112 pad rows
             (G.hunyuan(119056, 119000), O.geometry_hunyuan(119056, 119000, 185, 0.3))]
    for g, o in pairs:
        assert (g.n_blocks, g.nq_blocks, g.text_keys, g.kv_len, g.kv_zero_from, g.text_end_block, g.text_q_valid,
                g.first_frame_blocks) == (o.n_blocks, o.nq_blocks, o.text_keys, o.kv_len, o.kv_zero_from,
                                          o.text_end_block, o.text_q_valid, o.first_frame_blocks)
        assert g.gap == o.gap
    g = G.hunyuan(119056, 119000)
    assert (g.n_blocks, g.nq_blocks, g.vis_len, g.gap, g.kv_len) == (931, 929, 118800, 112, 119112)
    lib = N.lib()
    d = _desc(g)
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) > 0, lib.rsa_last_error_string()
    d.vis_len = 118801                 # nq_blocks no longer equals ceil(vis_len/128)? (still 929) -> n_blocks stays valid
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) > 0
    d.vis_len = 118912 + 1             # 930 visual blocks: inconsistent with nq_blocks = 929
    assert lib.rsa_attn_workspace_bytes(C.byref(d)) == 0


def test_hot_path_refuses_cpu_tensors():
    q = torch.zeros(1, 2, 1024, 128, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.Plan(q, q, q, G.wan(1024), 2, 0.3)
    with pytest.raises(RuntimeError):
        ops.permute_rows(torch.zeros(1, 4, 8), torch.arange(4))


def test_peer_route_validation():
    """The gather / scatter entry points check the route before touching the device."""
    lib = N.lib()
    d = _desc(G.hunyuan(1280, 1224))
    r = N.PeerRoute()
    p = N.PrepDesc()
    p.rows, p.dst_row = 1280, 0
    r.n_ranks, r.rank, r.rows_per_rank, r.heads_total = 2, 0, 600, 4            # 2 x 600 != 1280
    assert lib.rsa_qkv_prep_gather(C.byref(p), C.byref(d), C.byref(r), 16, 16, 16, 0, None, 0, None) == -1
    assert b"rows" in lib.rsa_last_error_string()
    r.rows_per_rank = 640
    r.heads_total = 5                                                             # != n_ranks * heads
    assert lib.rsa_qkv_prep_gather(C.byref(p), C.byref(d), C.byref(r), 16, 16, 16, 0, None, 0, None) == -1
    r.heads_total = 4                                                             # tables still null
    assert lib.rsa_qkv_prep_gather(C.byref(p), C.byref(d), C.byref(r), 16, 16, 16, 0, None, 0, None) == -1
    assert b"table" in lib.rsa_last_error_string()
    d2 = _desc(G.hunyuan(1256, 1200))                                             # ragged visual segment: refused
    r.rows_per_rank = 628
    assert lib.rsa_qkv_prep_gather(C.byref(p), C.byref(d2), C.byref(r), 16, 16, 16, 0, None, 0, None) in (-1, -2)


def test_attention_grid_order_visits_every_tile_once():
    """Kernel 4's 1-D grid (attention_grid_slot, the inline function the kernel itself calls, through its host-side debug
    entry): for every mix of visual / text tile counts, head counts and front-set sizes, the CTAs of a launch visit each
    (head, query tile) exactly once; text pairs of the last `front` heads come first; with an odd number of visual tiles
    and an even number of text tiles the tail is re-paired (odd visual tile alone, text tiles together); the former order
    (flag) is head by head with the (2p, 2p+1) pairing."""
    lib = N.lib()
    out = (C.c_int * 5)()

    def slot(i, nqt, nqv, n_bh, front, former=0):
        lib.rsa_debug_attention_grid_slot(i, nqt, nqv, n_bh, front, former, C.byref(out))
        return tuple(out)

    for nqv in range(0, 10):
        for ntxt in range(0, 5):
            nqt = nqv + ntxt
            if nqt == 0:
                continue
            n_pairs = (nqt + 1) // 2
            for n_bh in (1, 2, 3, 7):
                for front in {0, 1, 2, n_bh, n_bh + 3}:
                    if nqv < 2:
                        front = n_bh                 # what the library passes when there are no visual pairs
                    elif ntxt == 0:
                        front = 0
                    for former in (0, 1):
                        for nqv_arg in {nqv, (1 << 20) if ntxt == 0 else nqv}:    # rsa_masked_attention's "all visual"
                            seen = {}
                            for i in range(n_pairs * n_bh):
                                bh, pair, t0, t1, rep = slot(i, nqt, nqv_arg, n_bh, front, former)
                                assert 0 <= bh < n_bh and 0 <= pair < n_pairs and 0 <= t0 < nqt
                                if former:
                                    assert (bh, pair, t0, t1, rep) == (i // n_pairs, n_pairs - 1 - i % n_pairs,
                                                                       2 * pair, 2 * pair + 1, 0)
                                for t in (t0, t1):
                                    if t < nqt:
                                        seen[(bh, t)] = seen.get((bh, t), 0) + 1
                            assert len(seen) == n_bh * nqt and set(seen.values()) == {1}
    # HunyuanVideo 129 frames: 929 visual + 2 text tiles, 24 heads, text pairs of the last 3 heads in front
    nqt, nqv, n_bh, front = 931, 929, 24, 3
    # (two "text pairs" per head here: the re-paired text tiles, and the odd visual tile on its own)
    assert [slot(i, nqt, nqv, n_bh, front) for i in range(4)] == [(21, 465, 929, 930, 1), (21, 464, 928, 931, 1),
                                                                  (22, 465, 929, 930, 1), (22, 464, 928, 931, 1)]
    assert slot(6, nqt, nqv, n_bh, front) == (0, 465, 929, 930, 1)      # head 0 in order: its text tiles ...
    assert slot(7, nqt, nqv, n_bh, front) == (0, 464, 928, 931, 1)      # ... the odd visual tile alone ...
    assert slot(8, nqt, nqv, n_bh, front) == (0, 463, 926, 927, 0)      # ... then the visual pairs, descending
    assert slot(6 + 21 * 466, nqt, nqv, n_bh, front) == (21, 463, 926, 927, 0)   # last heads: their visual pairs
    assert slot(24 * 466 - 1, nqt, nqv, n_bh, front) == (23, 0, 0, 1, 0)
    # and the number of front heads the library derives from a descriptor (148 SMs assumed without a device)
    d = _desc(G.hunyuan(119056, 119000), h=24, top_k=185)
    assert lib.rsa_debug_front_text_heads(C.byref(d)) == 3
    d = _desc(G.flux(66048, 512), h=24, top_k=51)
    assert lib.rsa_debug_front_text_heads(C.byref(d)) == 7
    d = _desc(G.wan(75600, 28), h=40, top_k=147)
    assert lib.rsa_debug_front_text_heads(C.byref(d)) == 0
    d = _desc(G.hunyuan(1280, 1224))                                   # tiny: every head is in the front set
    assert lib.rsa_debug_front_text_heads(C.byref(d)) == 2


def test_gilbert_recursion_helpers_against_reference():
    """sgn / in_bounds / gilbert_xyz2d_r keep their reference names and results (`gilbert_helpers.json`: the reference's
    own functions on seeded random frames -- shifted origins, index offsets, flipped major axis; oracle/make_golden.py).
    The recursion runs in csrc/gilbert.cc (rsa_gilbert_xyz2d_r); a point outside the frame is an error, not a value."""
    import json

    import utils.jenga_gilbert as J
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gilbert_helpers.json")) as f:
        g = json.load(f)
    for v, want in g["sgn"].items():
        assert J.sgn(int(v)) == want
    for c in g["cases"]:
        a = c["args"]
        assert J.in_bounds(*a[1:]) == c["in_bounds"]
        assert J.in_bounds(*c["outside"], *a[4:]) == c["outside_in_bounds"]
        assert J.gilbert_xyz2d_r(*a) == c["index"]
        if not c["outside_in_bounds"]:
            with pytest.raises(ValueError):
                J.gilbert_xyz2d_r(a[0], *c["outside"], *a[4:])


# names of the reference that this repo does not provide, each with the reason
NOT_PROVIDED = {
    "jenga_gilbert": {"transpose_gilbert_mapping", "sliced_gilbert_mapping", "block_wise_mapping",
                      "sliced_gilbert_block_neighbor_mapping",            # variants no reference script calls
                      "visualize_gilbert_curve", "visualize_gilbert_curves_comparison"},   # matplotlib visualisers
}


def test_api_surface_matches_reference():
    """`api_signatures.json` lists every function and class (__init__, __call__) the reference's hot-path modules define,
    with parameter names, kinds and defaults (inspect.signature on the unmodified reference, oracle/make_golden.py).  The
    mirror modules must define the same names with the same parameters in the same order and the same defaults; they may
    give a default where the reference has none and may append parameters that have defaults (mask re-use, fused
    pre-attention steps)."""
    import importlib
    import inspect
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_signatures.json")) as f:
        api = json.load(f)

    def sig(fn):
        return [(p.name, p.kind.name, None if p.default is inspect._empty else repr(p.default))
                for p in inspect.signature(fn).parameters.values()]

    def compare(ref_params, fn, where, problems):
        ours = sig(fn)
        for i, (name, kind, default) in enumerate(ref_params):
            if i >= len(ours):
                problems.append(f"{where}: parameter {name} is missing")
                continue
            oname, okind, odefault = ours[i]
            if oname != name or okind != kind:
                problems.append(f"{where}: parameter {i} is {oname} ({okind}), the reference has {name} ({kind})")
            elif default is not None and odefault != default:
                problems.append(f"{where}: {name} defaults to {odefault}, the reference to {default}")
        for name, kind, default in ours[len(ref_params):]:
            if default is None and kind not in ("VAR_POSITIONAL", "VAR_KEYWORD"):
                problems.append(f"{where}: extra parameter {name} has no default")

    problems = []
    for mod, entry in api.items():
        ours = importlib.import_module("utils.jenga_gilbert" if mod == "jenga_gilbert" else "rectified_spaattn." + mod)
        for name, e in entry.items():
            if name in NOT_PROVIDED.get(mod, ()):
                assert not hasattr(ours, name)
                continue
            obj = getattr(ours, name, None)
            if obj is None:
                problems.append(f"{mod}.{name} is missing")
            elif e["kind"] == "function":
                compare(e["params"], obj, f"{mod}.{name}", problems)
            else:
                compare(e["init"], obj.__init__, f"{mod}.{name}.__init__", problems)
                if "call" in e:
                    compare(e["call"], obj.__call__, f"{mod}.{name}.__call__", problems)
    assert not problems, "\n".join(problems)


def test_attn_processor_helpers():
    """get_attn_processors / set_attn_processor (reference attn_processor.py:6-62) on a small module tree of fake
    attention layers: key naming, one processor for all, a dict keyed like the getter's (consumed while it is applied),
    ValueError on a dict of the wrong size.  The expected values are what the reference itself returns on this tree
    (checked in the build container)."""
    import torch

    from rectified_spaattn.attn_processor import get_attn_processors, set_attn_processor

    class Attn(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p, self.inner = "default", torch.nn.Linear(2, 2)

        def get_processor(self):
            return self.p

        def set_processor(self, p):
            self.p = p

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.attn1, self.attn2, self.ff = Attn(), Attn(), torch.nn.Linear(2, 2)

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.blocks, self.single = torch.nn.ModuleList([Block(), Block()]), Attn()
            self.proj = torch.nn.Linear(2, 2)

    m = Model()
    keys = ["blocks.0.attn1.processor", "blocks.0.attn2.processor", "blocks.1.attn1.processor",
            "blocks.1.attn2.processor", "single.processor"]
    assert sorted(get_attn_processors(m)) == keys and set(get_attn_processors(m).values()) == {"default"}
    set_attn_processor(m, "X")
    assert get_attn_processors(m) == {k: "X" for k in keys}
    procs = {k: k.upper() for k in keys}
    arg = dict(procs)
    set_attn_processor(m, arg)
    assert get_attn_processors(m) == procs and arg == {}
    with pytest.raises(ValueError, match="number of processors 1 does not match"):
        set_attn_processor(m, {"a": 1})


def test_attn_helpers_against_reference_values():
    """get_cu_seqlens / get_attn_mask / get_flash_attn_params of rectified_spaattn/attn.py (reference :34-57, :157-178).
    The reference allocates on "cuda" unconditionally; the expected values below are its own results with that
    allocation redirected to the CPU in the build container."""
    import torch

    from rectified_spaattn import attn as A
    cu = A.get_cu_seqlens(1000, 256, [200], device="cpu")
    assert cu.dtype == torch.int32 and cu.tolist() == [0, 1200, 1256]
    cu = A.get_cu_seqlens(64, 16, [3, 16, 0], device="cpu")
    assert cu.tolist() == [0, 67, 80, 160, 160, 224, 240]
    m = A.get_attn_mask(64, 16, [3, 16, 0], device="cpu")
    assert m.shape == (3, 1, 1, 80) and m.dtype == torch.bool
    assert m[:, 0, 0].sum(1).tolist() == [67, 80, 64] and bool(m[0, 0, 0, :67].all()) and not bool(m[0, 0, 0, 67:].any())
    cq, ck, sq, sk = A.get_flash_attn_params(1000, 256, [200], device="cpu")
    assert cq.tolist() == ck.tolist() == [0, 1200, 1256] and (sq, sk) == (1256, 1256)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.
    No Python file of the package, and no C/CUDA source, refers to it; bench.py imports it inside the CPU arm (class
    CpuArm) only -- and that arm imports nothing of the product (no rsa_b200, so no librsa_b200.so in its process)."""
    import re
    pkg = os.path.join(REPO, "rectified-spaattn_b200")
    offenders = []
    for root, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d not in ("build", "__pycache__")]
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(root, fn), encoding="utf-8", errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "oracle/" in text:
                    offenders.append(os.path.join(root, fn))
    assert not offenders, offenders
    bench_src = open(os.path.join(REPO, "bench.py"), encoding="utf-8").read()
    body = bench_src.split("class CpuArm", 1)[1].split("\ndef cpu_c1_full", 1)[0]
    imports = re.findall(r"^\s*from oracle import .*$", bench_src, re.M)
    assert imports and all(line.strip() in body for line in imports)
    assert not re.search(r"^\s*(from|import)\s+(rsa_b200|rectified_spaattn|utils)\b", body, re.M)


def test_product_library_has_one_attention_kernel():
    """north_star: "no multi-backend dispatch".  The mma.sync cross-check implementation of kernel 4 lives in the tests'
    own library (tests/xcheck/librsa_xcheck.so); the product library exports no switch and contains no such kernel."""
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "rsa_set_attention_impl" not in syms and "attn_mma" not in syms
    everything = subprocess.run(["nm", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "attn_mma" not in everything
    assert not hasattr(ops, "set_attention_impl")
    xlib = os.path.join(REPO, "tests", "xcheck", "librsa_xcheck.so")
    assert os.path.exists(xlib), "build() compiles the tests' cross-check library too"
    xsyms = subprocess.run(["nm", "-D", "--defined-only", xlib], capture_output=True, text=True, check=True).stdout
    assert "rsa_xcheck_attention" in xsyms
