"""ORACLE support (test infrastructure): the small synthetic cases shared by make_golden.py and tests/."""
import numpy as np

from oracle import rsa_oracle as O

CASES = {
    # name: (family, grid(t,h,w), text_len, num_true_delta, heads, top_k, p, regime, seed)
    "wan_c1": ("wan", (4, 16, 16), 0, 0, 2, 2, 0.3, "walk", 1),
    "wan_ragged": ("wan", (3, 10, 13), 0, 0, 2, 1, 0.3, "walk", 2),
    "wan_iid": ("wan", (4, 16, 16), 0, 0, 2, 3, 0.5, "iid", 3),
    "hunyuan_small": ("hunyuan", (4, 16, 16), 256, 200, 2, 2, 0.3, "walk", 4),
    "hunyuan_fulltext": ("hunyuan", (2, 16, 16), 256, 256, 2, 1, 0.3, "cluster", 5),
    "flux_small": ("flux", (1, 32, 32), 512, 512, 2, 1, 0.3, "walk", 6),
    "cog_small": ("cogvideo", (4, 16, 16), 226, 226, 2, 2, 0.3, "walk", 7),
    "hunyuan_mid": ("hunyuan", (8, 16, 32), 256, 77, 2, 6, 0.3, "walk", 8),
    # ragged visual segment (780 = 6 blocks + 12 tokens), the small analogue of HunyuanVideo 129 frames (118 800 + 256):
    # the reference cannot run it as given (it raises, rectified_hunyuan_attn.py:356); see geometry_hunyuan
    "hunyuan_ragged": ("hunyuan", (5, 12, 13), 256, 200, 2, 2, 0.3, "walk", 9),
}

# Cases outside CASES (the GPU parity tests iterate over CASES and assume 128 columns): name -> the CASES tuple + head_dim.
# cog_d64 = CogVideoX's real head dimension (the reference kernel takes Lk in {16, 32, 64, 128}, wan21 :121); it pins the
# oracle's 64-column arithmetic (scale 64^-1/2, pooled statistics, GAPR, R, C) against the unmodified reference on CPU.
EXTRA_CASES = {
    "cog_d64": ("cogvideo", (8, 16, 32), 226, 226, 2, 6, 0.3, "walk", 31, 64),
}

# cases the unmodified reference cannot run on the caller's tensors (fixtures are made on the explicitly padded layout)
REFERENCE_NEEDS_PADDED_LAYOUT = ("hunyuan_ragged",)


def case_inputs(name):
    spec = CASES[name] if name in CASES else EXTRA_CASES[name]
    fam, (t, h, w), text_len, ntrue_d, heads, top_k, p, regime, seed = spec[:9]
    head_dim = spec[9] if len(spec) > 9 else 128
    nv = t * h * w
    s = nv + text_len
    q, k, v = O.synth_qkv(heads, s, head_dim, regime, seed)
    return fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v


# Mid-size cases (one head each): large enough for the selection kernel's wider sorting-network instantiations
# (rectified-spaattn_b200/csrc/block_select.cu: 256 threads x 1 / 2 / 4 / 8 entries per thread) and for the
# joint family's text-aggregate insertion when the visual blocks alone fill a power of two (512), small enough that
# the unmodified reference builds their masks on CPU in seconds (oracle/make_golden.py `mid`).  Only the mask-building
# stages are pinned at these sizes (probabilities, GAPR bytes, selection, R, C); the attention itself is size-independent
# and covered by CASES.  top_k = int((1 - sa_drop_rate) * img_blocks) as in the reference scripts.
#   name: (family, grid(t,h,w), text_len, num_true_delta, heads, top_k, p, regime, seed)
MID_CASES = {
    "wan_300": ("wan", (3, 100, 128), 0, 0, 1, O.select_block_num(0.75, 300), 0.3, "walk", 41),            # 300 entries
    "flux_512": ("flux", (1, 256, 256), 512, 512, 1, O.select_block_num(0.9, 512), 0.3, "walk", 42),       # 512 + aggregate
    "flux_513": ("flux", (1, 216, 304), 512, 512, 1, O.select_block_num(0.9, 513), 0.3, "cluster", 43),    # 514 entries
    "hunyuan_600": ("hunyuan", (6, 100, 128), 256, 200, 1, O.select_block_num(0.8, 600), 0.3, "iid", 44),  # threshold > top_k
    "hunyuan_1100": ("hunyuan", (11, 100, 128), 256, 200, 1, O.select_block_num(0.8, 1100), 0.3, "walk", 45),
    # every K block is a copy of the first block of its group of four: probabilities come in exact ties of four, and
    # the cut (top_k = 75 = 18 groups + 3) falls inside a tie group -- the tie-break contract (probability desc, index asc)
    "wan_ties": ("wan", (3, 100, 128), 0, 0, 1, O.select_block_num(0.75, 300), 0.3, "ties", 46),
}


def mid_case_inputs(name):
    fam, (t, h, w), text_len, ntrue_d, heads, top_k, p, regime, seed = MID_CASES[name]
    nv = t * h * w
    s = nv + text_len
    q, k, v = O.synth_qkv(heads, s, 128, "walk" if regime == "ties" else regime, seed)
    if regime == "ties":
        nb = nv // 128
        kb = k[:, :, : nb * 128].reshape(k.shape[0], k.shape[1], nb, 128, k.shape[3])
        kb[:] = kb[:, :, (np.arange(nb) // 4) * 4]
    return fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v
