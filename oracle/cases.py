"""ORACLE support (test infrastructure): the small synthetic cases shared by make_golden.py and tests/."""
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
