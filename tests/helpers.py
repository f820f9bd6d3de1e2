"""Shared helpers for the parity tests: the same synthetic case through the oracle and through the C ABI."""
import numpy as np

from oracle import cases as C
from oracle import gilbert_oracle as GO
from oracle import rsa_oracle as O


def oracle_geometry(fam, nv, s, text_len, ntrue_d, top_k, p, t):
    if fam == "wan":
        return O.geometry_wan(s, top_k, p, (nv + 127) // 128 // t)
    if fam == "hunyuan":
        return O.geometry_hunyuan(s, nv + ntrue_d, top_k, p)
    if fam == "flux":
        return O.geometry_flux(s, text_len, top_k, p)
    return O.geometry_cogvideo(s, text_len, top_k, p)


def product_geometry(fam, nv, s, text_len, ntrue_d, t):
    from rsa_b200 import geometry as G
    if fam == "wan":
        return G.wan(s, (nv + 127) // 128 // t)
    if fam == "hunyuan":
        return G.hunyuan(s, nv + ntrue_d)
    if fam == "flux":
        return G.flux(s, text_len)
    return G.cogvideo(s, text_len)


def load_case(name):
    fam, (t, h, w), nv, s, text_len, ntrue_d, heads, top_k, p, q, k, v = C.case_inputs(name)
    nbr = GO.gilbert_block_neighbors(t, h, w)
    ogeo = oracle_geometry(fam, nv, s, text_len, ntrue_d, top_k, p, t)
    return dict(fam=fam, grid=(t, h, w), nv=nv, s=s, text_len=text_len, ntrue_d=ntrue_d, heads=heads, top_k=top_k,
                p=p, q=q, k=k, v=v, nbr=nbr, ogeo=ogeo)


def cos_sim(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


def assert_prep_close(got, ref, what=""):
    """Kernel 0 / prep-oracle bar against the PyTorch op sequence (arrays of bf16 values held in fp32, last axis 128).

    The normalised value of an element can differ by one bf16 ulp (rsqrt and the reduction order of mean(x^2) differ by
    fp32 ulps between implementations); the rotation then mixes the two elements of a pair, so the error of an output
    is bounded by an ulp of the PAIR's magnitude, not of its own (a small output next to a large partner).  Hence:
    |got - ref| <= 2^-7 * |pair| + 1e-6 everywhere, and >= 99.9 % of the elements identical."""
    got = np.asarray(got, dtype=np.float32)
    ref = np.asarray(ref, dtype=np.float32)
    pair = np.sqrt(ref[..., 0::2] ** 2 + ref[..., 1::2] ** 2).repeat(2, axis=-1)
    tol = pair * 2.0 ** -7 + 1e-6
    worst = float(np.max(np.abs(got - ref) / tol))
    assert worst <= 1.0, f"{what}: worst deviation {worst:.2f} x the bound"
    same = float(np.mean(got == ref))
    assert same >= 0.999, f"{what}: only {same:.5f} of the elements identical"
