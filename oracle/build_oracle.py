"""ORACLE build recipe (test infrastructure): compiles oracle_kernels.c into oracle/_build/liboracle.so."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(_HERE, "oracle_kernels.c")
    out = os.path.join(out_dir, "liboracle.so")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    # -ffp-contract=off: every fused operation in the oracle is an explicit fmaf()
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", out, "-lm"])
    return out


if __name__ == "__main__":
    print(build(force=True))
