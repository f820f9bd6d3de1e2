"""ORACLE support (test infrastructure): import the UNMODIFIED reference beside this repo.

The reference lives at /root/reference in the build container and (copied by oracle/stage_reference.py,
git-ignored) at baseline/_ref on the GPU box.  diffusers and matplotlib are not installed, so empty stub
modules stand in for the names the reference imports at module scope.  The reference keeps the
package names `rectified_spaattn` and `utils`, which this repo mirrors, so it is aliased as
`ref_rectified_spaattn` / `ref_utils` (namespace packages, no __init__.py).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "rectified_spaattn")):
            return cand
    return None


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


def load(names=("rectified_wan21_attn",)):
    """Returns {name: module} for reference modules under rectified_spaattn/, plus 'jenga_gilbert'."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not found (neither /root/reference nor baseline/_ref)")
    try:
        import diffusers  # noqa: F401
    except Exception:
        for n in ("diffusers", "diffusers.models", "diffusers.models.transformers"):
            _stub(n)
        _stub("diffusers.models.attention_processor", Attention=object, AttentionProcessor=object)
        _stub("diffusers.models.transformers.transformer_wan", _get_qkv_projections=None,
              _get_added_kv_projections=None)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        for n in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits"):
            _stub(n)
        _stub("mpl_toolkits.mplot3d", Axes3D=object)
    if "ref_rectified_spaattn" not in sys.modules:
        pkg = types.ModuleType("ref_rectified_spaattn")
        pkg.__path__ = [os.path.join(root, "rectified_spaattn")]
        sys.modules["ref_rectified_spaattn"] = pkg
        upkg = types.ModuleType("ref_utils")
        upkg.__path__ = [os.path.join(root, "utils")]
        sys.modules["ref_utils"] = upkg
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for n in names:
            if n == "jenga_gilbert":
                out[n] = importlib.import_module("ref_utils.jenga_gilbert")
            else:
                out[n] = importlib.import_module("ref_rectified_spaattn." + n)
    return out


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
