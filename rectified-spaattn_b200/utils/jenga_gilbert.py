"""Drop-in mirror of the reference utils/jenga_gilbert.py (host geometry), backed by the C++ implementation in
csrc/gilbert.cc.  Same names, arguments and return types:
  gilbert_xyz2d(x, y, z, width, height, depth, axis_order)            reference :12-54
  sgn / in_bounds / gilbert_xyz2d_r                                   reference :57, :60-81, :84-288 (the recursion's
                                                                      helpers, kept importable; the recursion itself
                                                                      runs in csrc/gilbert.cc as a loop)
  gilbert_mapping(t, h, w, transpose_order=None, axis_order)          reference :458-504 -> two Python lists
  gilbert_block_neighbor_mapping(t, h, w, block_size=128, ...)        reference :613-693 -> torch.bool [NB, NB]
The reference's unused sliced/transposed variants and matplotlib visualisers are out of scope.
"""
import torch

from rsa_b200 import ops as _ops


def gilbert_mapping(t, h, w, transpose_order=None, axis_order=("w", "h", "t")):
    if transpose_order is not None:
        raise NotImplementedError("transpose_order is unused by every reference script and is not provided")
    l2h, h2l = _ops.gilbert_mapping(int(t), int(h), int(w), axis_order)
    return l2h.tolist(), h2l.tolist()


def gilbert_block_neighbor_mapping(t, h, w, block_size=128, transpose_order=None, axis_order=("w", "h", "t")):
    if transpose_order is not None:
        raise NotImplementedError("transpose_order is unused by every reference script and is not provided")
    return _ops.gilbert_block_neighbors(int(t), int(h), int(w), int(block_size), axis_order)


def gilbert_xyz2d(x, y, z, width, height, depth, axis_order=None):
    # single-point query through the full mapping of the same box (setup-time helper, not a hot path)
    l2h, _ = _ops.gilbert_mapping(int(depth), int(height), int(width), axis_order)
    return int(l2h[(z * height + y) * width + x])


def sgn(x):
    return (x > 0) - (x < 0)


def in_bounds(x, y, z, x_s, y_s, z_s, ax, ay, az, bx, by, bz, cx, cy, cz):
    """Is (x, y, z) inside the box with origin (x_s, y_s, z_s) spanned by a + b + c?  Along an axis with a negative
    extent d the box covers (origin + d, origin], otherwise [origin, origin + d)."""
    for p, o, d in ((x, x_s, ax + bx + cx), (y, y_s, ay + by + cy), (z, z_s, az + bz + cz)):
        if (p > o or p <= o + d) if d < 0 else (p < o or p >= o + d):
            return False
    return True


def gilbert_xyz2d_r(cur_idx, x_dst, y_dst, z_dst, x, y, z, ax, ay, az, bx, by, bz, cx, cy, cz):
    from rsa_b200 import native as _native
    r = _native.lib().rsa_gilbert_xyz2d_r(*(int(v) for v in (cur_idx, x_dst, y_dst, z_dst, x, y, z, ax, ay, az, bx, by,
                                                             bz, cx, cy, cz)))
    if r < 0:
        raise ValueError(_native.lib().rsa_last_error_string().decode())
    return int(r)


def build_multi_curve(latent_time, latent_height, latent_width, axis_order_list, device="cuda"):
    """Same result as build_multi_curve in the reference scripts (scripts/main_hunyuan.py:23-42):
    [[linear_to_hilbert (long, device), hilbert_order (long, device), block_neighbor_list (bool, CPU)], ...]."""
    t, h, w = int(latent_time), int(latent_height), int(latent_width)
    if (h * w) % 4 != 0:
        raise ValueError(f"latent_height_ * latent_width_ must be divisible by 4, but got {h * w}")
    out = []
    for axis_order in axis_order_list:
        l2h, h2l = _ops.gilbert_mapping(t, h, w, axis_order)
        nbr = _ops.gilbert_block_neighbors(t, h, w, 128, axis_order)
        out.append([l2h.to(device), h2l.to(device), nbr])
    return out
