"""ORACLE (test infrastructure, not product code) -- Gilbert-curve geometry.

CPU restatement of the reference's host-side geometry:
  * point query  gilbert_xyz2d / gilbert_xyz2d_r      utils/jenga_gilbert.py:12-54, 84-288
  * bounds test  in_bounds / sgn                      utils/jenga_gilbert.py:57-81
  * mapping      gilbert_mapping                      utils/jenga_gilbert.py:458-504
  * neighbours   gilbert_block_neighbor_mapping       utils/jenga_gilbert.py:613-693
  * permute      x[:, hilbert_order] / x[:, linear_to_hilbert]   scripts/main_hunyuan.py:88-89, 183

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
Pinned against the reference by tests/golden/gilbert_*.json (made by oracle/make_golden.py, which imports
the reference in the build container).

The reference recursion is tail-recursive everywhere, so the restatement is a loop over a "frame"
(origin o, major a, mid b, minor c -- each an integer 3-vector).  Python's floor division on negative
components (a // 2) is part of the algorithm and is kept.
"""
from __future__ import annotations

import numpy as np


def _sgn(v):
    return (v > 0) - (v < 0)


def _vsum(v):
    return v[0] + v[1] + v[2]


def _add(*vs):
    return (sum(v[0] for v in vs), sum(v[1] for v in vs), sum(v[2] for v in vs))


def _neg(v):
    return (-v[0], -v[1], -v[2])


def _sub(u, v):
    return (u[0] - v[0], u[1] - v[1], u[2] - v[2])


def _inside(p, o, a, b, c):
    """Point p inside the box spanned from origin o by a+b+c (jenga_gilbert.py:60-81)."""
    for k in range(3):
        d = a[k] + b[k] + c[k]
        if d < 0:
            if p[k] > o[k] or p[k] <= o[k] + d:
                return False
        else:
            if p[k] < o[k] or p[k] >= o[k] + d:
                return False
    return True


def gilbert_index(x, y, z, width, height, depth, axis_order=("w", "h", "t")):
    """Index along the generalised Hilbert curve of voxel (x,y,z) in a width x height x depth box.

    Follows gilbert_xyz2d (jenga_gilbert.py:12-54) + gilbert_xyz2d_r (:84-288)."""
    if axis_order is not None:
        vec = {"w": (width, 0, 0), "h": (0, height, 0), "t": (0, 0, depth)}
        a, b, c = vec[axis_order[0]], vec[axis_order[1]], vec[axis_order[2]]
    elif width >= height and width >= depth:
        a, b, c = (width, 0, 0), (0, height, 0), (0, 0, depth)
    elif height >= width and height >= depth:
        a, b, c = (0, height, 0), (width, 0, 0), (0, 0, depth)
    else:
        a, b, c = (0, 0, depth), (width, 0, 0), (0, height, 0)

    p = (x, y, z)
    o = (0, 0, 0)
    idx = 0
    while True:
        w, h, d = abs(_vsum(a)), abs(_vsum(b)), abs(_vsum(c))
        da = tuple(_sgn(v) for v in a)
        db = tuple(_sgn(v) for v in b)
        dc = tuple(_sgn(v) for v in c)
        # straight runs (:100-107)
        if h == 1 and d == 1:
            return idx + sum(da[k] * (p[k] - o[k]) for k in range(3))
        if w == 1 and d == 1:
            return idx + sum(db[k] * (p[k] - o[k]) for k in range(3))
        if w == 1 and h == 1:
            return idx + sum(dc[k] * (p[k] - o[k]) for k in range(3))

        a2 = tuple(v // 2 for v in a)
        b2 = tuple(v // 2 for v in b)
        c2 = tuple(v // 2 for v in c)
        w2, h2, d2 = abs(_vsum(a2)), abs(_vsum(b2)), abs(_vsum(c2))
        # prefer even steps (:118-125)
        if (w2 % 2) and w > 2:
            a2 = _add(a2, da)
        if (h2 % 2) and h > 2:
            b2 = _add(b2, db)
        if (d2 % 2) and d > 2:
            c2 = _add(c2, dc)

        if 2 * w > 3 * h and 2 * w > 3 * d:
            # wide: split the major axis only (:128-147)
            if _inside(p, o, a2, b, c):
                a = a2
                continue
            idx += abs(_vsum(a2) * _vsum(b) * _vsum(c))
            o, a = _add(o, a2), _sub(a, a2)
            continue

        if 3 * h > 4 * d:
            # do not split the minor axis (:150-184)
            if _inside(p, o, b2, c, a2):
                a, b, c = b2, c, a2
                continue
            idx += abs(_vsum(b2) * _vsum(c) * _vsum(a2))
            o1 = _add(o, b2)
            if _inside(p, o1, a, _sub(b, b2), c):
                o, b = o1, _sub(b, b2)
                continue
            idx += abs(_vsum(a) * _vsum(_sub(b, b2)) * _vsum(c))
            o = _add(o, _sub(a, da), _sub(b2, db))
            a, b, c = _neg(b2), c, _neg(_sub(a, a2))
            continue

        if 3 * d > 4 * h:
            # do not split the mid axis (:187-219)
            if _inside(p, o, c2, a2, b):
                a, b, c = c2, a2, b
                continue
            idx += abs(_vsum(c2) * _vsum(a2) * _vsum(b))
            o1 = _add(o, c2)
            if _inside(p, o1, a, b, _sub(c, c2)):
                o, c = o1, _sub(c, c2)
                continue
            idx += abs(_vsum(a) * _vsum(b) * _vsum(_sub(c, c2)))
            o = _add(o, _sub(a, da), _sub(c2, dc))
            a, b, c = _neg(c2), _neg(_sub(a, a2)), b
            continue

        # regular case: split all three (:222-288), five octant groups in curve order
        if _inside(p, o, b2, c2, a2):
            a, b, c = b2, c2, a2
            continue
        idx += abs(_vsum(b2) * _vsum(c2) * _vsum(a2))

        o1 = _add(o, b2)
        if _inside(p, o1, c, a2, _sub(b, b2)):
            o, a, b, c = o1, c, a2, _sub(b, b2)
            continue
        idx += abs(_vsum(c) * _vsum(a2) * _vsum(_sub(b, b2)))

        o2 = _add(o, _sub(b2, db), _sub(c, dc))
        if _inside(p, o2, a, _neg(b2), _neg(_sub(c, c2))):
            o, b, c = o2, _neg(b2), _neg(_sub(c, c2))
            continue
        idx += abs(_vsum(a) * _vsum(_neg(b2)) * _vsum(_neg(_sub(c, c2))))

        o3 = _add(o, _sub(a, da), b2, _sub(c, dc))
        if _inside(p, o3, _neg(c), _neg(_sub(a, a2)), _sub(b, b2)):
            o, a, b, c = o3, _neg(c), _neg(_sub(a, a2)), _sub(b, b2)
            continue
        idx += abs(_vsum(_neg(c)) * _vsum(_neg(_sub(a, a2))) * _vsum(_sub(b, b2)))

        o = _add(o, _sub(a, da), _sub(b2, db))
        a, b, c = _neg(b2), c2, _neg(_sub(a, a2))


def gilbert_mapping(t, h, w, axis_order=("w", "h", "t")):
    """(linear_to_hilbert, hilbert_to_linear) as int64 arrays; linear = z*h*w + y*w + x (:458-504)."""
    n = t * h * w
    l2h = np.zeros(n, dtype=np.int64)
    h2l = np.zeros(n, dtype=np.int64)
    for z in range(t):
        for y in range(h):
            for x in range(w):
                lin = (z * h + y) * w + x
                g = gilbert_index(x, y, z, w, h, t, axis_order)
                l2h[lin] = g
                h2l[g] = lin
    return l2h, h2l


def gilbert_block_neighbors(t, h, w, block_size=128, axis_order=("w", "h", "t"), l2h=None):
    """bool [NB, NB]: blocks (curve index // block_size) that own 26-adjacent voxels, self included (:613-693)."""
    n = t * h * w
    nb = (n + block_size - 1) // block_size
    if l2h is None:
        l2h, _ = gilbert_mapping(t, h, w, axis_order)
    color = (np.asarray(l2h).reshape(t, h, w) // block_size).astype(np.int64)  # [z, y, x]
    out = np.zeros((nb, nb), dtype=bool)
    out[np.arange(nb), np.arange(nb)] = True
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == 0 and dy == 0 and dz == 0:
                    continue
                zs = slice(max(0, -dz), t - max(0, dz))
                ys = slice(max(0, -dy), h - max(0, dy))
                xs = slice(max(0, -dx), w - max(0, dx))
                zn = slice(max(0, dz), t - max(0, -dz))
                yn = slice(max(0, dy), h - max(0, -dy))
                xn = slice(max(0, dx), w - max(0, -dx))
                out[color[zs, ys, xs].ravel(), color[zn, yn, xn].ravel()] = True
    return out


def permute_rows(x, index):
    """out[b, i, :] = x[b, index[i], :]  (scripts/main_hunyuan.py:88-89, :183 -- both directions are gathers)."""
    return x[:, np.asarray(index)]
