// gilbert.cc -- host geometry: generalised Hilbert ("Gilbert") curve index, token mappings and the
// block-neighbour matrix.  Replaces the pure-Python setup code of the reference
//   utils/jenga_gilbert.py:12-54, 84-288 (point query), :458-504 (mapping), :613-693 (neighbours)
// which takes 3.2 s + 4.5 s at the HunyuanVideo grid; this runs in milliseconds.
//
// The reference recursion only ever tail-calls itself, so the point query is a loop over a frame
// (origin o, major a, mid b, minor c).  Halving uses floor division (Python's //) on signed components.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "rsa.h"

namespace rsa {
void set_error(const char* fmt, ...);
}

namespace {

struct V3 {
  int64_t x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline int64_t sum(V3 a) { return a.x + a.y + a.z; }
inline int64_t sgn(int64_t v) { return (v > 0) - (v < 0); }
inline V3 dir(V3 a) { return {sgn(a.x), sgn(a.y), sgn(a.z)}; }
inline int64_t fdiv2(int64_t v) { return v >> 1; }  // floor division by two for negatives as well
inline V3 half(V3 a) { return {fdiv2(a.x), fdiv2(a.y), fdiv2(a.z)}; }
inline int64_t iabs(int64_t v) { return v < 0 ? -v : v; }
inline int64_t dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

inline bool axis_in(int64_t p, int64_t o, int64_t d) {
  return d < 0 ? !(p > o || p <= o + d) : !(p < o || p >= o + d);
}
inline bool inside(V3 p, V3 o, V3 a, V3 b, V3 c) {
  V3 d = a + b + c;
  return axis_in(p.x, o.x, d.x) && axis_in(p.y, o.y, d.y) && axis_in(p.z, o.z, d.z);
}

int64_t gilbert_index(V3 p, V3 a, V3 b, V3 c) {
  V3 o{0, 0, 0};
  int64_t idx = 0;
  for (;;) {
    const int64_t w = iabs(sum(a)), h = iabs(sum(b)), d = iabs(sum(c));
    const V3 da = dir(a), db = dir(b), dc = dir(c);
    if (h == 1 && d == 1) return idx + dot(da, p - o);
    if (w == 1 && d == 1) return idx + dot(db, p - o);
    if (w == 1 && h == 1) return idx + dot(dc, p - o);

    V3 a2 = half(a), b2 = half(b), c2 = half(c);
    const int64_t w2 = iabs(sum(a2)), h2 = iabs(sum(b2)), d2 = iabs(sum(c2));
    if ((w2 & 1) && w > 2) a2 = a2 + da;
    if ((h2 & 1) && h > 2) b2 = b2 + db;
    if ((d2 & 1) && d > 2) c2 = c2 + dc;

    if (2 * w > 3 * h && 2 * w > 3 * d) {  // long box: cut the major axis only
      if (inside(p, o, a2, b, c)) {
        a = a2;
        continue;
      }
      idx += iabs(sum(a2) * sum(b) * sum(c));
      o = o + a2;
      a = a - a2;
      continue;
    }
    if (3 * h > 4 * d) {  // minor axis stays whole
      if (inside(p, o, b2, c, a2)) {
        V3 na = b2, nb = c, nc = a2;
        a = na, b = nb, c = nc;
        continue;
      }
      idx += iabs(sum(b2) * sum(c) * sum(a2));
      if (inside(p, o + b2, a, b - b2, c)) {
        o = o + b2;
        b = b - b2;
        continue;
      }
      idx += iabs(sum(a) * sum(b - b2) * sum(c));
      o = o + (a - da) + (b2 - db);
      V3 na = -b2, nb = c, nc = -(a - a2);
      a = na, b = nb, c = nc;
      continue;
    }
    if (3 * d > 4 * h) {  // mid axis stays whole
      if (inside(p, o, c2, a2, b)) {
        V3 na = c2, nb = a2, nc = b;
        a = na, b = nb, c = nc;
        continue;
      }
      idx += iabs(sum(c2) * sum(a2) * sum(b));
      if (inside(p, o + c2, a, b, c - c2)) {
        o = o + c2;
        c = c - c2;
        continue;
      }
      idx += iabs(sum(a) * sum(b) * sum(c - c2));
      o = o + (a - da) + (c2 - dc);
      V3 na = -c2, nb = -(a - a2), nc = b;
      a = na, b = nb, c = nc;
      continue;
    }
    // all three axes cut: five sub-boxes visited in curve order
    if (inside(p, o, b2, c2, a2)) {
      V3 na = b2, nb = c2, nc = a2;
      a = na, b = nb, c = nc;
      continue;
    }
    idx += iabs(sum(b2) * sum(c2) * sum(a2));
    if (inside(p, o + b2, c, a2, b - b2)) {
      o = o + b2;
      V3 na = c, nb = a2, nc = b - b2;
      a = na, b = nb, c = nc;
      continue;
    }
    idx += iabs(sum(c) * sum(a2) * sum(b - b2));
    {
      V3 o2 = o + (b2 - db) + (c - dc);
      if (inside(p, o2, a, -b2, -(c - c2))) {
        o = o2;
        V3 nb = -b2, nc = -(c - c2);
        b = nb, c = nc;
        continue;
      }
    }
    idx += iabs(sum(a) * sum(-b2) * sum(-(c - c2)));
    {
      V3 o3 = o + (a - da) + b2 + (c - dc);
      if (inside(p, o3, -c, -(a - a2), b - b2)) {
        o = o3;
        V3 na = -c, nb = -(a - a2), nc = b - b2;
        a = na, b = nb, c = nc;
        continue;
      }
    }
    idx += iabs(sum(-c) * sum(-(a - a2)) * sum(b - b2));
    o = o + (a - da) + (b2 - db);
    V3 na = -b2, nb = c2, nc = -(a - a2);
    a = na, b = nb, c = nc;
  }
}

bool frame_for(int t, int h, int w, const char* axis_order, V3* a, V3* b, V3* c) {
  const V3 W{w, 0, 0}, H{0, h, 0}, T{0, 0, t};
  if (axis_order == nullptr) {  // size-based default (jenga_gilbert.py:34-54)
    if (w >= h && w >= t) {
      *a = W, *b = H, *c = T;
    } else if (h >= w && h >= t) {
      *a = H, *b = W, *c = T;
    } else {
      *a = T, *b = W, *c = H;
    }
    return true;
  }
  if (strlen(axis_order) != 3) return false;
  V3* dst[3] = {a, b, c};
  for (int i = 0; i < 3; ++i) {
    switch (axis_order[i]) {
      case 'w': *dst[i] = W; break;
      case 'h': *dst[i] = H; break;
      case 't': *dst[i] = T; break;
      default: return false;
    }
  }
  return true;
}

}  // namespace

extern "C" int64_t rsa_gilbert_xyz2d_r(int64_t cur_idx, int64_t x_dst, int64_t y_dst, int64_t z_dst, int64_t x,
                                       int64_t y, int64_t z, int64_t ax, int64_t ay, int64_t az, int64_t bx, int64_t by,
                                       int64_t bz, int64_t cx, int64_t cy, int64_t cz) {
  const V3 p{x_dst, y_dst, z_dst}, o{x, y, z}, a{ax, ay, az}, b{bx, by, bz}, c{cx, cy, cz};
  if (sum(a) == 0 || sum(b) == 0 || sum(c) == 0 || !inside(p, o, a, b, c)) {
    rsa::set_error("rsa_gilbert_xyz2d_r: the point is not inside the frame (or the frame is degenerate)");
    return -1;
  }
  return cur_idx + gilbert_index(p - o, a, b, c);  // the curve depends on the frame only through p - o
}

extern "C" int rsa_gilbert_map(int t, int h, int w, const char* axis_order, int64_t* l2h, int64_t* h2l) {
  if (t <= 0 || h <= 0 || w <= 0) {
    rsa::set_error("rsa_gilbert_map: grid sizes must be positive (t=%d h=%d w=%d)", t, h, w);
    return RSA_ERR_ARG;
  }
  if (!l2h && !h2l) {
    rsa::set_error("rsa_gilbert_map: both output pointers are null");
    return RSA_ERR_ARG;
  }
  V3 a, b, c;
  if (!frame_for(t, h, w, axis_order, &a, &b, &c)) {
    rsa::set_error("rsa_gilbert_map: axis_order must be three characters over {w,h,t}");
    return RSA_ERR_ARG;
  }
  const int64_t n = (int64_t)t * h * w;
  for (int z = 0; z < t; ++z)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const int64_t lin = ((int64_t)z * h + y) * w + x;
        const int64_t g = gilbert_index(V3{x, y, z}, a, b, c);
        if (g < 0 || g >= n) {
          rsa::set_error("rsa_gilbert_map: curve index %lld out of range at (%d,%d,%d)", (long long)g, x, y, z);
          return RSA_ERR_ARG;
        }
        if (l2h) l2h[lin] = g;
        if (h2l) h2l[g] = lin;
      }
  return RSA_OK;
}

extern "C" int rsa_gilbert_block_neighbors(int t, int h, int w, int block_size, const char* axis_order,
                                           uint8_t* out) {
  if (t <= 0 || h <= 0 || w <= 0 || block_size <= 0 || !out) {
    rsa::set_error("rsa_gilbert_block_neighbors: bad argument");
    return RSA_ERR_ARG;
  }
  const int64_t n = (int64_t)t * h * w;
  const int64_t nb = (n + block_size - 1) / block_size;
  std::vector<int64_t> l2h((size_t)n);
  int rc = rsa_gilbert_map(t, h, w, axis_order, l2h.data(), nullptr);
  if (rc != RSA_OK) return rc;
  std::vector<int32_t> color((size_t)n);
  for (int64_t i = 0; i < n; ++i) color[(size_t)i] = (int32_t)(l2h[(size_t)i] / block_size);
  memset(out, 0, (size_t)(nb * nb));
  for (int z = 0; z < t; ++z)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const int32_t me = color[(size_t)(((int64_t)z * h + y) * w + x)];
        out[(size_t)me * nb + me] = 1;
        for (int dz = -1; dz <= 1; ++dz) {
          const int nz = z + dz;
          if (nz < 0 || nz >= t) continue;
          for (int dy = -1; dy <= 1; ++dy) {
            const int ny = y + dy;
            if (ny < 0 || ny >= h) continue;
            for (int dx = -1; dx <= 1; ++dx) {
              const int nx = x + dx;
              if (nx < 0 || nx >= w) continue;
              out[(size_t)me * nb + color[(size_t)(((int64_t)nz * h + ny) * w + nx)]] = 1;
            }
          }
        }
      }
  return RSA_OK;
}
