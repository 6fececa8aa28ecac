// Device build of the oriented-cylinder bounds of the closest-triangle search (PillNode, moments.h).
//
// The reference's igl::AABB (/root/reference/include/igl/AABB.cpp:812-978) bounds every subtree by an axis-aligned box.
// The result of a closest-point query does not depend on which bounds prune the search, only on their being
// conservative, so the bounds here are chosen for the far field: per cluster (a child lane of the 4-ary tree: an
// internal node's subtree or one face) a cylinder about the cluster's area-weighted mean normal.
//
// Clusters are contiguous ranges of the depth-first leaf order (wn.h: CountLevel / StartLevel).  Small clusters are
// bounded by one thread each, large ones by one block each; three passes over the cluster's fp64 vertices:
//   1. normal sum, area sum, axis-aligned extent       -> axis (fp32, re-normalised in fp64)
//   2. extent in an orthonormal frame (u, v, axis)      -> centre (fp32)
//   3. max distance from the axis line, max |height|    -> R, t (rounded up)
// All of it in the mesh's SCALED coordinates (x sigma, a power of two with max |coordinate| * sigma in [0.5, 1): the fp32
// stage can neither overflow nor go denormal whatever the units of the mesh).
// Everything is exact min/max arithmetic in fp64 against the fp32 values the kernel will read, so the cylinder
// contains the fp64 geometry; the kernel's own fp32 rounding is covered by the slack in traverse.h: lane_init.
#pragma once
#ifdef IGLB200_BUILDER_TU
#include "wn.h"

namespace pill {
using fwn::EMPTY; using fwn::INTERNAL_BIT; using fwn::PillNode;

static constexpr uint32_t BIG = 256;   // clusters with more faces get a block

struct Acc1 { double n[3], area, lo[3], hi[3]; };
struct Acc2 { double lo[3], hi[3]; };
struct Acc3 { double lat2, zabs; };
struct Frame { double a[3], u[3], v[3]; float af[3]; };

DEVFN void acc1_init(Acc1 &A) { for (int d = 0; d < 3; ++d) { A.n[d] = 0.0; A.lo[d] = INFINITY; A.hi[d] = -INFINITY; } A.area = 0.0; }
DEVFN void acc1_face(Acc1 &A, const double *tt, double sg) {
  double t[9];
  for (int k = 0; k < 9; ++k) t[k] = tt[k] * sg;
  const double e1[3] = {t[3] - t[0], t[4] - t[1], t[5] - t[2]}, e2[3] = {t[6] - t[0], t[7] - t[1], t[8] - t[2]};
  const double c[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
  const double l = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  if (l == l && l < INFINITY) { A.n[0] += c[0]; A.n[1] += c[1]; A.n[2] += c[2]; A.area += l; }
  for (int k = 0; k < 3; ++k) for (int d = 0; d < 3; ++d) { A.lo[d] = fmin(A.lo[d], t[3 * k + d]); A.hi[d] = fmax(A.hi[d], t[3 * k + d]); }
}
DEVFN void acc1_merge(Acc1 &A, const Acc1 &B) {
  for (int d = 0; d < 3; ++d) { A.n[d] += B.n[d]; A.lo[d] = fmin(A.lo[d], B.lo[d]); A.hi[d] = fmax(A.hi[d], B.hi[d]); }
  A.area += B.area;
}
DEVFN Frame make_frame(const Acc1 &A) {
  Frame f;
  double n[3] = {A.n[0], A.n[1], A.n[2]};
  double l = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  if (!(l >= 0.25 * A.area) || !(l > 0.0) || !(l < INFINITY)) {
    // the normals cancel (a closed or strongly curved cluster): the thinnest axis of the extent
    const double e[3] = {A.hi[0] - A.lo[0], A.hi[1] - A.lo[1], A.hi[2] - A.lo[2]};
    const int k = (e[0] <= e[1] && e[0] <= e[2]) ? 0 : (e[1] <= e[2] ? 1 : 2);
    n[0] = n[1] = n[2] = 0.0; n[k] = 1.0; l = 1.0;
  }
  for (int d = 0; d < 3; ++d) f.af[d] = (float)(n[d] / l);
  // the kernel uses the fp32 axis as if it were unit: measure everything against its exact direction
  double a[3] = {(double)f.af[0], (double)f.af[1], (double)f.af[2]};
  const double al = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  for (int d = 0; d < 3; ++d) f.a[d] = a[d] / al;
  const int m = (fabs(f.a[0]) <= fabs(f.a[1]) && fabs(f.a[0]) <= fabs(f.a[2])) ? 0 : (fabs(f.a[1]) <= fabs(f.a[2]) ? 1 : 2);
  double e[3] = {0.0, 0.0, 0.0};
  e[m] = 1.0;
  double u[3] = {f.a[1] * e[2] - f.a[2] * e[1], f.a[2] * e[0] - f.a[0] * e[2], f.a[0] * e[1] - f.a[1] * e[0]};
  const double ul = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  for (int d = 0; d < 3; ++d) f.u[d] = u[d] / ul;
  f.v[0] = f.a[1] * f.u[2] - f.a[2] * f.u[1]; f.v[1] = f.a[2] * f.u[0] - f.a[0] * f.u[2]; f.v[2] = f.a[0] * f.u[1] - f.a[1] * f.u[0];
  return f;
}
DEVFN void acc2_init(Acc2 &A) { for (int d = 0; d < 3; ++d) { A.lo[d] = INFINITY; A.hi[d] = -INFINITY; } }
DEVFN void acc2_face(Acc2 &A, const Frame &f, const double *tt, double sg) {
  double t[9];
  for (int k = 0; k < 9; ++k) t[k] = tt[k] * sg;
  for (int k = 0; k < 3; ++k) {
    const double *x = t + 3 * k;
    const double p[3] = {x[0] * f.u[0] + x[1] * f.u[1] + x[2] * f.u[2], x[0] * f.v[0] + x[1] * f.v[1] + x[2] * f.v[2],
                         x[0] * f.a[0] + x[1] * f.a[1] + x[2] * f.a[2]};
    for (int d = 0; d < 3; ++d) { A.lo[d] = fmin(A.lo[d], p[d]); A.hi[d] = fmax(A.hi[d], p[d]); }
  }
}
DEVFN void acc2_merge(Acc2 &A, const Acc2 &B) { for (int d = 0; d < 3; ++d) { A.lo[d] = fmin(A.lo[d], B.lo[d]); A.hi[d] = fmax(A.hi[d], B.hi[d]); } }
DEVFN void centre_of(const Frame &f, const Acc2 &A, float cf[3]) {
  const double m[3] = {0.5 * (A.lo[0] + A.hi[0]), 0.5 * (A.lo[1] + A.hi[1]), 0.5 * (A.lo[2] + A.hi[2])};
  for (int d = 0; d < 3; ++d) cf[d] = (float)(m[0] * f.u[d] + m[1] * f.v[d] + m[2] * f.a[d]);
}
DEVFN void acc3_face(Acc3 &A, const Frame &f, const float cf[3], const double *tt, double sg) {
  double t[9];
  for (int k = 0; k < 9; ++k) t[k] = tt[k] * sg;
  for (int k = 0; k < 3; ++k) {
    const double w[3] = {t[3 * k] - (double)cf[0], t[3 * k + 1] - (double)cf[1], t[3 * k + 2] - (double)cf[2]};
    const double z = w[0] * f.a[0] + w[1] * f.a[1] + w[2] * f.a[2];
    const double l[3] = {w[0] - z * f.a[0], w[1] - z * f.a[1], w[2] - z * f.a[2]};
    A.lat2 = fmax(A.lat2, l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    A.zabs = fmax(A.zabs, fabs(z));
  }
}
// conservative fp32 of a non-negative fp64 extent: its own rounding in fp64 (a few ulp of the coordinates) and the cast
DEVFN float extent_up(double x, double scale) {
  const double y = x * (1.0 + 1e-9) + scale * 1e-14;
  float f = (float)y;
  if ((double)f < y) f = plat::u2f(plat::f2u(f) + 1u);   // next float up (y >= 0)
  return f;
}
DEVFN float f_dn(double d) { float f = (float)d; if ((double)f > d) f = plat::u2f(f > 0.f ? plat::f2u(f) - 1u : (f < 0.f ? plat::f2u(f) + 1u : 0x80000001u)); return f; }
DEVFN float f_upw(double d) { float f = (float)d; if ((double)f < d) f = plat::u2f(f > 0.f ? plat::f2u(f) + 1u : (f < 0.f ? plat::f2u(f) - 1u : 0x00000001u)); return f; }
// A lane keeps whichever bound has the smaller surface area — the measure of how often a far query's sphere reaches a
// convex body (mean shadow = area / 4): the cylinder for clusters that are flat in SOME direction (smooth surfaces, single
// faces), the axis-aligned box for needle-like faces and for clusters as thick as they are wide (noise, the top of the tree).
DEVFN void store_lane(PillNode *nodes, uint32_t cluster, const Frame &f, const float cf[3], const Acc3 &A, double scale, const Acc1 &ext,
                      uint32_t *kinds) {
  PillNode &n = nodes[cluster >> 2];
  const int l = (int)(cluster & 3u);
  {
    const double e[3] = {ext.hi[0] - ext.lo[0], ext.hi[1] - ext.lo[1], ext.hi[2] - ext.lo[2]};
    double sbox = 2.0 * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0]);
    const double R0 = sqrt(A.lat2), t0 = A.zabs;
    double spill = 6.283185307179586 * R0 * (R0 + 2.0 * t0);
    const int mode = (int)scale;
    if (mode == 0) sbox = INFINITY;
    if (mode == 3) sbox = -1.0;
    if (mode == 2) {
      // far-field measure: a body of lateral radius R and half-thickness t about the surface is reached by the sphere of a
      // query at distance D when it lies within R + sqrt(2 D t) of the foot point; D = a quarter of the scene (scaled: 0.5)
      const double D = 0.5;
      const double tb = 0.5 * (fabs(f.a[0]) * e[0] + fabs(f.a[1]) * e[1] + fabs(f.a[2]) * e[2]);
      const double rb2 = 0.25 * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) - tb * tb;
      const double rb = sqrt(rb2 > 0.0 ? rb2 : 0.0);
      sbox = (rb + sqrt(2.0 * D * tb)); sbox *= sbox;
      spill = (R0 + sqrt(2.0 * D * t0)); spill *= spill;
    }
    if (sbox < spill && e[0] == e[0] && e[1] == e[1] && e[2] == e[2]) {
      ((float *)&n.cx)[l] = f_dn(ext.lo[0]); ((float *)&n.cy)[l] = f_dn(ext.lo[1]); ((float *)&n.cz)[l] = f_dn(ext.lo[2]);
      ((float *)&n.nx)[l] = f_upw(ext.hi[0]); ((float *)&n.ny)[l] = f_upw(ext.hi[1]); ((float *)&n.nz)[l] = f_upw(ext.hi[2]);
      ((float *)&n.R)[l] = 0.f; ((float *)&n.t)[l] = 0.f;
      kinds[cluster] = 1u;
      return;
    }
  }
  kinds[cluster] = 0u;
  ((float *)&n.cx)[l] = cf[0]; ((float *)&n.cy)[l] = cf[1]; ((float *)&n.cz)[l] = cf[2];
  ((float *)&n.nx)[l] = f.af[0]; ((float *)&n.ny)[l] = f.af[1]; ((float *)&n.nz)[l] = f.af[2];
  float R = extent_up(sqrt(A.lat2), scale), t = extent_up(A.zabs, scale);
  if (!(R == R) || !(t == t)) { R = INFINITY; t = INFINITY; }   // non-finite geometry: never pruned
  ((float *)&n.R)[l] = R; ((float *)&n.t)[l] = t;
}

struct InitNodes {   // children, lane masks, absent lanes
  const uint32_t *nodes; PillNode *out;
  DEVFN void operator()(int64_t i) const {
    PillNode n;
    uint32_t ch[4], imask = 0, tmask = 0;
    for (int l = 0; l < 4; ++l) {
      ch[l] = nodes[4 * i + l];
      if (ch[l] == EMPTY) continue;
      if (ch[l] & INTERNAL_BIT) imask |= 1u << l; else tmask |= 1u << l;
    }
    n.child = make_uint4(ch[0], ch[1], ch[2], ch[3]);
    n.pad = make_uint4(imask, tmask, 0u, 0u);
    n.cx = n.cy = n.cz = n.nx = n.ny = n.nz = n.R = make_float4(0.f, 0.f, 0.f, 0.f);
    n.t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    out[i] = n;
  }
};

struct SetKinds {   // pad.z = lanes bounded by a box
  const uint32_t *kinds; PillNode *out;
  DEVFN void operator()(int64_t i) const {
    out[i].pad.z = (kinds[4 * i] & 1u) | ((kinds[4 * i + 1] & 1u) << 1) | ((kinds[4 * i + 2] & 1u) << 2) | ((kinds[4 * i + 3] & 1u) << 3);
  }
};

struct SmallClusters {   // one thread per cluster of <= BIG faces
  const double *tri64; const uint32_t *leafface, *start, *cnt; PillNode *out; double sigma; uint32_t *kinds; int mode;
  DEVFN void operator()(int64_t c) const {
    const uint32_t n = cnt[c];
    if (n == 0 || n > BIG) return;
    const uint32_t s = start[c];
    Acc1 a1; acc1_init(a1);
    for (uint32_t k = 0; k < n; ++k) acc1_face(a1, tri64 + 9 * (size_t)leafface[s + k], sigma);
    const Frame f = make_frame(a1);
    Acc2 a2; acc2_init(a2);
    for (uint32_t k = 0; k < n; ++k) acc2_face(a2, f, tri64 + 9 * (size_t)leafface[s + k], sigma);
    float cf[3]; centre_of(f, a2, cf);
    Acc3 a3{0.0, 0.0};
    for (uint32_t k = 0; k < n; ++k) acc3_face(a3, f, cf, tri64 + 9 * (size_t)leafface[s + k], sigma);
    store_lane(out, (uint32_t)c, f, cf, a3, (double)mode, a1, kinds);
  }
};
struct FlagBig { const uint32_t *cnt; uint32_t *flag; int64_t nc; DEVFN void operator()(int64_t c) const { flag[c] = (c < nc && cnt[c] > BIG) ? 1u : 0u; } };
struct ListBig { const uint32_t *flag, *scan; uint32_t *list; DEVFN void operator()(int64_t c) const { if (flag[c]) list[scan[c]] = (uint32_t)c; } };

struct BigClusters {   // one block per cluster of > BIG faces (fixed reduction tree: deterministic)
  const double *tri64; const uint32_t *leafface, *start, *cnt, *list; PillNode *out; double sigma; uint32_t *kinds; int mode;
#ifdef IGLB200_EMUL
  DEVFN void tile(uint32_t k) const {
    const uint32_t c = list[k], n = cnt[c], s = start[c];
    Acc1 a1; acc1_init(a1);
    for (uint32_t e = 0; e < n; ++e) acc1_face(a1, tri64 + 9 * (size_t)leafface[s + e], sigma);
    const Frame f = make_frame(a1);
    Acc2 a2; acc2_init(a2);
    for (uint32_t e = 0; e < n; ++e) acc2_face(a2, f, tri64 + 9 * (size_t)leafface[s + e], sigma);
    float cf[3]; centre_of(f, a2, cf);
    Acc3 a3{0.0, 0.0};
    for (uint32_t e = 0; e < n; ++e) acc3_face(a3, f, cf, tri64 + 9 * (size_t)leafface[s + e], sigma);
    store_lane(out, c, f, cf, a3, (double)mode, a1, kinds);
  }
#else
  template <class A, class M> static __device__ A block_reduce(A v, A *sh, M merge) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (unsigned o = blockDim.x / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) merge(sh[threadIdx.x], sh[threadIdx.x + o]);
      __syncthreads();
    }
    const A r = sh[0];
    __syncthreads();
    return r;
  }
  __device__ void tile(uint32_t k) const {
    __shared__ Acc1 sh1[256];
    Acc2 *sh2 = (Acc2 *)sh1; Acc3 *sh3 = (Acc3 *)sh1;
    const uint32_t c = list[k], n = cnt[c], s = start[c];
    Acc1 a1; acc1_init(a1);
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) acc1_face(a1, tri64 + 9 * (size_t)leafface[s + e], sigma);
    a1 = block_reduce(a1, sh1, [](Acc1 &x, const Acc1 &y) { acc1_merge(x, y); });
    const Frame f = make_frame(a1);
    Acc2 a2; acc2_init(a2);
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) acc2_face(a2, f, tri64 + 9 * (size_t)leafface[s + e], sigma);
    a2 = block_reduce(a2, sh2, [](Acc2 &x, const Acc2 &y) { acc2_merge(x, y); });
    float cf[3]; centre_of(f, a2, cf);
    Acc3 a3{0.0, 0.0};
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) acc3_face(a3, f, cf, tri64 + 9 * (size_t)leafface[s + e], sigma);
    a3 = block_reduce(a3, sh3, [](Acc3 &x, const Acc3 &y) { x.lat2 = fmax(x.lat2, y.lat2); x.zabs = fmax(x.zabs, y.zabs); });
    if (threadIdx.x == 0) store_lane(out, c, f, cf, a3, (double)mode, a1, kinds);
  }
#endif
};

// `cnt`, `start`, `leafface`: the cluster ranges (4 per node) in depth-first leaf order
inline void build(const fwn::Tree &tree, const double *tri64, const uint32_t *cnt, const uint32_t *start, const uint32_t *leafface,
                  double sigma, PillNode *out, stream_t s) {
  using plat::DevBuf;
  const int64_t nn = tree.nnodes, nc = 4 * nn;
  if (nn == 0) return;
  const char *mode_env = std::getenv("IGLB200_BOUND_MODE");   // tuning aid: 0 cylinders only, 1 smaller surface area (default), 2 far-field measure, 3 boxes only
  const int mode = mode_env ? std::atoi(mode_env) : 1;
  DevBuf<uint32_t> kinds((size_t)nc);
  plat::dev_memset(kinds.p, 0, (size_t)4 * nc, s);
  plat::for_n(nn, InitNodes{tree.nodes.p, out}, s);
  plat::for_n(nc, SmallClusters{tri64, leafface, start, cnt, out, sigma, kinds.p, mode}, s);
  DevBuf<uint32_t> flag((size_t)nc + 1), scan((size_t)nc + 1), list((size_t)nc + 1);
  plat::for_n(nc + 1, FlagBig{cnt, flag.p, nc}, s);
  plat::exclusive_scan_u32(flag.p, scan.p, nc + 1, s);
  plat::for_n(nc, ListBig{flag.p, scan.p, list.p}, s);
  plat::tiles_dyn(scan.p + nc, BigClusters{tri64, leafface, start, cnt, list.p, out, sigma, kinds.p, mode}, s);
  plat::for_n(nn, SetKinds{kinds.p, out}, s);
  plat::sync(s);
}

}  // namespace pill
#endif
