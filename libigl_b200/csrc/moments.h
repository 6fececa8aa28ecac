// Bottom-up per-node data for both traversals, one thread per node, one launch
// per tree level (deepest first).
//
//  * order-2 Taylor moments of the soup fast winding number, restating the fp32
//    arithmetic of UT_SolidAngle::init's functors
//      triangle: /root/reference/include/igl/FastWindingNumberForSoups.h:6668-6890
//      node:     :6892-7118
//    operation by operation (same association order, reciprocal-multiply where
//    the reference's vector/scalar operator uses it (:3423-3452), no FMA), so
//    the cluster centre P and radius^2 that drive the descend test
//    (:7185-7191) come out bit-identical;
//  * conservative fp32 child boxes (rounded outward from the fp64 vertices) for
//    the closest-triangle traversal.
//
// Compiled with -fmad=false.
#pragma once
#include "fwn_build.h"

namespace fwn {

// 23 rows x 4 child lanes, row order = the reference's BoxData (:6489-6519)
enum Row {
  R_MAXP2 = 0, R_PX, R_PY, R_PZ, R_NX, R_NY, R_NZ, R_NXX, R_NYY, R_NZZ, R_NXY_NYX, R_NYZ_NZY, R_NZX_NXZ,
  R_NXXX, R_NYYY, R_NZZZ, R_SUMPERM, R_2NXXY_NYXX, R_2NXXZ_NZXX, R_2NYYZ_NZYY, R_2NYYX_NXYY, R_2NZZX_NXZZ,
  R_2NZZY_NYZZ, R_COUNT
};

struct alignas(16) FwnNode {  // 384 B
  uint4 child;
  float4 d[R_COUNT];
};
struct alignas(128) CpNode {  // 128 B = one L2 line: child boxes SoA over the 4 lanes
  float4 lox, loy, loz, hix, hiy, hiz;
  uint4 child;
  uint4 pad;  // .x = bitmask of lanes that are internal nodes, .y = lanes that are faces
};

// Node of the closest-triangle search: every child lane (an internal cluster or a single face) is bounded by an
// ORIENTED CYLINDER ("pill box"): centre c, unit axis n, radius R about the axis, half-height t along it —
//     lower bound of the distance from p:  sqrt( max(|(p-c).n| - t, 0)^2 + max(|(p-c) - ((p-c).n) n| - R, 0)^2 ).
// The axis is the cluster's area-weighted mean normal, so t is the cluster's deviation from its mean plane (curvature,
// noise) — not, as with an axis-aligned box, the height of its tilted faces.  For a query at distance d the candidates
// are the clusters within ~sqrt(2 d t) of the foot point: independent of the mesh resolution for smooth surfaces
// (boxes: t ~ face size).  A face lane is the same bound with t ~ 0: a disc.  8 rows x 4 lanes + children + masks = 160 B
// = five 32-byte sectors.  Absent lanes carry t = -inf (bound = +inf).
struct alignas(32) PillNode {
  float4 cx, cy, cz, nx, ny, nz, R, t;
  uint4 child;
  uint4 pad;  // .x = bitmask of lanes that are internal nodes, .y = lanes that are faces, .z = lanes bounded by an axis-aligned
              // box instead (rows cx,cy,cz = lo, nx,ny,nz = hi): whichever has the smaller surface area (pill.h)
};
static_assert(sizeof(PillNode) == 160, "PillNode layout");

// tri32: 3 float4 per face, a,b,c = (float)vertex in .xyz — the reference's V.cast<float>()
// (signed_distance.cpp:66); the winding number and the moments read exactly these.
static constexpr int TRI32_STRIDE = 3;
// triq: 4 float4 per face for the fp32 stage of the closest-triangle search (traverse.h):
// the face in its own orthonormal frame, derived from the fp64 vertices:
//   [o.xyz | L] [u.xyz | cx] [v.xyz | cy] [n.xyz | unused]
// o = start of the longest edge (as float), u along that edge, n the unit normal, v = n x u;
// in that frame the vertices are (0,0), (L,0), (cx,cy) with cy >= 0 (cy == 0: segment / point).
static constexpr int TRIQ_STRIDE = 4;

struct Local {  // what a subtree hands to its parent (LocalData :6611-6642)
  Box3 box;
  float P[3], areaP[3], N[3], area;
  float Nii[3], Nxy, Nyx, Nyz, Nzy, Nzx, Nxz;
  float Niii[3], sumperm, n2xxy, n2xxz, n2yyz, n2yyx, n2zzx, n2zzy;
  Box3 cpbox;  // conservative box of the fp64 geometry
};

#ifdef IGLB200_BUILDER_TU
DEVFN float f_round_down(double d) {
#if defined(__CUDA_ARCH__)
  return __double2float_rd(d);
#else
  float f = (float)d;
  if ((double)f > d) f = std::nextafterf(f, -INFINITY);
  return f;
#endif
}
DEVFN float f_round_up(double d) {
#if defined(__CUDA_ARCH__)
  return __double2float_ru(d);
#else
  float f = (float)d;
  if ((double)f < d) f = std::nextafterf(f, INFINITY);
  return f;
#endif
}

struct F3 { float v[3]; };

// compute_integrals lambda (:6765-6838); a,b,c sorted along axis i
DEVFN void tri_integrals(const F3 &a, const F3 &b, const F3 &c, const float *P, float *ii, float *ij, float *ik, int i) {
  float oab[3], oac[3], ocb[3];
  for (int d = 0; d < 3; ++d) { oab[d] = b.v[d] - a.v[d]; oac[d] = c.v[d] - a.v[d]; ocb[d] = b.v[d] - c.v[d]; }
  const float t = oab[i] / oac[i];
  const int j = (i == 2) ? 0 : (i + 1);
  const int k = (j == 2) ? 0 : (j + 1);
  const float jdiff = t * oac[j] - oab[j];
  const float kdiff = t * oac[k] - oab[k];
  const float ca0 = (jdiff * oab[k] - kdiff * oab[j]), ca1 = kdiff * oab[i], ca2 = jdiff * oab[i];
  const float cc0 = (jdiff * ocb[k] - kdiff * ocb[j]), cc1 = kdiff * ocb[i], cc2 = jdiff * ocb[i];
  float l2 = ca0 * ca0; l2 += ca1 * ca1; l2 += ca2 * ca2;
  const float area_scale_a = sqrtf(l2);
  l2 = cc0 * cc0; l2 += cc1 * cc1; l2 += cc2 * cc2;
  const float area_scale_c = sqrtf(l2);
  const float Pai = a.v[i] - P[i];
  const float Pci = c.v[i] - P[i];
  const float two3 = (float)(2.0 / 3.0), one3 = (float)(1.0 / 3.0);
  const float int_ii_a = area_scale_a * (0.5f * Pai * Pai + two3 * Pai * oab[i] + 0.25f * oab[i] * oab[i]);
  const float int_ii_c = area_scale_c * (0.5f * Pci * Pci + two3 * Pci * ocb[i] + 0.25f * ocb[i] * ocb[i]);
  *ii = int_ii_a + int_ii_c;
  int jk = j;
  float *integral = ij;
  float diff = jdiff;
  while (true) {
    if (integral) {
      const float obmidj = b.v[jk] + 0.5f * diff;
      const float oabmidj = obmidj - a.v[jk];
      const float ocbmidj = obmidj - c.v[jk];
      const float Paj = a.v[jk] - P[jk];
      const float Pcj = c.v[jk] - P[jk];
      const float int_ij_a = area_scale_a * (0.5f * Pai * Paj + one3 * Pai * oabmidj + one3 * Paj * oab[i] + 0.25f * oab[i] * oabmidj);
      const float int_ij_c = area_scale_c * (0.5f * Pci * Pcj + one3 * Pci * ocbmidj + one3 * Pcj * ocb[i] + 0.25f * ocb[i] * ocbmidj);
      *integral = int_ij_a + int_ij_c;
    }
    if (jk == k) break;
    jk = k;
    integral = ik;
    diff = kdiff;
  }
}

DEVFN void sort_axis(const F3 *values, int axis, int *o) {  // (:6733-6757)
  o[0] = 0; o[1] = 1; o[2] = 2;
  int t;
  if (values[0].v[axis] > values[1].v[axis]) { t = o[0]; o[0] = o[1]; o[1] = t; }
  if (values[o[0]].v[axis] > values[2].v[axis]) { t = o[0]; o[0] = o[2]; o[2] = t; }
  if (values[o[1]].v[axis] > values[o[2]].v[axis]) { t = o[1]; o[1] = o[2]; o[2] = t; }
}

// item functor (:6668-6890)
DEVFN void tri_local(const float4 *tri32, const double *tri64, const TriBox *boxes, uint32_t f, int order, Local &L) {
  F3 values[3];
  for (int k = 0; k < 3; ++k) { const float4 p = tri32[TRI32_STRIDE * (size_t)f + k]; values[k].v[0] = p.x; values[k].v[1] = p.y; values[k].v[2] = p.z; }
  const F3 &a = values[0], &b = values[1], &c = values[2];
  float ab[3], ac[3];
  for (int d = 0; d < 3; ++d) { ab[d] = b.v[d] - a.v[d]; ac[d] = c.v[d] - a.v[d]; }
  L.box = load_box(boxes, f);
  float N[3];
  N[0] = 0.5f * (ab[1] * ac[2] - ab[2] * ac[1]);
  N[1] = 0.5f * (ab[2] * ac[0] - ab[0] * ac[2]);
  N[2] = 0.5f * (ab[0] * ac[1] - ab[1] * ac[0]);
  float area2 = N[0] * N[0]; area2 += N[1] * N[1]; area2 += N[2] * N[2];
  const float area = sqrtf(area2);
  const float third = 1 / 3.0f;  // operator/(T): multiply by the reciprocal (:3446-3449)
  float P[3];
  for (int d = 0; d < 3; ++d) P[d] = ((a.v[d] + b.v[d]) + c.v[d]) * third;
  for (int d = 0; d < 3; ++d) { L.P[d] = P[d]; L.areaP[d] = P[d] * area; L.N[d] = N[d]; }
  L.area = area;
  for (int d = 0; d < 3; ++d) { L.Nii[d] = 0.f; L.Niii[d] = 0.f; }
  L.Nxy = L.Nyx = L.Nyz = L.Nzy = L.Nzx = L.Nxz = 0.f;
  L.sumperm = L.n2xxy = L.n2xxz = L.n2yyz = L.n2yyx = L.n2zzx = L.n2zzy = 0.f;
  // conservative box of the double-precision triangle
  for (int d = 0; d < 3; ++d) {
    const double v0 = tri64[9 * (size_t)f + d], v1 = tri64[9 * (size_t)f + 3 + d], v2 = tri64[9 * (size_t)f + 6 + d];
    L.cpbox.lo[d] = f_round_down(fmin(fmin(v0, v1), v2));
    L.cpbox.hi[d] = f_round_up(fmax(fmax(v0, v1), v2));
  }
  if (order < 2 || area == 0.f) return;
  const float inv_area = 1 / area;
  const float n[3] = {N[0] * inv_area, N[1] * inv_area, N[2] * inv_area};
  int ox[3], oy[3], oz[3];
  sort_axis(values, 0, ox);
  sort_axis(values, 1, oy);
  sort_axis(values, 2, oz);
  const float dx = values[ox[2]].v[0] - values[ox[0]].v[0];
  const float dy = values[oy[2]].v[1] - values[oy[0]].v[1];
  const float dz = values[oz[2]].v[2] - values[oz[0]].v[2];
  float ixx = 0, ixy = 0, iyy = 0, iyz = 0, izz = 0, izx = 0;
  if (dx > 0) tri_integrals(values[ox[0]], values[ox[1]], values[ox[2]], P, &ixx, ((dx >= dy && dy > 0) ? &ixy : nullptr), ((dx >= dz && dz > 0) ? &izx : nullptr), 0);
  if (dy > 0) tri_integrals(values[oy[0]], values[oy[1]], values[oy[2]], P, &iyy, ((dy >= dz && dz > 0) ? &iyz : nullptr), ((dx < dy && dx > 0) ? &ixy : nullptr), 1);
  if (dz > 0) tri_integrals(values[oz[0]], values[oz[1]], values[oz[2]], P, &izz, ((dx < dz && dx > 0) ? &izx : nullptr), ((dy < dz && dy > 0) ? &iyz : nullptr), 2);
  L.Niii[0] = ixx * n[0]; L.Niii[1] = iyy * n[1]; L.Niii[2] = izz * n[2];
  L.sumperm = 2 * (n[0] * iyz + n[1] * izx + n[2] * ixy);
  const float Nxxy = n[0] * ixy, Nxxz = n[0] * izx, Nyyz = n[1] * iyz, Nyyx = n[1] * ixy, Nzzx = n[2] * izx, Nzzy = n[2] * iyz;
  L.n2xxy = 2 * Nxxy + n[1] * ixx;
  L.n2xxz = 2 * Nxxz + n[2] * ixx;
  L.n2yyz = 2 * Nyyz + n[2] * iyy;
  L.n2yyx = 2 * Nyyx + n[0] * iyy;
  L.n2zzx = 2 * Nzzx + n[0] * izz;
  L.n2zzy = 2 * Nzzy + n[1] * izz;
}

DEVFN void set_lane(float4 &r, int lane, float v) { ((float *)&r)[lane] = v; }

// post functor (:6892-7118) + closest-point child boxes
struct NodeMoments {
  const uint32_t *nodes;  // 4 per node
  const float4 *tri32;
  const double *tri64;
  const TriBox *boxes;
  Local *locals;  // per node
  FwnNode *fnodes;
  CpNode *cnodes;
  uint32_t first;  // first node of this level
  int order;
  DEVFN void operator()(int64_t k) const {
    const uint32_t nodei = first + (uint32_t)k;
    uint32_t child[4];
    for (int i = 0; i < 4; ++i) child[i] = nodes[4 * (size_t)nodei + i];
    Local cd[4];
    int nchildren = 0;
    for (int i = 0; i < 4; ++i) {
      if (child[i] == EMPTY) break;
      if (child[i] & INTERNAL_BIT) cd[i] = locals[child[i] & ~INTERNAL_BIT];
      else tri_local(tri32, tri64, boxes, child[i], order, cd[i]);
      ++nchildren;
    }
    FwnNode fn;
    CpNode cn;
    fn.child = make_uint4(child[0], child[1], child[2], child[3]);
    cn.child = fn.child;
    {
      uint32_t imask = 0, tmask = 0;  // lanes holding internal nodes / faces
      for (int i = 0; i < 4; ++i) {
        if (child[i] == EMPTY) continue;
        if (child[i] & INTERNAL_BIT) imask |= 1u << i; else tmask |= 1u << i;
      }
      cn.pad = make_uint4(imask, tmask, 0, 0);
    }
    for (int r = 0; r < R_COUNT; ++r) fn.d[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    Local out;
    float N[3] = {cd[0].N[0], cd[0].N[1], cd[0].N[2]};
    float areaP[3] = {cd[0].areaP[0], cd[0].areaP[1], cd[0].areaP[2]};
    float area = cd[0].area;
    for (int i = 1; i < nchildren; ++i) {
      for (int d = 0; d < 3; ++d) { N[d] += cd[i].N[d]; areaP[d] += cd[i].areaP[d]; }
      area += cd[i].area;
    }
    for (int i = 0; i < nchildren; ++i) {
      set_lane(fn.d[R_NX], i, cd[i].N[0]); set_lane(fn.d[R_NY], i, cd[i].N[1]); set_lane(fn.d[R_NZ], i, cd[i].N[2]);
      set_lane(fn.d[R_PX], i, cd[i].P[0]); set_lane(fn.d[R_PY], i, cd[i].P[1]); set_lane(fn.d[R_PZ], i, cd[i].P[2]);
    }
    for (int d = 0; d < 3; ++d) { out.N[d] = N[d]; out.areaP[d] = areaP[d]; }
    out.area = area;
    Box3 box = cd[0].box, cpbox = cd[0].cpbox;
    for (int i = 1; i < nchildren; ++i) { box_combine(box, cd[i].box); box_combine(cpbox, cd[i].cpbox); }
    float avgP[3];
    if (area > 0) {
      const float inv = 1 / area;
      for (int d = 0; d < 3; ++d) avgP[d] = areaP[d] * inv;
    } else {
      for (int d = 0; d < 3; ++d) avgP[d] = 0.5f * (box.lo[d] + box.hi[d]);
    }
    for (int d = 0; d < 3; ++d) out.P[d] = avgP[d];
    out.box = box;
    out.cpbox = cpbox;
    for (int i = 0; i < 4; ++i) {
      float r2 = INFINITY;
      if (i < nchildren) {
        float m[3];
        for (int d = 0; d < 3; ++d) {
          const float x = cd[i].P[d] - cd[i].box.lo[d], y = cd[i].box.hi[d] - cd[i].P[d];
          m[d] = (x < y) ? y : x;  // SYSmax
        }
        r2 = m[0] * m[0]; r2 += m[1] * m[1]; r2 += m[2] * m[2];
      }
      set_lane(fn.d[R_MAXP2], i, r2);
      const float BIG = INFINITY;
      set_lane(cn.lox, i, i < nchildren ? cd[i].cpbox.lo[0] : BIG); set_lane(cn.hix, i, i < nchildren ? cd[i].cpbox.hi[0] : -BIG);
      set_lane(cn.loy, i, i < nchildren ? cd[i].cpbox.lo[1] : BIG); set_lane(cn.hiy, i, i < nchildren ? cd[i].cpbox.hi[1] : -BIG);
      set_lane(cn.loz, i, i < nchildren ? cd[i].cpbox.lo[2] : BIG); set_lane(cn.hiz, i, i < nchildren ? cd[i].cpbox.hi[2] : -BIG);
    }
    if (order >= 1) {
      for (int d = 0; d < 3; ++d) { out.Nii[d] = cd[0].Nii[d]; out.Niii[d] = cd[0].Niii[d]; }
      out.Nxy = out.Nyx = out.Nyz = out.Nzy = out.Nzx = out.Nxz = 0.f;
      out.sumperm = cd[0].sumperm;
      out.n2xxy = cd[0].n2xxy; out.n2xxz = cd[0].n2xxz; out.n2yyz = cd[0].n2yyz;
      out.n2yyx = cd[0].n2yyx; out.n2zzx = cd[0].n2zzx; out.n2zzy = cd[0].n2zzy;
      for (int i = 1; i < nchildren; ++i) {
        for (int d = 0; d < 3; ++d) { out.Nii[d] += cd[i].Nii[d]; out.Niii[d] += cd[i].Niii[d]; }
        out.sumperm += cd[i].sumperm;
        out.n2xxy += cd[i].n2xxy; out.n2xxz += cd[i].n2xxz; out.n2yyz += cd[i].n2yyz;
        out.n2yyx += cd[i].n2yyx; out.n2zzx += cd[i].n2zzx; out.n2zzy += cd[i].n2zzy;
      }
      for (int i = 0; i < nchildren; ++i) {
        set_lane(fn.d[R_NXX], i, cd[i].Nii[0]); set_lane(fn.d[R_NYY], i, cd[i].Nii[1]); set_lane(fn.d[R_NZZ], i, cd[i].Nii[2]);
        set_lane(fn.d[R_NXY_NYX], i, cd[i].Nxy + cd[i].Nyx);
        set_lane(fn.d[R_NYZ_NZY], i, cd[i].Nyz + cd[i].Nzy);
        set_lane(fn.d[R_NZX_NXZ], i, cd[i].Nzx + cd[i].Nxz);
        set_lane(fn.d[R_NXXX], i, cd[i].Niii[0]); set_lane(fn.d[R_NYYY], i, cd[i].Niii[1]); set_lane(fn.d[R_NZZZ], i, cd[i].Niii[2]);
        set_lane(fn.d[R_SUMPERM], i, cd[i].sumperm);
        set_lane(fn.d[R_2NXXY_NYXX], i, cd[i].n2xxy); set_lane(fn.d[R_2NXXZ_NZXX], i, cd[i].n2xxz);
        set_lane(fn.d[R_2NYYZ_NZYY], i, cd[i].n2yyz); set_lane(fn.d[R_2NYYX_NXYY], i, cd[i].n2yyx);
        set_lane(fn.d[R_2NZZX_NXZZ], i, cd[i].n2zzx); set_lane(fn.d[R_2NZZY_NYZZ], i, cd[i].n2zzy);
      }
      for (int i = 0; i < nchildren; ++i) {
        const Local &c = cd[i];
        float dp[3];
        for (int d = 0; d < 3; ++d) dp[d] = c.P[d] - avgP[d];
        const float *Nc = c.N;
        for (int d = 0; d < 3; ++d) out.Nii[d] += Nc[d] * dp[d];
        const float Nxy = c.Nxy + Nc[0] * dp[1];
        const float Nyx = c.Nyx + Nc[1] * dp[0];
        const float Nyz = c.Nyz + Nc[1] * dp[2];
        const float Nzy = c.Nzy + Nc[2] * dp[1];
        const float Nzx = c.Nzx + Nc[2] * dp[0];
        const float Nxz = c.Nxz + Nc[0] * dp[2];
        out.Nxy += Nxy; out.Nyx += Nyx; out.Nyz += Nyz; out.Nzy += Nzy; out.Nzx += Nzx; out.Nxz += Nxz;
        if (order >= 2) {
          for (int d = 0; d < 3; ++d) out.Niii[d] += 2.0f * dp[d] * c.Nii[d] + dp[d] * dp[d] * Nc[d];
          out.sumperm += (dp[0] * (Nyz + Nzy) + dp[1] * (Nzx + Nxz) + dp[2] * (Nxy + Nyx));
          out.n2xxy += 2 * (dp[1] * c.Nii[0] + dp[0] * c.Nxy + Nc[0] * dp[0] * dp[1]) + 2 * c.Nyx * dp[0] + Nc[1] * dp[0] * dp[0];
          out.n2xxz += 2 * (dp[2] * c.Nii[0] + dp[0] * c.Nxz + Nc[0] * dp[0] * dp[2]) + 2 * c.Nzx * dp[0] + Nc[2] * dp[0] * dp[0];
          out.n2yyz += 2 * (dp[2] * c.Nii[1] + dp[1] * c.Nyz + Nc[1] * dp[1] * dp[2]) + 2 * c.Nzy * dp[1] + Nc[2] * dp[1] * dp[1];
          out.n2yyx += 2 * (dp[0] * c.Nii[1] + dp[1] * c.Nyx + Nc[1] * dp[1] * dp[0]) + 2 * c.Nxy * dp[1] + Nc[0] * dp[1] * dp[1];
          out.n2zzx += 2 * (dp[0] * c.Nii[2] + dp[2] * c.Nzx + Nc[2] * dp[2] * dp[0]) + 2 * c.Nxz * dp[2] + Nc[0] * dp[2] * dp[2];
          out.n2zzy += 2 * (dp[1] * c.Nii[2] + dp[2] * c.Nzy + Nc[2] * dp[2] * dp[1]) + 2 * c.Nyz * dp[2] + Nc[1] * dp[2] * dp[2];
        }
      }
    } else {
      for (int d = 0; d < 3; ++d) { out.Nii[d] = 0.f; out.Niii[d] = 0.f; }
      out.Nxy = out.Nyx = out.Nyz = out.Nzy = out.Nzx = out.Nxz = 0.f;
      out.sumperm = out.n2xxy = out.n2xxz = out.n2yyz = out.n2yyx = out.n2zzx = out.n2zzy = 0.f;
    }
    locals[nodei] = out;
    fnodes[nodei] = fn;
    cnodes[nodei] = cn;
  }
};

#endif  // IGLB200_BUILDER_TU

}  // namespace fwn
