// Voxel grids on either side of the signed-distance path, device resident end to end
// (mesh -> SDF grid -> iso-surface without a host round trip):
//   igl::grid            /root/reference/include/igl/grid.cpp:12-45
//   igl::voxel_grid      /root/reference/include/igl/voxel_grid.cpp:14-66 (box overload; :68-88 = box from V)
//   igl::flood_fill      /root/reference/include/igl/flood_fill.cpp:11-98
//   igl::marching_cubes  /root/reference/include/igl/marching_cubes.cpp:34-112 + march_cube.cpp:24-142
//   igl::offset_surface  /root/reference/include/igl/offset_surface.cpp:18-48 (composed by the host wrappers)
//
// Not a port.  The reference's flood fill and marching cubes are single sequential sweeps whose
// results depend on the sweep order (a hash map hands out vertex ids in order of first visit; the
// flood value is carried along the scan).  Here the order is reproduced in closed form:
//   * flood_fill: every signed cell floods back over the NaN cells since the previous signed one,
//     and what is left keeps the carried value.  So a NaN cell ends up with (a) the next signed
//     value of its row, else the row's last one; (b) in a row without any, the first signed value
//     of the next signed row of the sheet, else of the sheet's last signed row; (c) in a sheet
//     without any, the same rule over sheets.  One pass over rows, one over sheets, one fill.
//   * marching_cubes: a cut edge gets its vertex from the FIRST cube (x fastest) that contains it;
//     that "owner" is a closed-form function of the edge.  Every cube counts the cut edges it owns,
//     an exclusive scan turns counts into vertex ids (= the reference's insertion order: cubes in
//     scan order, edges 0..11 inside a cube), and triangles look their three ids up at the owners.
//     The vertex is interpolated from the owner's first endpoint exactly as the reference does
//     (un-fused: GV(i) + t*(GV(j) - GV(i)), t = (iso - a)/(b - a)), so V and F are bit-identical.
#pragma once
#include "platform.h"
#include "traverse.h"
#include "mc_tables.h"
#include <stdexcept>
#include <cmath>
#include <limits>

struct iglb200_grid {
  int side[3] = {0, 0, 0};
  int64_t n = 0;
  plat::DevBuf<double> GV;  // column-major 3 x n (the layout the query kernels read)
  plat::DevBuf<double> S;   // n values, x fastest
  bool has_S = false;
};
struct iglb200_surface {
  int64_t nV = 0, nF = 0;
  plat::DevBuf<double> V;    // nV x 3 row-major
  plat::DevBuf<int32_t> F;   // nF x 3 row-major
  plat::DevBuf<int32_t> E;   // nV x 2: the grid edge (i < j) each vertex lies on = its E2V key (march_cube.cpp:52-60)
};

namespace grid {
using fwn::dadd; using fwn::dsub; using fwn::dmul; using fwn::ddiv;

// ------------------------------------------------------------------ voxel_grid
struct VoxelParams { int side[3]; double ratio[3], off[3]; };

// host arithmetic of voxel_grid.cpp:22-65 (a dozen scalar operations)
inline VoxelParams voxel_params(const double bmin[3], const double bmax[3], int in_s, int pad_count) {
  VoxelParams vp;
  double diag[3];
  for (int c = 0; c < 3; ++c) {
    diag[c] = bmax[c] - bmin[c];
    if (!std::isfinite(bmin[c]) || !std::isfinite(bmax[c])) throw std::invalid_argument("voxel_grid: box is not finite");
  }
  if (pad_count < 0) throw std::invalid_argument("voxel_grid: pad_count < 0");
  if (!(in_s > pad_count * 2 + 1)) throw std::invalid_argument("voxel_grid: s should be > 2*pad_count+1");   // assert at :31
  int si = 0;
  for (int c = 1; c < 3; ++c) if (diag[c] > diag[si]) si = c;   // maxCoeff: first maximum (:24)
  const double s_len = diag[si];
  if (!(s_len > 0.0)) throw std::invalid_argument("voxel_grid: empty box");
  const double s = in_s - 2 * pad_count;
  vp.side[si] = (int)s;
  for (int i = 0; i < 3; ++i)
    if (i != si) vp.side[i] = (int)std::ceil(s * (bmax[i] - bmin[i]) / s_len);
  for (int i = 0; i < 3; ++i) vp.side[i] += 2 * pad_count;
  double A[3];
  for (int c = 0; c < 3; ++c) {
    const double ps = (double)pad_count / ((double)vp.side[c] - 1.);
    A[c] = diag[c] / (1.0 - 2. * ps);
  }
  int ai = 0;
  for (int c = 1; c < 3; ++c) if (A[c] > A[ai]) ai = c;
  for (int c = 0; c < 3; ++c) vp.ratio[c] = A[ai] * ((double)vp.side[c] - 1.0) / (double)(vp.side[ai] - 1.0);
  // GV.colwise().mean() (:64): along axis c every value k/(side-1)*ratio occurs n/side times, so the
  // column mean is the mean of those side values (the reference sums all n rows; Eigen's reduction
  // order is unspecified, the two agree to ~1e-16 relative)
  for (int c = 0; c < 3; ++c) {
    double sum = 0.0;
    for (int k = 0; k < vp.side[c]; ++k) sum += ((double)k / (double)(vp.side[c] - 1)) * vp.ratio[c];
    vp.off[c] = (bmin[c] + bmax[c]) / 2.0 - sum / (double)vp.side[c];
  }
  return vp;
}

struct FillGrid {   // grid.cpp:22-44 then voxel_grid.cpp:63,66: (sub/(res-1)) * ratio + offset
  int sx, sy; int64_t n;
  double dx, dy, dz;   // side - 1
  double rx, ry, rz, ox, oy, oz;
  double *GV;
  DEVFN void operator()(int64_t g) const {
    const int x = (int)(g % sx), y = (int)((g / sx) % sy), z = (int)(g / ((int64_t)sx * sy));
    GV[g] = dadd(dmul(ddiv((double)x, dx), rx), ox);
    GV[n + g] = dadd(dmul(ddiv((double)y, dy), ry), oy);
    GV[2 * n + g] = dadd(dmul(ddiv((double)z, dz), rz), oz);
  }
};

inline iglb200_grid *create(const double bmin[3], const double bmax[3], int s, int pad_count, stream_t st) {
  const VoxelParams vp = voxel_params(bmin, bmax, s, pad_count);
  int64_t n = 1;
  for (int c = 0; c < 3; ++c) {
    if (vp.side[c] < 1) throw std::invalid_argument("voxel_grid: empty side");
    n *= vp.side[c];
    if (n > 0x7FFFFFFF) throw std::invalid_argument("voxel_grid: more than 2^31-1 grid vertices (the reference indexes them with int)");
  }
  iglb200_grid *g = new iglb200_grid;
  try {
    for (int c = 0; c < 3; ++c) g->side[c] = vp.side[c];
    g->n = n;
    g->GV.alloc((size_t)3 * n);
    g->S.alloc((size_t)n);
    plat::Span sp("voxel_grid", st);
    plat::for_n(n, FillGrid{vp.side[0], vp.side[1], n, (double)(vp.side[0] - 1), (double)(vp.side[1] - 1), (double)(vp.side[2] - 1),
                            vp.ratio[0], vp.ratio[1], vp.ratio[2], vp.off[0], vp.off[1], vp.off[2], g->GV.p}, st);
    plat::sync(st);
  } catch (...) { delete g; throw; }
  return g;
}

// ------------------------------------------------------------------ flood_fill
DEVFN bool is_nan(double v) { return v != v; }
DEVFN double qnan() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7FF8000000000000ll);
#else
  return std::numeric_limits<double>::quiet_NaN();
#endif
}

struct FloodRows {   // per row: a NaN cell takes the next signed value of its row, else the last one
  int nx; double *S; double *rowval;   // rowval: first signed value of the row, NaN if the row has none
  DEVFN void operator()(int64_t r) const {
    double *row = S + r * nx;
    double nxt = qnan();
    for (int x = nx - 1; x >= 0; --x) {
      const double v = row[x];
      if (!is_nan(v)) nxt = v;
      else if (!is_nan(nxt)) row[x] = nxt;
    }
    rowval[r] = nxt;
    if (is_nan(nxt)) return;
    double last = nxt;
    for (int x = 0; x < nx; ++x) {
      const double v = row[x];
      if (is_nan(v)) row[x] = last; else last = v;
    }
  }
};
// the same rule one level up: for an entry without a value, the next valued entry's value, else the last valued one's
DEVFN void fill_gaps(const double *val, double *fill, int64_t n, double &firstval) {
  const double none = qnan();
  double nxt = none;
  for (int64_t i = n - 1; i >= 0; --i) { if (!is_nan(val[i])) nxt = val[i]; fill[i] = is_nan(val[i]) ? nxt : none; }
  double last = none;
  for (int64_t i = 0; i < n; ++i) { if (!is_nan(val[i])) last = val[i]; else if (is_nan(fill[i])) fill[i] = last; }
  firstval = nxt;   // after the backward sweep: value of the first valued entry (NaN if none)
}
struct FloodSheets {   // per sheet over its rows
  int ny; const double *rowval; double *rowfill; double *sheetval;
  DEVFN void operator()(int64_t z) const { fill_gaps(rowval + z * ny, rowfill + z * ny, ny, sheetval[z]); }
};
struct FloodVolume {   // one thread over the sheets
  int nz; const double *sheetval; double *sheetfill;
  DEVFN void operator()(int64_t) const { double unused; fill_gaps(sheetval, sheetfill, nz, unused); }
};
struct FloodFill {     // rows without any signed cell take their row's, else their sheet's fill value
  int nx, ny; double *S; const double *rowval, *rowfill, *sheetfill;
  DEVFN void operator()(int64_t g) const {
    const int64_t r = g / nx;
    if (!is_nan(rowval[r])) return;
    double v = rowfill[r];
    if (is_nan(v)) v = sheetfill[r / ny];
    if (!is_nan(v)) S[g] = v;
  }
};

inline void flood_fill(const int side[3], double *S, stream_t st) {
  const int nx = side[0], ny = side[1], nz = side[2];
  const int64_t rows = (int64_t)ny * nz, n = rows * nx;
  if (n <= 0) return;
  plat::DevBuf<double> rowval((size_t)rows), rowfill((size_t)rows), sheetval((size_t)nz), sheetfill((size_t)nz);
  { plat::Span sp("flood_fill", st);
  plat::for_n(rows, FloodRows{nx, S, rowval.p}, st);
  plat::for_n(nz, FloodSheets{ny, rowval.p, rowfill.p, sheetval.p}, st);
  plat::for_n(1, FloodVolume{nz, sheetval.p, sheetfill.p}, st);
  plat::for_n(n, FloodFill{nx, ny, S, rowval.p, rowfill.p, sheetfill.p}, st); }
  plat::sync(st);
}

// ------------------------------------------------------------------ marching_cubes
// cube corners and edges in the numbering of marching_cubes.cpp:52 (ioffset) and the case tables:
// corner c sits at (CX,CY,CZ)[c]; edge e joins corners EA[e] -> EB[e] (the direction matters: the
// vertex is interpolated from EA).
#ifdef IGLB200_EMUL
#define MC_TABLE static const
#else
#define MC_TABLE static __device__ const
#endif
MC_TABLE uint64_t MC_TRI[256] = {IGLB200_MC_TRI_WORDS};
MC_TABLE unsigned char MC_EA[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
MC_TABLE unsigned char MC_EB[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
MC_TABLE unsigned char MC_CX[8] = {0, 1, 1, 0, 0, 1, 1, 0};
MC_TABLE unsigned char MC_CY[8] = {0, 0, 1, 1, 0, 0, 1, 1};
MC_TABLE unsigned char MC_CZ[8] = {0, 0, 0, 0, 1, 1, 1, 1};
// edge id from its direction and the offsets of its line inside the cube (the two other axes, in xyz order)
MC_TABLE unsigned char MC_EDGE_ID[3][2][2] = {{{0, 4}, {2, 6}},     // x: [y][z]
                                              {{3, 7}, {1, 5}},     // y: [x][z]
                                              {{8, 11}, {9, 10}}};  // z: [x][y]

struct EdgeGeom { int dir, l[3]; };   // direction and position (0/1 per axis, 0 along dir) of edge e inside its cube
DEVFN EdgeGeom edge_geom(int e) {
  const int a = MC_EA[e], b = MC_EB[e];
  EdgeGeom g;
  g.dir = MC_CX[a] != MC_CX[b] ? 0 : (MC_CY[a] != MC_CY[b] ? 1 : 2);
  g.l[0] = MC_CX[a] < MC_CX[b] ? MC_CX[a] : MC_CX[b];
  g.l[1] = MC_CY[a] < MC_CY[b] ? MC_CY[a] : MC_CY[b];
  g.l[2] = MC_CZ[a] < MC_CZ[b] ? MC_CZ[a] : MC_CZ[b];
  return g;
}
DEVFN int edge_id(int dir, const int l[3]) {
  return dir == 0 ? MC_EDGE_ID[0][l[1]][l[2]] : dir == 1 ? MC_EDGE_ID[1][l[0]][l[2]] : MC_EDGE_ID[2][l[0]][l[1]];
}

struct McDims {
  int nx, ny, nz;
  HD int64_t cubes() const { return (int64_t)(nx - 1) * (ny - 1) * (nz - 1); }
  DEVFN void cube_xyz(int64_t q, int c[3]) const {
    c[0] = (int)(q % (nx - 1)); c[1] = (int)((q / (nx - 1)) % (ny - 1)); c[2] = (int)(q / ((int64_t)(nx - 1) * (ny - 1)));
  }
  DEVFN int64_t cube_index(const int c[3]) const { return c[0] + (int64_t)(nx - 1) * (c[1] + (int64_t)(ny - 1) * c[2]); }
  DEVFN int64_t vertex(const int c[3], int corner) const {
    return (c[0] + MC_CX[corner]) + (int64_t)nx * ((c[1] + MC_CY[corner]) + (int64_t)ny * (c[2] + MC_CZ[corner]));
  }
};
// first cube in scan order (x fastest, then y, then z) that contains edge e of cube c, and the edge's id there
DEVFN void edge_owner(const int c[3], int e, int oc[3], int &oe) {
  const EdgeGeom g = edge_geom(e);
  int lo[3];
  for (int k = 0; k < 3; ++k) {
    if (k == g.dir) { oc[k] = c[k]; lo[k] = 0; continue; }
    const int p = c[k] + g.l[k];          // grid coordinate of the edge's line
    oc[k] = p > 0 ? p - 1 : 0;            // the cube before it, if there is one
    lo[k] = p - oc[k];
  }
  oe = edge_id(g.dir, lo);
}
DEVFN int popc12(uint32_t m) { int n = 0; for (int i = 0; i < 12; ++i) n += (m >> i) & 1u; return n; }

struct McClassify {   // march_cube.cpp:79-87 per cube, plus which cut edges this cube is the first to meet
  McDims d; const double *S; double shift, iso;
  uint32_t *info;     // bits 0-7 corner flags, bits 8-19 owned cut edges
  uint32_t *vcount, *fcount;
  DEVFN void operator()(int64_t q) const {
    int c[3]; d.cube_xyz(q, c);
    uint32_t cf = 0;
    for (int k = 0; k < 8; ++k) if (dsub(S[d.vertex(c, k)], shift) > iso) cf |= 1u << k;
    uint32_t own = 0;
    if (cf != 0 && cf != 255) {
      for (int e = 0; e < 12; ++e) {
        if ((((cf >> MC_EA[e]) ^ (cf >> MC_EB[e])) & 1u) == 0) continue;   // aiCubeEdgeFlags: cut iff the corners differ
        int oc[3], oe; edge_owner(c, e, oc, oe);
        if (oc[0] == c[0] && oc[1] == c[1] && oc[2] == c[2]) own |= 1u << e;
      }
    }
    const uint64_t w = MC_TRI[cf];
    int nf = 0;
    while (nf < 5 && ((w >> (12 * nf)) & 0xF) != 0xF) ++nf;
    info[q] = cf | (own << 8);
    vcount[q] = (uint32_t)popc12(own);
    fcount[q] = (uint32_t)nf;
  }
};
struct McVertices {   // march_cube.cpp:100-123 for the edges a cube owns
  McDims d; const double *S; const double *GV; int64_t n; double shift, iso;
  const uint32_t *info, *vbase;
  double *V; int32_t *E;
  DEVFN void operator()(int64_t q) const {
    uint32_t own = info[q] >> 8;
    if (!own) return;
    int c[3]; d.cube_xyz(q, c);
    uint32_t v = vbase[q];
    for (int e = 0; e < 12; ++e) {
      if (!((own >> e) & 1u)) continue;
      const int64_t i = d.vertex(c, MC_EA[e]), j = d.vertex(c, MC_EB[e]);
      const double a = dsub(S[i], shift), b = dsub(S[j], shift);
      const double t = ddiv(dsub(iso, a), dsub(b, a));
      for (int k = 0; k < 3; ++k) {
        const double gi = GV[k * n + i], gj = GV[k * n + j];
        V[3 * (size_t)v + k] = dadd(gi, dmul(t, dsub(gj, gi)));
      }
      E[2 * (size_t)v] = (int32_t)(i < j ? i : j);
      E[2 * (size_t)v + 1] = (int32_t)(i < j ? j : i);
      ++v;
    }
  }
};
struct McFaces {      // march_cube.cpp:126-141: ids of the three edges' vertices, looked up at the owners
  McDims d;
  const uint32_t *info, *vbase, *fbase;
  int32_t *F;
  DEVFN void operator()(int64_t q) const {
    const uint32_t cf = info[q] & 255u;
    const uint64_t w = MC_TRI[cf];
    if ((w & 0xF) == 0xF) return;
    int c[3]; d.cube_xyz(q, c);
    int32_t ev[12];
    for (int e = 0; e < 12; ++e) {
      ev[e] = -1;
      if ((((cf >> MC_EA[e]) ^ (cf >> MC_EB[e])) & 1u) == 0) continue;
      int oc[3], oe; edge_owner(c, e, oc, oe);
      const int64_t oq = d.cube_index(oc);
      const uint32_t own = info[oq] >> 8;
      ev[e] = (int32_t)(vbase[oq] + (uint32_t)popc12(own & ((1u << oe) - 1u)));
    }
    size_t f = fbase[q];
    for (int t = 0; t < 5; ++t) {
      if (((w >> (12 * t)) & 0xF) == 0xF) break;
      for (int k = 0; k < 3; ++k) F[3 * f + k] = ev[(w >> (12 * t + 4 * k)) & 0xF];
      ++f;
    }
  }
};

// 64-bit totals of the per-cube counts: the 32-bit scans below must not wrap silently
struct Sum64 {
  const uint32_t *a, *b; int64_t n; unsigned long long *out;   // out[0] += sum a, out[1] += sum b; 1024 entries per call
  DEVFN void operator()(int64_t k) const {
    unsigned long long sa = 0, sb = 0;
    const int64_t hi = (k + 1) * 1024 < n ? (k + 1) * 1024 : n;
    for (int64_t i = k * 1024; i < hi; ++i) { sa += a[i]; sb += b[i]; }
#ifdef IGLB200_EMUL
    out[0] += sa; out[1] += sb;
#else
    atomicAdd(&out[0], sa); atomicAdd(&out[1], sb);
#endif
  }
};
inline void check_totals(const uint32_t *vcount, const uint32_t *fcount, int64_t nc, stream_t st) {
  plat::DevBuf<unsigned long long> tot(2);
  plat::dev_memset(tot.p, 0, 16, st);
  plat::for_n((nc + 1023) / 1024, Sum64{vcount, fcount, nc, tot.p}, st);
  unsigned long long h[2];
  plat::d2h(h, tot.p, 16, st);
  plat::sync(st);
  if (h[0] > 0x7FFFFFFFull || h[1] > 0x7FFFFFFFull) throw std::invalid_argument("marching_cubes: more than 2^31-1 vertices or faces");
}

// S: n values (x fastest); GV: column-major 3 x n; the iso-surface of (S - shift) at `iso`
inline iglb200_surface *marching_cubes(const double *S, const double *GV, int nx, int ny, int nz, double shift, double iso, stream_t st) {
  if (nx < 1 || ny < 1 || nz < 1) throw std::invalid_argument("marching_cubes: empty grid");
  const int64_t n = (int64_t)nx * ny * nz;
  if (n > 0x7FFFFFFF) throw std::invalid_argument("marching_cubes: more than 2^31-1 grid vertices");
  iglb200_surface *out = new iglb200_surface;
  try {
    const McDims d{nx, ny, nz};
    const int64_t nc = (nx > 1 && ny > 1 && nz > 1) ? d.cubes() : 0;
    if (nc > 0) {
      plat::Span sp("marching_cubes", st);
      plat::DevBuf<uint32_t> info((size_t)nc), vcount((size_t)nc + 1), fcount((size_t)nc + 1), vbase((size_t)nc + 1), fbase((size_t)nc + 1);
      plat::for_n(nc, McClassify{d, S, shift, iso, info.p, vcount.p, fcount.p}, st);
      plat::dev_memset(vcount.p + nc, 0, 4, st);
      plat::dev_memset(fcount.p + nc, 0, 4, st);
      check_totals(vcount.p, fcount.p, nc, st);
      plat::exclusive_scan_u32(vcount.p, vbase.p, nc + 1, st);   // entry nc = total
      plat::exclusive_scan_u32(fcount.p, fbase.p, nc + 1, st);
      uint32_t tot[2];
      plat::d2h(&tot[0], vbase.p + nc, 4, st);
      plat::d2h(&tot[1], fbase.p + nc, 4, st);
      plat::sync(st);
      out->nV = tot[0]; out->nF = tot[1];
      if ((uint64_t)out->nV > 0x7FFFFFFFull) throw std::invalid_argument("marching_cubes: more than 2^31-1 vertices");
      vcount.release(); fcount.release();
      out->V.alloc((size_t)3 * out->nV); out->E.alloc((size_t)2 * out->nV); out->F.alloc((size_t)3 * out->nF);
      if (out->nV) plat::for_n(nc, McVertices{d, S, GV, n, shift, iso, info.p, vbase.p, out->V.p, out->E.p}, st);
      if (out->nF) plat::for_n(nc, McFaces{d, info.p, vbase.p, fbase.p, out->F.p}, st);
      plat::sync(st);
    }
  } catch (...) { delete out; throw; }
  return out;
}

struct Iota32 { uint32_t *v; DEVFN void operator()(int64_t i) const { v[i] = (uint32_t)i; } };
// ------------------------------------------------------------------ marching_cubes over an explicit cube list
// igl::marching_cubes(S, GV, GI, isovalue, V, F, E2V) (marching_cubes.cpp:114-160): cube c has the corners
// GI(c, 0..7) (any vertex indices, e.g. of a sparse voxel grid).  There is no grid structure to derive the first
// cube of an edge from, so the hash map is replaced by a sort: every cube emits its cut edges in order (cube, edge),
// a stable radix sort by the edge key (i < j packed) groups the occurrences of an edge with its first visitor in
// front, and the groups ranked by that first visitor's emission index are the reference's vertex ids.
struct SpClassify {
  const double *S; const int32_t *GI; int64_t gi_rs, gi_cs; double iso;
  unsigned char *cf8; uint32_t *ecount, *fcount;
  DEVFN void operator()(int64_t c) const {
    uint32_t cf = 0;
    for (int k = 0; k < 8; ++k) if (S[GI[c * gi_rs + k * gi_cs]] > iso) cf |= 1u << k;
    int ne = 0;
    for (int e = 0; e < 12; ++e) ne += (int)(((cf >> MC_EA[e]) ^ (cf >> MC_EB[e])) & 1u);
    const uint64_t w = MC_TRI[cf];
    int nf = 0;
    while (nf < 5 && ((w >> (12 * nf)) & 0xF) != 0xF) ++nf;
    cf8[c] = (unsigned char)cf; ecount[c] = (uint32_t)ne; fcount[c] = (uint32_t)nf;
  }
};
struct SpEmitEdges {   // emission index = ebase[c] + rank of the edge among the cube's cut edges
  const int32_t *GI; int64_t gi_rs, gi_cs;
  const unsigned char *cf8; const uint32_t *ebase;
  uint64_t *key; uint32_t *val, *cube_edge;
  DEVFN void operator()(int64_t c) const {
    const uint32_t cf = cf8[c];
    uint32_t idx = ebase[c];
    for (int e = 0; e < 12; ++e) {
      if ((((cf >> MC_EA[e]) ^ (cf >> MC_EB[e])) & 1u) == 0) continue;
      const uint32_t i = (uint32_t)GI[c * gi_rs + MC_EA[e] * gi_cs], j = (uint32_t)GI[c * gi_rs + MC_EB[e] * gi_cs];
      key[idx] = i < j ? ((uint64_t)i | ((uint64_t)j << 32)) : ((uint64_t)j | ((uint64_t)i << 32));   // march_cube.cpp:52-60
      val[idx] = idx;
      cube_edge[idx] = (uint32_t)(c * 12 + e);
      ++idx;
    }
  }
};
struct SpHeads { const uint64_t *skey; uint32_t *head; DEVFN void operator()(int64_t p) const { head[p] = (p == 0 || skey[p - 1] != skey[p]) ? 1u : 0u; } };
struct SpOwners {   // group g = number of heads before p; its owner is the emission index at the head
  const uint32_t *head, *gscan, *sval; uint32_t *owner, *gid_of_entry;
  DEVFN void operator()(int64_t p) const {
    const uint32_t g = gscan[p] + head[p] - 1u;
    if (head[p]) owner[g] = sval[p];
    gid_of_entry[sval[p]] = g;
  }
};
struct SpRank { const uint32_t *gsorted; uint32_t *vid; DEVFN void operator()(int64_t r) const { vid[gsorted[r]] = (uint32_t)r; } };
struct SpVertices {
  const double *S; const double *GV; int64_t n; const int32_t *GI; int64_t gi_rs, gi_cs; double iso;
  const uint32_t *owner, *vid, *cube_edge;
  double *V; int32_t *E;
  DEVFN void operator()(int64_t g) const {
    const uint32_t ce = cube_edge[owner[g]], v = vid[g];
    const int64_t c = ce / 12; const int e = (int)(ce % 12);
    const int64_t i = GI[c * gi_rs + MC_EA[e] * gi_cs], j = GI[c * gi_rs + MC_EB[e] * gi_cs];
    const double a = S[i], b = S[j];
    const double t = ddiv(dsub(iso, a), dsub(b, a));
    for (int k = 0; k < 3; ++k) {
      const double gi = GV[k * n + i], gj = GV[k * n + j];
      V[3 * (size_t)v + k] = dadd(gi, dmul(t, dsub(gj, gi)));
    }
    E[2 * (size_t)v] = (int32_t)(i < j ? i : j);
    E[2 * (size_t)v + 1] = (int32_t)(i < j ? j : i);
  }
};
struct SpFaces {
  const unsigned char *cf8; const uint32_t *ebase, *fbase, *gid_of_entry, *vid;
  int32_t *F;
  DEVFN void operator()(int64_t c) const {
    const uint32_t cf = cf8[c];
    const uint64_t w = MC_TRI[cf];
    if ((w & 0xF) == 0xF) return;
    int32_t ev[12];
    uint32_t idx = ebase[c];
    for (int e = 0; e < 12; ++e) {
      ev[e] = -1;
      if ((((cf >> MC_EA[e]) ^ (cf >> MC_EB[e])) & 1u) == 0) continue;
      ev[e] = (int32_t)vid[gid_of_entry[idx++]];
    }
    size_t f = fbase[c];
    for (int t = 0; t < 5; ++t) {
      if (((w >> (12 * t)) & 0xF) == 0xF) break;
      for (int k = 0; k < 3; ++k) F[3 * f + k] = ev[(w >> (12 * t + 4 * k)) & 0xF];
      ++f;
    }
  }
};
struct SpCheckIndices { const int32_t *GI; int64_t gi_rs, gi_cs, n; uint32_t *bad; DEVFN void operator()(int64_t q) const {
  const int32_t v = GI[(q / 8) * gi_rs + (q % 8) * gi_cs]; if (v < 0 || v >= n) plat::atomic_max(bad, 1u); } };

// S: n values; GV: column-major 3 x n; GI: m x 8 device int32 with the given strides
inline iglb200_surface *marching_cubes_sparse(const double *S, const double *GV, int64_t n, const int32_t *GI, int64_t gi_rs, int64_t gi_cs,
                                              int64_t m, double iso, stream_t st) {
  using plat::DevBuf;
  if (m < 0 || n < 0 || n > 0x7FFFFFFF || m > (int64_t)(0xFFFFFFFFu / 12)) throw std::invalid_argument("marching_cubes: too many cubes or vertices");
  iglb200_surface *out = new iglb200_surface;
  try {
    if (m == 0) return out;
    plat::Span sp("marching_cubes", st);
    DevBuf<uint32_t> bad(1);
    plat::dev_memset(bad.p, 0, 4, st);
    plat::for_n(8 * m, SpCheckIndices{GI, gi_rs, gi_cs, n, bad.p}, st);
    uint32_t hbad = 0;
    plat::d2h(&hbad, bad.p, 4, st);
    plat::sync(st);
    if (hbad) throw std::invalid_argument("marching_cubes: GI contains a vertex index outside [0, #GV)");
    DevBuf<unsigned char> cf8((size_t)m);
    DevBuf<uint32_t> ecount((size_t)m + 1), fcount((size_t)m + 1), ebase((size_t)m + 1), fbase((size_t)m + 1);
    plat::for_n(m, SpClassify{S, GI, gi_rs, gi_cs, iso, cf8.p, ecount.p, fcount.p}, st);
    plat::dev_memset(ecount.p + m, 0, 4, st);
    plat::dev_memset(fcount.p + m, 0, 4, st);
    plat::exclusive_scan_u32(ecount.p, ebase.p, m + 1, st);
    plat::exclusive_scan_u32(fcount.p, fbase.p, m + 1, st);
    uint32_t tot[2];
    plat::d2h(&tot[0], ebase.p + m, 4, st);
    plat::d2h(&tot[1], fbase.p + m, 4, st);
    plat::sync(st);
    const int64_t ne = tot[0];
    out->nF = tot[1];
    ecount.release(); fcount.release();
    if (ne == 0) return out;
    DevBuf<uint64_t> key((size_t)ne), skey((size_t)ne);
    DevBuf<uint32_t> val((size_t)ne), sval((size_t)ne), cube_edge((size_t)ne);
    plat::for_n(m, SpEmitEdges{GI, gi_rs, gi_cs, cf8.p, ebase.p, key.p, val.p, cube_edge.p}, st);
    {
      const size_t tb = plat::sort_pairs64_temp_bytes(ne);
      DevBuf<char> tmp(tb);
      plat::sort_pairs_u64(tmp.p, tb, key.p, skey.p, val.p, sval.p, ne, st);   // stable: first visitor first in its group
    }
    key.release(); val.release();
    DevBuf<uint32_t> head((size_t)ne + 1), gscan((size_t)ne + 1), gid((size_t)ne);
    plat::for_n(ne, SpHeads{skey.p, head.p}, st);
    plat::dev_memset(head.p + ne, 0, 4, st);
    plat::exclusive_scan_u32(head.p, gscan.p, ne + 1, st);
    uint32_t ng = 0;
    plat::d2h(&ng, gscan.p + ne, 4, st);
    plat::sync(st);
    out->nV = ng;
    DevBuf<uint32_t> owner((size_t)ng), gidx((size_t)ng), owner_sorted((size_t)ng), gsorted((size_t)ng), vid((size_t)ng);
    plat::for_n(ne, SpOwners{head.p, gscan.p, sval.p, owner.p, gid.p}, st);
    plat::for_n(ng, Iota32{gidx.p}, st);
    {
      const size_t tb = plat::sort_pairs_temp_bytes(ng);
      DevBuf<char> tmp(tb);
      plat::sort_pairs_u32(tmp.p, tb, owner.p, owner_sorted.p, gidx.p, gsorted.p, ng, 32, st);   // vertex ids in order of first visit
    }
    plat::for_n(ng, SpRank{gsorted.p, vid.p}, st);
    out->V.alloc((size_t)3 * ng); out->E.alloc((size_t)2 * ng); out->F.alloc((size_t)3 * out->nF + 1);
    plat::for_n(ng, SpVertices{S, GV, n, GI, gi_rs, gi_cs, iso, owner.p, vid.p, cube_edge.p, out->V.p, out->E.p}, st);
    if (out->nF) plat::for_n(m, SpFaces{cf8.p, ebase.p, fbase.p, gid.p, vid.p, out->F.p}, st);
    plat::sync(st);
  } catch (...) { delete out; throw; }
  return out;
}

}  // namespace grid
