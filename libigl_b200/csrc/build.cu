// Device build of the mesh handle: gather faces, 4-ary cluster tree
// (fwn_build.h), bottom-up moments and closest-point boxes (moments.h).
// Compiled with -fmad=false: every fp32 value that feeds a branch in the
// reference must be reproduced bit for bit.
#define IGLB200_BUILDER_TU 1
#include "mesh.h"
#include "wn.h"
#include "pill.h"

namespace fwn {

// V (any stride, f32 or f64) + F -> per-face fp64 and fp32 vertices.
// fp32 = (float)fp64, the reference's V.cast<float>() (signed_distance.cpp:66).
template <class VT> struct GatherFaces {
  const VT *V;
  int64_t nV, v_rs, v_cs;
  const int32_t *F;
  int64_t f_rs, f_cs;
  double *tri64;
  float4 *tri32;
  uint32_t *err;      // [0] = out-of-range index seen
  uint32_t *absacc;   // order-encoded max |coordinate|
  uint32_t *bbacc;    // [6] order-encoded fp32 hull (rounded outward) of the geometry
  DEVFN void operator()(int64_t f) const {
    float amax = 0.f;
    for (int k = 0; k < 3; ++k) {
      int64_t vi = F[f * f_rs + k * f_cs];
      if (vi < 0 || vi >= nV) { plat::atomic_max(err, 1u); vi = 0; }
      double x[3];
      for (int d = 0; d < 3; ++d) x[d] = (double)V[vi * v_rs + d * v_cs];
      for (int d = 0; d < 3; ++d) tri64[9 * f + 3 * k + d] = x[d];
      tri32[TRI32_STRIDE * f + k] = make_float4((float)x[0], (float)x[1], (float)x[2], 0.f);
      for (int d = 0; d < 3; ++d) {
        const float up = f_round_up(fabs(x[d]));
        if (up == up) amax = fmaxf(amax, up);
        const float lo = f_round_down(x[d]), hi = f_round_up(x[d]);
        if (lo == lo && hi == hi) { plat::atomic_min(&bbacc[d], plat::f2ord(lo)); plat::atomic_max(&bbacc[3 + d], plat::f2ord(hi)); }
      }
    }
    plat::atomic_max(absacc, plat::f2ord(amax));
  }
};

// fp32-stage record of the closest-triangle search (traverse.h: tri_dist_f32): the face in its own orthonormal frame,
// computed in fp64 from the vertices SCALED by sigma (a power of two, exact)
struct FaceRecords {
  const double *tri64; float4 *triq; double sigma;
  DEVFN void operator()(int64_t f) const {
    double vtx[3][3];
    for (int k = 0; k < 3; ++k) for (int d = 0; d < 3; ++d) vtx[k][d] = tri64[9 * f + 3 * k + d] * sigma;
    int k0 = 0;
    double best = -1.0;
    for (int k = 0; k < 3; ++k) {
      double l2 = 0.0;
      for (int d = 0; d < 3; ++d) { const double e = vtx[(k + 1) % 3][d] - vtx[k][d]; l2 += e * e; }
      if (l2 > best) { best = l2; k0 = k; }
    }
    const double *A = vtx[k0], *B = vtx[(k0 + 1) % 3], *Cc = vtx[(k0 + 2) % 3];
    double u[3], ac[3], n[3], v[3];
    const double Ln = sqrt(best > 0.0 ? best : 0.0);
    for (int d = 0; d < 3; ++d) { u[d] = (Ln > 0.0) ? (B[d] - A[d]) / Ln : (d == 0 ? 1.0 : 0.0); ac[d] = Cc[d] - A[d]; }
    n[0] = u[1] * ac[2] - u[2] * ac[1]; n[1] = u[2] * ac[0] - u[0] * ac[2]; n[2] = u[0] * ac[1] - u[1] * ac[0];
    double nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double acl = sqrt(ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2]);
    bool flat = !(nl > 1e-14 * acl) || !(nl == nl) || !(Ln > 0.0);
    if (!flat) {
      // the cross product of a thin face is mostly rounding noise: force n orthogonal to u.
      // A wrong rotation of n about u by theta moves the apex by height*theta ~ the noise itself.
      for (int d = 0; d < 3; ++d) n[d] /= nl;
      const double nu = n[0] * u[0] + n[1] * u[1] + n[2] * u[2];
      for (int d = 0; d < 3; ++d) n[d] -= nu * u[d];
      nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      flat = !(nl > 0.5);
    }
    if (flat) {
      // segment or point: any unit normal orthogonal to u will do
      const int m = (fabs(u[0]) <= fabs(u[1]) && fabs(u[0]) <= fabs(u[2])) ? 0 : (fabs(u[1]) <= fabs(u[2]) ? 1 : 2);
      double e[3] = {0.0, 0.0, 0.0};
      e[m] = 1.0;
      n[0] = u[1] * e[2] - u[2] * e[1]; n[1] = u[2] * e[0] - u[0] * e[2]; n[2] = u[0] * e[1] - u[1] * e[0];
      nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    }
    for (int d = 0; d < 3; ++d) n[d] /= nl;
    v[0] = n[1] * u[2] - n[2] * u[1]; v[1] = n[2] * u[0] - n[0] * u[2]; v[2] = n[0] * u[1] - n[1] * u[0];
    double cx = u[0] * ac[0] + u[1] * ac[1] + u[2] * ac[2];
    double cy = flat ? 0.0 : (v[0] * ac[0] + v[1] * ac[1] + v[2] * ac[2]);
    if (flat) {
      // the apex lies on the line of the longest edge (within rounding): clamp it onto the edge
      cx = cx < 0.0 ? 0.0 : (cx > Ln ? Ln : cx);
    } else if (cy < 0.0) { cy = 0.0; }
    float4 *q = triq + TRIQ_STRIDE * f;
    q[0] = make_float4((float)A[0], (float)A[1], (float)A[2], (float)Ln);
    q[1] = make_float4((float)u[0], (float)u[1], (float)u[2], (float)cx);
    q[2] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)cy);
    q[3] = make_float4((float)n[0], (float)n[1], (float)n[2], 0.f);
  }
};

// ---- pseudonormals (SIGNED_DISTANCE_TYPE_PSEUDONORMAL): per_face_normals.cpp:14-37,
// internal_angles.cpp:13-43, per_vertex_normals.cpp:40-104 (ANGLE), per_edge_normals.cpp:24-76 (UNIFORM).
// Sums run in the reference's order (faces by index) so the normals come out bit-identical:
// corners / half-edges are stably radix-sorted by vertex / undirected-edge key and each group is
// summed sequentially by the thread that finds its head.
static constexpr int PN_STRIDE = 21;

struct FaceNormalAngles {
  const double *tri64; double *pn; double *ang;  // ang: 3 per face
  DEVFN void operator()(int64_t f) const {
    const double *t = tri64 + 9 * f;
    const double v1[3] = {t[3] - t[0], t[4] - t[1], t[5] - t[2]}, v2[3] = {t[6] - t[0], t[7] - t[1], t[8] - t[2]};
    const double n[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
    const double r = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int d = 0; d < 3; ++d) pn[PN_STRIDE * f + d] = (r == 0) ? 0.0 : n[d] / r;
    for (int j = 0; j < 3; ++j) {
      const double *x = t + 3 * ((j + 2) % 3), *y = t + 3 * j, *z = t + 3 * ((j + 1) % 3);
      double a1[3], a2[3];
      for (int d = 0; d < 3; ++d) { a1[d] = x[d] - y[d]; a2[d] = z[d] - y[d]; }
      const double n1 = sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), n2 = sqrt(a2[0] * a2[0] + a2[1] * a2[1] + a2[2] * a2[2]);
      double u[3], w[3];
      for (int d = 0; d < 3; ++d) { u[d] = a1[d] / n1 - a2[d] / n2; w[d] = a1[d] / n1 + a2[d] / n2; }
      ang[3 * f + j] = 2.0 * atan(sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) / sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]));
    }
  }
};
struct CornerKeys {  // corner c = 3f+j: vertex key (32 bit) and undirected-edge key (64 bit, the edge opposite corner j)
  const int32_t *F; int64_t f_rs, f_cs;
  uint32_t *vkey, *val; uint64_t *ekey;
  DEVFN void operator()(int64_t c) const {
    const int64_t f = c / 3; const int j = (int)(c % 3);
    vkey[c] = (uint32_t)F[f * f_rs + j * f_cs];
    const uint64_t a = (uint32_t)F[f * f_rs + ((j + 1) % 3) * f_cs], b = (uint32_t)F[f * f_rs + ((j + 2) % 3) * f_cs];
    ekey[c] = ((a < b ? a : b) << 32) | (a < b ? b : a);
    val[c] = (uint32_t)c;
  }
};
DEVFN void pn_normalize(double *v) {
  const double z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (z > 0) { const double r = sqrt(z); v[0] /= r; v[1] /= r; v[2] /= r; }
}
struct VertexNormals {  // one thread per sorted corner; group heads do the (ordered) sum
  const uint32_t *key, *val; int64_t n; const double *ang; double *pn;
  DEVFN void operator()(int64_t i) const {
    if (i > 0 && key[i - 1] == key[i]) return;
    double acc[3] = {0, 0, 0};
    int64_t j = i;
    for (; j < n && key[j] == key[i]; ++j) {
      const uint32_t c = val[j];
      const double w = ang[c];
      for (int d = 0; d < 3; ++d) acc[d] += w * pn[PN_STRIDE * (int64_t)(c / 3) + d];
    }
    pn_normalize(acc);
    for (int64_t k = i; k < j; ++k) { const uint32_t c = val[k]; for (int d = 0; d < 3; ++d) pn[PN_STRIDE * (int64_t)(c / 3) + 3 + 3 * (c % 3) + d] = acc[d]; }
  }
};
struct EdgeNormals {
  const uint64_t *key; const uint32_t *val; int64_t n; double *pn;
  DEVFN void operator()(int64_t i) const {
    if (i > 0 && key[i - 1] == key[i]) return;
    double acc[3] = {0, 0, 0};
    int64_t j = i;
    for (; j < n && key[j] == key[i]; ++j) { const uint32_t c = val[j]; for (int d = 0; d < 3; ++d) acc[d] += pn[PN_STRIDE * (int64_t)(c / 3) + d]; }
    pn_normalize(acc);
    for (int64_t k = i; k < j; ++k) { const uint32_t c = val[k]; for (int d = 0; d < 3; ++d) pn[PN_STRIDE * (int64_t)(c / 3) + 12 + 3 * (c % 3) + d] = acc[d]; }
  }
};

static void build_pseudonormals(const double *tri64, const int32_t *dF, int64_t f_rs, int64_t f_cs, int64_t nF, double *pn, stream_t s) {
  using plat::DevBuf;
  if (nF == 0) return;
  const int64_t nc = 3 * nF;
  DevBuf<double> ang((size_t)nc);
  DevBuf<uint32_t> vkey((size_t)nc), vkey2((size_t)nc), val((size_t)nc), val2((size_t)nc);
  DevBuf<uint64_t> ekey((size_t)nc), ekey2((size_t)nc);
  plat::for_n(nF, FaceNormalAngles{tri64, pn, ang.p}, s);
  plat::for_n(nc, CornerKeys{dF, f_rs, f_cs, vkey.p, val.p, ekey.p}, s);
  {
    const size_t tb = plat::sort_pairs_temp_bytes(nc);
    DevBuf<char> tmp(tb);
    plat::sort_pairs_u32(tmp.p, tb, vkey.p, vkey2.p, val.p, val2.p, nc, 32, s);
    plat::for_n(nc, VertexNormals{vkey2.p, val2.p, nc, ang.p, pn}, s);
    plat::sync(s);
  }
  {
    const size_t tb = plat::sort_pairs64_temp_bytes(nc);
    DevBuf<char> tmp(tb);
    plat::sort_pairs_u64(tmp.p, tb, ekey.p, ekey2.p, val.p, val2.p, nc, s);
    plat::for_n(nc, EdgeNormals{ekey2.p, val2.p, nc, pn}, s);
    plat::sync(s);
  }
}

static int tree_levels(const Tree &t) { return (int)t.level_start.size() - 1; }

iglb200_mesh *create_mesh(const void *V, int v_is_f64, int64_t nV, int64_t v_rs, int64_t v_cs, const int32_t *F,
                          int64_t nF, int64_t f_rs, int64_t f_cs, int order, int flags, stream_t s) {
  using plat::DevBuf;
  iglb200_mesh *m = new iglb200_mesh;
  std::memset(&m->h, 0, sizeof(m->h));
  m->h.magic = IGLB200_ARENA_MAGIC;
  m->h.nF = nF; m->h.nV = nV; m->h.order = order;
  try {
    // raw uploads.  Only the bytes actually spanned by the strided views are copied.
    const size_t vsz = v_is_f64 ? 8 : 4;
    const size_t v_span = (nV > 0) ? (size_t)((nV - 1) * v_rs + 2 * v_cs + 1) : 0;
    const size_t f_span = (nF > 0) ? (size_t)((nF - 1) * f_rs + 2 * f_cs + 1) : 0;
    DevBuf<char> dV(v_span * vsz);
    DevBuf<int32_t> dF(f_span);
    plat::h2d(dV.p, V, v_span * vsz, s);
    plat::h2d(dF.p, F, f_span * 4, s);
    DevBuf<double> tri64((size_t)9 * nF);
    DevBuf<float4> tri32((size_t)TRI32_STRIDE * nF), triq((size_t)TRIQ_STRIDE * nF);
    DevBuf<uint32_t> acc(8);
    {
      const uint32_t init[8] = {0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
      plat::h2d(acc.p, init, sizeof(init), s);
    }
    if (v_is_f64)
      plat::for_n(nF, GatherFaces<double>{(const double *)dV.p, nV, v_rs, v_cs, dF.p, f_rs, f_cs, tri64.p, tri32.p, acc.p, acc.p + 1, acc.p + 2}, s);
    else
      plat::for_n(nF, GatherFaces<float>{(const float *)dV.p, nV, v_rs, v_cs, dF.p, f_rs, f_cs, tri64.p, tri32.p, acc.p, acc.p + 1, acc.p + 2}, s);
    uint32_t hacc[8];
    plat::d2h(hacc, acc.p, sizeof(hacc), s);
    plat::sync(s);
    if (hacc[0]) throw std::invalid_argument("F contains a vertex index outside [0, #V)");
    m->h.absmax = nF ? plat::ord2f(hacc[1]) : 0.f;
    {
      // sigma = 2^-e with absmax = f * 2^e, f in [0.5, 1)
      int e = 0;
      const double am = (double)m->h.absmax;
      m->h.sigma = 1.0;
      if (am > 0.0 && am < INFINITY) { std::frexp(am, &e); m->h.sigma = std::ldexp(1.0, -e); }
    }
    plat::for_n(nF, FaceRecords{tri64.p, triq.p, m->h.sigma}, s);
    for (int d = 0; d < 3; ++d) {
      m->h.bbox_lo[d] = nF ? (double)plat::ord2f(hacc[2 + d]) : 0.0;
      m->h.bbox_hi[d] = nF ? (double)plat::ord2f(hacc[5 + d]) : 0.0;
    }
    dV.release();

    DevBuf<TriBox> boxes;
    Tree tree;
    build_tree(tri32.p, (uint32_t)nF, boxes, tree, s);
    const int64_t nn = tree.nnodes;
    m->h.nnodes = nn;
    m->h.nexcluded = tree.nexcluded;
    m->h.nlevels = nn ? tree_levels(tree) : 0;

    auto align = [](uint64_t x) { return (x + 255) & ~uint64_t(255); };
    uint64_t off = align(sizeof(ArenaHeader));
    m->h.off_tri64 = off; off = align(off + (uint64_t)72 * nF);
    m->h.off_tri32 = off; off = align(off + (uint64_t)16 * TRI32_STRIDE * nF);
    m->h.off_triq = off; off = align(off + (uint64_t)16 * TRIQ_STRIDE * nF);
    m->h.off_fnodes = off; off = align(off + (uint64_t)sizeof(FwnNode) * nn);
    // the box nodes serve the exact winding number only (wn.h descends into clusters whose box holds the query)
    DevBuf<CpNode> cn_tmp;
    const bool keep_boxes = (flags & 2) != 0;
    m->h.off_cnodes = 0;
    if (keep_boxes) { m->h.off_cnodes = off; off = align(off + (uint64_t)sizeof(CpNode) * nn); }
    else cn_tmp.alloc((size_t)nn);
    m->h.off_pnodes = off; off = align(off + (uint64_t)sizeof(PillNode) * nn);
    m->h.off_pn = 0;
    if (flags & 1) { m->h.off_pn = off; off = align(off + (uint64_t)8 * PN_STRIDE * nF); }
    m->h.off_wn_cnt = m->h.off_wn_start = m->h.off_wn_bstart = m->h.off_wn_bent = m->h.off_wn_leafface = m->h.n_wn_bent = 0;
    wn::Built wb;
    wn::Ranges ranges;
    if (nn) wn::cluster_ranges(tree, nF, ranges, s);
    if ((flags & 2) && nn) {
      wn::build(tree, ranges, dF.p, f_rs, f_cs, nF, wb, s);
      m->h.n_wn_bent = wb.nent;
      m->h.off_wn_cnt = off; off = align(off + (uint64_t)16 * nn);
      m->h.off_wn_start = off; off = align(off + (uint64_t)16 * nn);
      m->h.off_wn_bstart = off; off = align(off + (uint64_t)4 * (4 * nn + 1));
      m->h.off_wn_bent = off; off = align(off + (uint64_t)8 * (wb.nent + 1));
      m->h.off_wn_leafface = off; off = align(off + (uint64_t)4 * nF);
    }
    m->h.bytes = off;
    m->arena = (char *)plat::dev_alloc(off);
    plat::d2d(m->arena + m->h.off_tri64, tri64.p, (size_t)72 * nF, s);
    plat::d2d(m->arena + m->h.off_tri32, tri32.p, (size_t)16 * TRI32_STRIDE * nF, s);
    plat::d2d(m->arena + m->h.off_triq, triq.p, (size_t)16 * TRIQ_STRIDE * nF, s);
    if (m->h.off_wn_cnt) {
      plat::d2d(m->arena + m->h.off_wn_cnt, wb.cnt.p, (size_t)16 * nn, s);
      plat::d2d(m->arena + m->h.off_wn_start, wb.start.p, (size_t)16 * nn, s);
      plat::d2d(m->arena + m->h.off_wn_bstart, wb.bstart.p, (size_t)4 * (4 * nn + 1), s);
      plat::d2d(m->arena + m->h.off_wn_bent, wb.bent.p, (size_t)8 * wb.nent, s);
      plat::d2d(m->arena + m->h.off_wn_leafface, wb.leafface.p, (size_t)4 * (nF - tree.nexcluded), s);
    }
    if (nn) {
      DevBuf<Local> locals((size_t)nn);
      for (int l = tree_levels(tree) - 1; l >= 0; --l) {
        const uint32_t first = tree.level_start[l], cnt = tree.level_start[l + 1] - first;
        plat::for_n(cnt, NodeMoments{tree.nodes.p, (const float4 *)(m->arena + m->h.off_tri32), (const double *)(m->arena + m->h.off_tri64),
                                     boxes.p, locals.p, (FwnNode *)(m->arena + m->h.off_fnodes),
                                     keep_boxes ? (CpNode *)(m->arena + m->h.off_cnodes) : cn_tmp.p, first, order}, s);
      }
      // oriented bounds of the closest-triangle search
      pill::build(tree, (const double *)(m->arena + m->h.off_tri64), ranges.cnt.p, ranges.start.p, ranges.leafface.p, m->h.sigma,
                  (PillNode *)(m->arena + m->h.off_pnodes), s);
      // faces dropped by the NaN/Inf filter = faces not referenced by any node
      plat::sync(s);
    }
    if (flags & 1) build_pseudonormals((const double *)(m->arena + m->h.off_tri64), dF.p, f_rs, f_cs, nF, (double *)(m->arena + m->h.off_pn), s);
    dF.release();
    plat::h2d(m->arena, &m->h, sizeof(ArenaHeader), s);
    plat::sync(s);
  } catch (...) {
    if (m->arena) plat::dev_free(m->arena);
    delete m;
    throw;
  }
  return m;
}

}  // namespace fwn
