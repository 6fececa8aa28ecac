// Point-cloud fast winding number on the device (config c4):
//   igl::octree                      /root/reference/include/igl/octree.cpp:16-168
//   fast_winding_number precompute   /root/reference/include/igl/fast_winding_number.cpp:28-124
//   fast_winding_number evaluation   /root/reference/include/igl/fast_winding_number.cpp:152-311
// double precision throughout, as every instantiation of the reference is.
//
// Not a port.  The reference builds the octree by recursion, keeps a list of ALL point
// indices in every cell and recomputes every cell's sums from those lists (O(n * depth) with
// a vector per cell).  Here:
//   * a point's path through the octree depends only on the point and the root cube, so each
//     point computes its own 42-level path code (the same `coordinate >= centre` comparisons
//     against the same recursively halved centres, so every point lands in the cell the
//     reference puts it in), the points are radix-sorted by code, and a cell is a contiguous
//     range of the sorted array;
//   * cells are created level by level (one thread per (cell, octant): binary search);
//   * area/normal sums and the expansion coefficients EC go bottom-up from the children
//     (parallel-axis shift); the radius R = max |P - CM| is an atomic max per ancestor;
//   * evaluation is one thread per query, depth-first with one stack entry per level.
// Cells are numbered breadth-first (the reference: depth-first); numbering is not observable.
// Depth is capped at 42 levels (the reference: 30 000): points closer than width*2^-42 share a leaf.
#pragma once
#include "platform.h"

struct iglb200_cloud {
  int64_t n = 0;
  int order = 2, nterms = 39;
  int64_t ncells = 0, nlevels = 0;
  uint64_t bytes = 0;
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};   // bounding box of the points
  // sorted points: position, area*normal
  plat::DevBuf<double> P, AN;
  // cells
  plat::DevBuf<uint32_t> cbegin, ccount;
  plat::DevBuf<int32_t> cchild;   // first child (8 consecutive cells) or -1
  plat::DevBuf<double> CM, R, EC; // [ncells*3], [ncells], [ncells*nterms]
};

namespace cloud {

static constexpr int MAX_LEVELS = 42;   // 2 x 21 octal digits in two 64-bit words
static constexpr double PI_4 = 4.0 * 3.1415926535897932384626433832795;  // igl/PI.h

template <class T> struct LoadPoints {  // raw (row-major n x 3, n x 3, n) -> doubles; bbox
  const T *P, *N, *A;
  double *p, *an, *a;
  unsigned long long *bb;  // [6] ordered-encoded min xyz, max xyz
  DEVFN void operator()(int64_t i) const {
    const double ar = (double)A[i];
    for (int d = 0; d < 3; ++d) {
      const double x = (double)P[3 * i + d];
      p[3 * i + d] = x;
      an[3 * i + d] = (double)N[3 * i + d] * ar;   // N.row(j)*A(j), fast_winding_number.cpp:247
      plat::atomic_min64(&bb[d], plat::d2ord(x));
      plat::atomic_max64(&bb[3 + d], plat::d2ord(x));
    }
    a[i] = ar;
  }
};

struct PathCodes {  // octree.cpp:43-51 (octant), :59-81 (child centre), one digit per level
  const double *p;
  double cx, cy, cz, w;
  uint64_t *hi, *lo;
  uint32_t *idx;
  DEVFN void operator()(int64_t i) const {
    const double x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    double c0 = cx, c1 = cy, c2 = cz, wd = w;
    uint64_t code[2] = {0, 0};
    for (int l = 0; l < MAX_LEVELS; ++l) {
      const uint64_t o = (x >= c0 ? 1u : 0u) | (y >= c1 ? 2u : 0u) | (z >= c2 ? 4u : 0u);
      code[l / 21] = (code[l / 21] << 3) | o;
      const double h = wd / 2;          // octree.cpp:103
      const double q = h / 2;           // translate_center(curr_center, h/2, i)
      c0 = c0 + ((o & 1u) ? q : -q);
      c1 = c1 + ((o & 2u) ? q : -q);
      c2 = c2 + ((o & 4u) ? q : -q);
      wd = h;
    }
    hi[i] = code[0]; lo[i] = code[1]; idx[i] = (uint32_t)i;
  }
};

struct GatherKey { const uint64_t *src; const uint32_t *idx; uint64_t *dst; DEVFN void operator()(int64_t i) const { dst[i] = src[idx[i]]; } };

struct Permute {  // gather the per-point arrays into sorted order
  const uint32_t *idx;
  const double *p, *an, *a;
  const uint64_t *hi, *lo;
  double *ps, *ans, *as;
  uint64_t *his, *los;
  DEVFN void operator()(int64_t i) const {
    const uint32_t j = idx[i];
    for (int d = 0; d < 3; ++d) { ps[3 * i + d] = p[3 * j + d]; ans[3 * i + d] = an[3 * j + d]; }
    as[i] = a[j]; his[i] = hi[j]; los[i] = lo[j];
  }
};

DEVFN uint32_t digit_at(const uint64_t *hi, const uint64_t *lo, uint32_t i, int level) {
  return (level < 21) ? (uint32_t)((hi[i] >> (3 * (20 - level))) & 7u) : (uint32_t)((lo[i] >> (3 * (41 - level))) & 7u);
}

struct MarkSplit {  // octree.cpp:98: split while the cell holds more than one point
  const uint32_t *ccount;
  uint32_t first, ncur; int level;
  uint32_t *flag;  // [ncur+1]
  DEVFN void operator()(int64_t k) const {
    flag[k] = (k < ncur && ccount[first + k] > 1 && level < MAX_LEVELS) ? 1u : 0u;
  }
};

struct MakeChildren {  // one thread per (split cell, octant)
  const uint32_t *flag, *rank;
  uint32_t first, next_first; int level;
  const uint64_t *hi, *lo;
  uint32_t *cbegin, *ccount; int32_t *cchild;
  double *ccen, *cwid;
  DEVFN void operator()(int64_t t) const {
    const uint32_t k = (uint32_t)(t >> 3), o = (uint32_t)(t & 7);
    const uint32_t cell = first + k;
    if (!flag[k]) { if (o == 0) cchild[cell] = -1; return; }
    const uint32_t base = next_first + 8u * rank[k];
    if (o == 0) cchild[cell] = (int32_t)base;
    const uint32_t b = cbegin[cell], e = b + ccount[cell];
    // first position in [b,e) whose digit at this level is >= o, and >= o+1
    uint32_t lo0 = b, hi0 = e;
    while (lo0 < hi0) { const uint32_t m = (lo0 + hi0) >> 1; if (digit_at(hi, lo, m, level) < o) lo0 = m + 1; else hi0 = m; }
    uint32_t lo1 = lo0, hi1 = e;
    while (lo1 < hi1) { const uint32_t m = (lo1 + hi1) >> 1; if (digit_at(hi, lo, m, level) < o + 1) lo1 = m + 1; else hi1 = m; }
    const uint32_t c = base + o;
    cbegin[c] = lo0; ccount[c] = lo1 - lo0;
    const double h = cwid[cell] / 2, q = h / 2;
    ccen[3 * c] = ccen[3 * cell] + ((o & 1u) ? q : -q);
    ccen[3 * c + 1] = ccen[3 * cell + 1] + ((o & 2u) ? q : -q);
    ccen[3 * c + 2] = ccen[3 * cell + 2] + ((o & 4u) ? q : -q);
    cwid[c] = h;
  }
};

// sums of one level from the level below (or from the points of a leaf)
struct CellSums {
  uint32_t first; int nterms;
  const uint32_t *cbegin, *ccount; const int32_t *cchild;
  const double *P, *AN, *A;
  double *SA, *SAP;   // sum A, sum A*P [3]
  double *CM, *EC;
  DEVFN void operator()(int64_t k) const {
    const uint32_t c = first + (uint32_t)k;
    double *ec = EC + (size_t)c * nterms;
    for (int t = 0; t < nterms; ++t) ec[t] = 0.0;
    const uint32_t cnt = ccount[c];
    if (cnt == 0) { SA[c] = 0.0; for (int d = 0; d < 3; ++d) { SAP[3 * c + d] = 0.0; CM[3 * c + d] = NAN; } return; }
    const int32_t ch = cchild[c];
    double sa = 0.0, sap[3] = {0, 0, 0}, e0[3] = {0, 0, 0};
    if (ch < 0) {
      const uint32_t b = cbegin[c];
      for (uint32_t j = b; j < b + cnt; ++j) {
        sa += A[j];
        for (int d = 0; d < 3; ++d) { sap[d] += A[j] * P[3 * j + d]; e0[d] += AN[3 * j + d]; }
      }
    } else {
      for (int o = 0; o < 8; ++o) {
        const uint32_t cc = (uint32_t)ch + o;
        if (ccount[cc] == 0) continue;
        sa += SA[cc];
        for (int d = 0; d < 3; ++d) { sap[d] += SAP[3 * cc + d]; e0[d] += EC[(size_t)cc * nterms + d]; }
      }
    }
    double cm[3];
    for (int d = 0; d < 3; ++d) cm[d] = sap[d] / sa;
    SA[c] = sa;
    for (int d = 0; d < 3; ++d) { SAP[3 * c + d] = sap[d]; CM[3 * c + d] = cm[d]; ec[d] = e0[d]; }
    if (nterms < 12) return;
    // E1[i + 3j] = sum (x-cm)_i * (a n)_j ;  E2[9l + i + 3j] = sum 1/2 (x-cm)_l (x-cm)_i (a n)_j
    if (ch < 0) {
      const uint32_t b = cbegin[c];
      for (uint32_t j = b; j < b + cnt; ++j) {
        double pt[3];
        for (int d = 0; d < 3; ++d) pt[d] = P[3 * j + d] - cm[d];
        const double *an = AN + 3 * (size_t)j;
        for (int t = 0; t < 9; ++t) ec[3 + t] += pt[t % 3] * an[t / 3];
        if (nterms == 39)
          for (int l = 0; l < 3; ++l)
            for (int t = 0; t < 9; ++t) ec[12 + 9 * l + t] += 0.5 * pt[l] * (pt[t % 3] * an[t / 3]);
      }
    } else {
      for (int o = 0; o < 8; ++o) {
        const uint32_t cc = (uint32_t)ch + o;
        if (ccount[cc] == 0) continue;
        const double *ce = EC + (size_t)cc * nterms;
        double dl[3];
        for (int d = 0; d < 3; ++d) dl[d] = CM[3 * cc + d] - cm[d];
        for (int t = 0; t < 9; ++t) { const int i = t % 3, j = t / 3; ec[3 + t] += ce[3 + t] + dl[i] * ce[j]; }
        if (nterms == 39)
          for (int l = 0; l < 3; ++l)
            for (int t = 0; t < 9; ++t) {
              const int i = t % 3, j = t / 3;
              ec[12 + 9 * l + t] += ce[12 + 9 * l + t] + 0.5 * dl[l] * ce[3 + i + 3 * j] + 0.5 * dl[i] * ce[3 + l + 3 * j] +
                                    0.5 * dl[l] * dl[i] * ce[j];
            }
      }
    }
  }
};

struct CellRadius {  // R = max |P - CM| over the cell's points: every point visits its ancestors
  const double *P; const uint64_t *hi, *lo;
  const int32_t *cchild; const double *CM;
  unsigned long long *Rbits;  // non-negative doubles order like their bit patterns
  DEVFN void operator()(int64_t i) const {
    const double x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
    uint32_t c = 0;
    for (int level = 0;; ++level) {
      const double dx = x - CM[3 * c], dy = y - CM[3 * c + 1], dz = z - CM[3 * c + 2];
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      unsigned long long u; memcpy(&u, &r, 8);
      plat::atomic_max64(&Rbits[c], u);
      const int32_t ch = cchild[c];
      if (ch < 0) break;
      c = (uint32_t)ch + digit_at(hi, lo, (uint32_t)i, level);
    }
  }
};

DEVFN double direct_eval(const double *loc, const double *an) {  // fast_winding_number.cpp:159-172
  const double ln = sqrt(loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2]);
  if (ln == 0) return 0.5;
  return (loc[0] * an[0] + loc[1] * an[1] + loc[2] * an[2]) / (PI_4 * (ln * ln * ln));
}

DEVFN double expansion_eval(const double *loc, const double *ec, int nterms) {  // :174-228
  double wn = direct_eval(loc, ec);
  const double r = sqrt(loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2]);
  double r5 = 0;
  if (nterms > 3) {
    const double r3 = PI_4 * r * r * r;
    r5 = r3 * r * r;
    const double d = 1.0 / r3, s = -3.0 / r5;
    double acc = 0;
    for (int k = 0; k < 9; ++k) {
      const int u = k % 3, v = k / 3;
      double sd = loc[u] * loc[v] * s;
      if (u == v) sd += d;
      acc += sd * ec[3 + k];
    }
    wn += acc;
  }
  if (nterms > 12) {
    const double r7 = r5 * r * r;
    for (int i = 0; i < 3; ++i) {
      double acc = 0;
      for (int k = 0; k < 9; ++k) {
        const int u = k % 3, v = k / 3;
        double rcd = 0;
        if (u == v) rcd += loc[i];
        if (u == i) rcd += loc[v];
        if (v == i) rcd += loc[u];
        const double td = 15.0 * loc[i] * (loc[u] * (loc[v] / r7)) + (-3.0 / r5) * rcd;
        acc += td * ec[12 + 9 * i + k];
      }
      wn += acc;
    }
  }
  return wn;
}

struct EvalQuery {  // helper() of fast_winding_number.cpp:233-292, depth-first
  const double *P, *AN;
  const uint32_t *cbegin, *ccount; const int32_t *cchild;
  const double *CM, *R, *EC;
  int nterms; int64_t ncloud;
  const double *Q; int64_t nq; double beta; double *WN;
  const uint32_t *perm;   // queries in Morton order: a warp walks (nearly) the same cells
  DEVFN double leaf_sum(uint32_t c, const double *q) const {
    double wn = 0;
    const uint32_t b = cbegin[c], e = b + ccount[c];
    for (uint32_t j = b; j < e; ++j) {
      double loc[3];
      for (int d = 0; d < 3; ++d) loc[d] = P[3 * (size_t)j + d] - q[d];
      wn += direct_eval(loc, AN + 3 * (size_t)j);
    }
    return wn;
  }
  // the cases without a walk: an empty cloud, or brute force (:302-309)
  DEVFN bool no_walk() const { return ncloud == 0 || !(beta > 0); }
  DEVFN void direct(int64_t t) const {
    const int64_t i = perm ? (int64_t)perm[t] : t;
    const double q[3] = {Q[i], Q[nq + i], Q[2 * nq + i]};
    double wn = 0;
    for (int64_t j = 0; j < ncloud; ++j) {
      double loc[3];
      for (int d = 0; d < 3; ++d) loc[d] = P[3 * j + d] - q[d];
      wn += direct_eval(loc, AN + 3 * j);
    }
    WN[i] = wn;
  }
  DEVFN void operator()(int64_t t) const {
    if (no_walk()) { direct(t); return; }
    const int64_t i = perm ? (int64_t)perm[t] : t;
    const double q[3] = {Q[i], Q[nq + i], Q[2 * nq + i]};
    double wn = 0;
    if (cchild[0] < 0) { WN[i] = leaf_sum(0, q); return; }
    uint32_t scell[MAX_LEVELS + 2];
    uint8_t snext[MAX_LEVELS + 2];
    int sp = 0;
    scell[0] = 0; snext[0] = 0;
    while (sp >= 0) {
      const uint32_t c = scell[sp];
      const uint32_t o = snext[sp];
      if (o == 8) { --sp; continue; }
      snext[sp] = (uint8_t)(o + 1);
      const uint32_t ch = (uint32_t)cchild[c] + o;
      if (ccount[ch] == 0) continue;
      double loc[3];
      for (int d = 0; d < 3; ++d) loc[d] = CM[3 * (size_t)ch + d] - q[d];
      const double nrm = sqrt(loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2]);
      const bool leaf = cchild[ch] < 0;
      if (nrm > beta * R[ch]) {
        if (leaf) wn += leaf_sum(ch, q);
        else wn += expansion_eval(loc, EC + (size_t)ch * nterms, nterms);
      } else if (leaf) {
        wn += leaf_sum(ch, q);   // queued in the reference, then summed directly at the next level
      } else {
        ++sp; scell[sp] = ch; snext[sp] = 0;
      }
    }
    WN[i] = wn;
  }
};

template <class T>
inline void upload_and_load(iglb200_cloud *c, const void *P, const void *N, const void *A, int64_t n, plat::DevBuf<double> &p,
                            plat::DevBuf<double> &an, plat::DevBuf<double> &a, unsigned long long *bb, stream_t s) {
  plat::DevBuf<T> rp((size_t)3 * n), rn((size_t)3 * n), ra((size_t)n);
  plat::h2d(rp.p, P, sizeof(T) * 3 * n, s);
  plat::h2d(rn.p, N, sizeof(T) * 3 * n, s);
  plat::h2d(ra.p, A, sizeof(T) * n, s);
  plat::for_n(n, LoadPoints<T>{rp.p, rn.p, ra.p, p.p, an.p, a.p, bb}, s);
  plat::sync(s);
  (void)c;
}

inline iglb200_cloud *create(const void *P, const void *N, const void *A, int is_f64, int64_t n, int order, stream_t s) {
  using plat::DevBuf;
  iglb200_cloud *c = new iglb200_cloud;
  try {
    c->n = n; c->order = order; c->nterms = order == 0 ? 3 : (order == 1 ? 12 : 39);
    if (n == 0) return c;
    if (n >= (int64_t)0x7FFFFFF0) throw std::invalid_argument("too many points");
    DevBuf<double> p((size_t)3 * n), an((size_t)3 * n), a((size_t)n);
    DevBuf<unsigned long long> bb(6);
    {
      unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
      plat::h2d(bb.p, init, sizeof(init), s);
    }
    if (is_f64) upload_and_load<double>(c, P, N, A, n, p, an, a, bb.p, s);
    else upload_and_load<float>(c, P, N, A, n, p, an, a, bb.p, s);
    unsigned long long hb[6];
    plat::d2h(hb, bb.p, sizeof(hb), s);
    plat::sync(s);
    double lo[3], hi[3];
    for (int d = 0; d < 3; ++d) { lo[d] = plat::ord2d(hb[d]); hi[d] = plat::ord2d(hb[3 + d]); c->lo[d] = lo[d]; c->hi[d] = hi[d]; }
    // octree.cpp:143-150
    const double cen[3] = {(lo[0] + hi[0]) / 2.0, (lo[1] + hi[1]) / 2.0, (lo[2] + hi[2]) / 2.0};
    const double width = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));

    DevBuf<uint64_t> khi((size_t)n), klo((size_t)n), k2((size_t)n);
    DevBuf<uint32_t> idx((size_t)n), idx2((size_t)n);
    plat::for_n(n, PathCodes{p.p, cen[0], cen[1], cen[2], width, khi.p, klo.p, idx.p}, s);
    {
      // LSD: stable sort by the low word, then by the high word
      const size_t tb = plat::sort_pairs64_temp_bytes(n);
      DevBuf<char> tmp(tb);
      plat::sort_pairs_u64(tmp.p, tb, klo.p, k2.p, idx.p, idx2.p, n, s);
      // keys of the second pass follow the first pass' permutation
      DevBuf<uint64_t> hperm((size_t)n);
      plat::for_n(n, GatherKey{khi.p, idx2.p, hperm.p}, s);
      plat::sort_pairs_u64(tmp.p, tb, hperm.p, k2.p, idx2.p, idx.p, n, s);
      plat::sync(s);
    }
    c->P.alloc((size_t)3 * n); c->AN.alloc((size_t)3 * n);
    DevBuf<double> As((size_t)n);
    DevBuf<uint64_t> his((size_t)n), los((size_t)n);
    plat::for_n(n, Permute{idx.p, p.p, an.p, a.p, khi.p, klo.p, c->P.p, c->AN.p, As.p, his.p, los.p}, s);
    plat::sync(s);
    p.release(); an.release(); a.release(); khi.release(); klo.release(); k2.release(); idx.release(); idx2.release();

    // cells, level by level
    size_t cap = (size_t)16 * n + 64;
    DevBuf<uint32_t> cbegin(cap), ccount(cap);
    DevBuf<int32_t> cchild(cap);
    DevBuf<double> ccen(3 * cap), cwid(cap);
    {
      const uint32_t b0 = 0, c0 = (uint32_t)n;
      plat::h2d(cbegin.p, &b0, 4, s); plat::h2d(ccount.p, &c0, 4, s);
      plat::h2d(ccen.p, cen, 24, s); plat::h2d(cwid.p, &width, 8, s);
    }
    std::vector<uint32_t> level_first = {0, 1};
    DevBuf<uint32_t> flag, rank;
    for (int level = 0;; ++level) {
      const uint32_t first = level_first[level], ncur = level_first[level + 1] - first;
      flag.alloc((size_t)ncur + 1); rank.alloc((size_t)ncur + 1);
      plat::for_n((int64_t)ncur + 1, MarkSplit{ccount.p, first, ncur, level, flag.p}, s);
      plat::exclusive_scan_u32(flag.p, rank.p, (int64_t)ncur + 1, s);
      uint32_t nsplit = 0;
      plat::d2h(&nsplit, rank.p + ncur, 4, s);
      plat::sync(s);
      const uint32_t next_first = first + ncur;
      if ((size_t)next_first + 8 * (size_t)nsplit > cap) {
        // clustered / duplicate points make long chains of cells: grow geometrically
        const size_t ncap = std::max(cap * 2, (size_t)next_first + 8 * (size_t)nsplit + 64);
        auto grow32 = [&](DevBuf<uint32_t> &b) { DevBuf<uint32_t> t(ncap); plat::d2d(t.p, b.p, 4 * cap, s); plat::sync(s); b = std::move(t); };
        grow32(cbegin); grow32(ccount);
        { DevBuf<int32_t> t(ncap); plat::d2d(t.p, cchild.p, 4 * cap, s); plat::sync(s); cchild = std::move(t); }
        { DevBuf<double> t(3 * ncap); plat::d2d(t.p, ccen.p, 24 * cap, s); plat::sync(s); ccen = std::move(t); }
        { DevBuf<double> t(ncap); plat::d2d(t.p, cwid.p, 8 * cap, s); plat::sync(s); cwid = std::move(t); }
        cap = ncap;
      }
      plat::for_n((int64_t)ncur * 8, MakeChildren{flag.p, rank.p, first, next_first, level, his.p, los.p, cbegin.p, ccount.p, cchild.p, ccen.p, cwid.p}, s);
      if (nsplit == 0) break;
      level_first.push_back(next_first + 8 * nsplit);
    }
    const int64_t nc = level_first.back();
    c->ncells = nc; c->nlevels = (int64_t)level_first.size() - 1;
    c->cbegin.alloc(nc); c->ccount.alloc(nc); c->cchild.alloc(nc);
    plat::d2d(c->cbegin.p, cbegin.p, 4 * (size_t)nc, s); plat::d2d(c->ccount.p, ccount.p, 4 * (size_t)nc, s);
    plat::d2d(c->cchild.p, cchild.p, 4 * (size_t)nc, s);
    plat::sync(s);
    cbegin.release(); ccount.release(); cchild.release(); ccen.release(); cwid.release();
    c->CM.alloc(3 * (size_t)nc); c->R.alloc((size_t)nc); c->EC.alloc((size_t)nc * c->nterms);
    DevBuf<double> SA((size_t)nc), SAP(3 * (size_t)nc);
    for (int level = (int)c->nlevels - 1; level >= 0; --level) {
      const uint32_t first = level_first[level], cnt = level_first[level + 1] - first;
      plat::for_n(cnt, CellSums{first, c->nterms, c->cbegin.p, c->ccount.p, c->cchild.p, c->P.p, c->AN.p, As.p, SA.p, SAP.p, c->CM.p, c->EC.p}, s);
    }
    plat::dev_memset(c->R.p, 0, 8 * (size_t)nc, s);
    plat::for_n(n, CellRadius{c->P.p, his.p, los.p, c->cchild.p, c->CM.p, (unsigned long long *)c->R.p}, s);
    plat::sync(s);
    c->bytes = c->P.bytes() + c->AN.bytes() + c->cbegin.bytes() + c->ccount.bytes() + c->cchild.bytes() + c->CM.bytes() + c->R.bytes() + c->EC.bytes();
  } catch (...) {
    delete c;
    throw;
  }
  return c;
}

inline void destroy(iglb200_cloud *c) { delete c; }

inline void info(const iglb200_cloud *c, int64_t *out) {
  out[0] = c->n; out[1] = c->ncells; out[2] = c->nlevels; out[3] = (int64_t)c->bytes; out[4] = c->order;
  out[5] = out[6] = out[7] = 0;
}

struct QueryKeys {   // 30-bit Morton code of a query inside the cloud's box doubled about its centre
  const double *Q; int64_t n;
  double lo[3], inv[3];
  uint32_t *keys, *vals;
  static DEVFN uint32_t spread(uint32_t v) {
    v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u;
    return v;
  }
  DEVFN void operator()(int64_t i) const {
    uint32_t c[3];
    for (int d = 0; d < 3; ++d) {
      double t = (Q[d * n + i] - lo[d]) * inv[d];
      t = t < 0.0 ? 0.0 : (t > 1023.0 ? 1023.0 : t);
      if (!(t == t)) t = 0.0;
      c[d] = (uint32_t)t;
    }
    keys[i] = spread(c[0]) | (spread(c[1]) << 1) | (spread(c[2]) << 2);
    vals[i] = (uint32_t)i;
  }
};

#ifndef IGLB200_EMUL
#ifndef CLOUD_THREADS
#define CLOUD_THREADS 64
#endif
#ifndef CLOUD_MIN_BLOCKS
#define CLOUD_MIN_BLOCKS 12   // 80 registers, 24 warps per SM: 36.8 ms on c4 (112 regs 47.7, 96 regs 39.6)
#endif
// The walk of EvalQuery for a PACKET: the 32 queries of a warp (neighbours in Morton order) descend the octree together,
// depth-first as each of them would alone, with one shared stack of (cell, next octant, lanes that opened the cell).  A
// lane that did not open a cell waits while the others walk it — as it would in a diverged warp — but cell centres, radii,
// expansion coefficients and leaf points are now warp-uniform loads.  Every lane adds exactly the terms of its own walk in
// the same order, so the result is bit-identical to the per-thread form.
static __device__ double cloud_eval_packet(const EvalQuery &f, bool alive, const double *q, uint4 *stack) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  double wn = 0;
  const unsigned m0 = __ballot_sync(FULL, alive);
  if (m0 == 0u) return 0.0;
  if (f.cchild[0] < 0) return alive ? f.leaf_sum(0, q) : 0.0;
  uint32_t cell = 0, mask = m0, next = 0;
  int sp = 0;
  while (true) {
    if (next == 8) {
      if (sp == 0) break;
      const uint4 e = stack[--sp];
      __syncwarp();
      cell = e.x; mask = e.y; next = e.z;
      continue;
    }
    const uint32_t ch = (uint32_t)f.cchild[cell] + next;
    ++next;
    if (f.ccount[ch] == 0) continue;
    const bool me = (mask >> lane) & 1u;
    double loc[3];
    for (int d = 0; d < 3; ++d) loc[d] = f.CM[3 * (size_t)ch + d] - q[d];
    const double nrm = sqrt(loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2]);
    if (f.cchild[ch] < 0) {   // a leaf is summed directly, near or far (warp-uniform)
      if (me) wn += f.leaf_sum(ch, q);
      continue;
    }
    const bool far = nrm > f.beta * f.R[ch];
    if (me && far) wn += expansion_eval(loc, f.EC + (size_t)ch * f.nterms, f.nterms);
    const unsigned open = __ballot_sync(FULL, me && !far);
    if (open) {
      if (lane == 0) stack[sp] = make_uint4(cell, mask, next, 0u);
      ++sp;
      __syncwarp();
      cell = ch; mask = open; next = 0;
    }
  }
  return wn;
}
static __global__ void __launch_bounds__(CLOUD_THREADS, CLOUD_MIN_BLOCKS) cloud_eval_kernel(int64_t n, EvalQuery f) {
  __shared__ uint4 stacks[CLOUD_THREADS / 32][MAX_LEVELS + 2];
  const bool walk = !f.no_walk();   // otherwise: nothing to sum, or brute force (uniform loads as it is)
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x; b < n; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = b + threadIdx.x;
    if (!walk) { if (t < n) f.direct(t); continue; }
    const bool alive = t < n;
    const int64_t i = alive ? (f.perm ? (int64_t)f.perm[t] : t) : 0;
    double q[3] = {0.0, 0.0, 0.0};
    if (alive) { q[0] = f.Q[i]; q[1] = f.Q[f.nq + i]; q[2] = f.Q[2 * f.nq + i]; }
    const double wn = cloud_eval_packet(f, alive, q, stacks[threadIdx.x >> 5]);
    if (alive) f.WN[i] = wn;
  }
}
#endif

inline void eval(const iglb200_cloud *c, const double *Q, int64_t nq, double beta, double *WN, stream_t s) {
  if (nq <= 0) return;
  if (nq > (int64_t)0x7FFFFFFF) throw std::invalid_argument("more than 2^31-1 queries in one call");
  // grow-only per-thread workspace for the sort (no cudaMalloc on the steady-state call path)
  struct Ws { plat::DevBuf<char> b; };
  plat::DevBuf<char> &ws = plat::per_device<Ws>().b;
  const bool sort = nq >= 1024 && c->n > 0 && beta > 0;
  uint32_t *perm_p = nullptr;
  if (sort) {
    plat::Span sp("morton_sort", s);
    const size_t a = ((size_t)nq * 4 + 255) & ~size_t(255), tb = plat::sort_pairs_temp_bytes(nq);
    if (ws.n < 4 * a + tb) { plat::sync(s); ws.alloc(4 * a + tb); }
    uint32_t *kin_p = (uint32_t *)ws.p, *kout_p = (uint32_t *)(ws.p + a), *vin_p = (uint32_t *)(ws.p + 2 * a);
    perm_p = (uint32_t *)(ws.p + 3 * a);
    struct { uint32_t *p; } kin{kin_p}, kout{kout_p}, vin{vin_p}, perm{perm_p};
    struct { char *p; } tmp{ws.p + 4 * a};
    QueryKeys qk;
    qk.Q = Q; qk.n = nq; qk.keys = kin.p; qk.vals = vin.p;
    for (int d = 0; d < 3; ++d) {
      const double cen = 0.5 * (c->lo[d] + c->hi[d]);
      double ext = c->hi[d] - c->lo[d];
      if (!(ext > 0.0)) ext = 1.0;
      qk.lo[d] = cen - ext; qk.inv[d] = 1024.0 / (2.0 * ext);
    }
    plat::for_n(nq, qk, s);
    plat::sort_pairs_u32(tmp.p, tb, kin.p, kout.p, vin.p, perm.p, nq, 30, s);
  }
  {
    plat::Span sp("cloud_eval", s);
    const EvalQuery f{c->P.p, c->AN.p, c->cbegin.p, c->ccount.p, c->cchild.p, c->CM.p, c->R.p, c->EC.p, c->nterms, c->n, Q, nq, beta, WN,
                      perm_p};
#ifdef IGLB200_EMUL
    plat::for_n(nq, f, s);
#else
    const int64_t blocks = (nq + CLOUD_THREADS - 1) / CLOUD_THREADS;
    cloud_eval_kernel<<<(unsigned)std::min<int64_t>(blocks, 0x7FFFFFFF), CLOUD_THREADS, 0, s>>>(nq, f);
    plat::launch_counter()++;
    PLAT_CHECK(cudaGetLastError());
#endif
  }
  if (sort) plat::sync(s);   // the shared workspace must be idle before the next call (possibly on another stream) reuses it
}

}  // namespace cloud
