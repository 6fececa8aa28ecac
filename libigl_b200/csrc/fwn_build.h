// Level-synchronous device construction of the 4-ary cluster tree used by the
// soup fast winding number, reproducing the reference's split decisions exactly.
//
// The far-field Taylor expansion has a truncation error (~3e-3) far above the
// parity bar (1e-5), so the device tree must contain the SAME clusters as the
// reference's UT_BVH<4> built with BOX_AREA
//   (/root/reference/include/igl/FastWindingNumberForSoups.h
//    init :4579-4634, initNode :4916-5047, multiSplit :5247-5346,
//    split :5348-5799, nthElement :5845-5898, partitionByCentre :5900-5945).
// Nothing is ported: the reference recurses depth-first on the CPU and
// partitions with a sequential three-way swap loop; here every tree level is
// processed by a handful of kernels, large ranges are binned and partitioned by
// all threads at once, and the sequential swap loop is replaced by its closed
// form (see `scatter_dest`): the permutation that loop produces is a pointer
// chain over prefix counts, which each element can walk on its own.
//
// What must match bit for bit: the ORDER of indices inside every range (ties in
// the small-range regimes depend on it), every fp32 comparison of
// half-surface-area costs, and the span index / pivot arithmetic.  This file is
// compiled with -fmad=false so no multiply-add is contracted.
//
// Node numbering is breadth-first (the reference's is depth-first); numbering
// does not influence any query result.  Child lane order is identical.
#pragma once
#include "platform.h"
#include <cfloat>

namespace fwn {

static constexpr uint32_t EMPTY = 0xFFFFFFFFu;        // FastWindingNumberForSoups.h:4221
static constexpr uint32_t INTERNAL_BIT = 0x80000000u; // :4222
static constexpr uint32_t NSPANS = 16;                // :4408
static constexpr uint32_t TILE = 2048;                // elements per partition tile
static constexpr uint32_t NONE = 0xFFFFFFFFu;

struct Box3 {
  float lo[3], hi[3];
};
// per-triangle box as stored in HBM: one 32-byte sector, two float4 loads
struct alignas(32) TriBox {
  float4 lo, hi;  // .w unused
};

DEVFN void box_empty(Box3 &b) {
  for (int a = 0; a < 3; ++a) { b.lo[a] = FLT_MAX; b.hi[a] = -FLT_MAX; }
}
DEVFN void box_combine(Box3 &b, const Box3 &s) {  // Box::combine :3983-3992
  for (int a = 0; a < 3; ++a) {
    b.lo[a] = (b.lo[a] < s.lo[a]) ? b.lo[a] : s.lo[a];
    b.hi[a] = (b.hi[a] > s.hi[a]) ? b.hi[a] : s.hi[a];
  }
}
DEVFN float half_area(const Box3 &b) {  // :4076-4081
  const float d0 = b.hi[0] - b.lo[0], d1 = b.hi[1] - b.lo[1], d2 = b.hi[2] - b.lo[2];
  return d0 * d1 + d1 * d2 + d2 * d0;
}
DEVFN Box3 load_box(const TriBox *boxes, uint32_t i) {
  const float4 l = boxes[i].lo, h = boxes[i].hi;
  Box3 b;
  b.lo[0] = l.x; b.lo[1] = l.y; b.lo[2] = l.z;
  b.hi[0] = h.x; b.hi[1] = h.y; b.hi[2] = h.z;
  return b;
}
DEVFN float load_key(const TriBox *boxes, uint32_t i, uint32_t axis) {  // utBoxCenter :4501-4505 (min+max)
  const float *l = (const float *)&boxes[i].lo, *h = (const float *)&boxes[i].hi;
  return l[axis] + h[axis];
}

// One tree node being expanded: up to four sub-ranges after three binary splits
struct Task {
  uint32_t node;     // node id to fill
  uint32_t nsub;     // sub-ranges so far
  uint32_t bnd[5];   // boundaries into idx[]
  Box3 sbox[4];
  float sarea[4];
  uint32_t choice;   // sub-range being split in the current phase
};

// A range too large for one thread (> 32 boxes): binned split
struct Job {
  uint32_t task, begin, count, axis;
  Box3 box;
  float axis_min_x2, axis_scale, axis_len;
  uint32_t tile_base, ntiles;
  // decided by Choose
  uint32_t split_index;  // NONE -> quick-select fallback
  float pivot;
  uint32_t left_count;
  // after the partition
  uint32_t nL, nE, split, rebox;
  Box3 lbox, rbox;
};

struct Bins {
  uint32_t cnt[NSPANS];
  uint32_t lo[NSPANS][3], hi[NSPANS][3];  // order-encoded floats
};
struct ReboxAcc {
  uint32_t lo[2][3], hi[2][3];
};

#ifdef IGLB200_BUILDER_TU  // only build.cu (compiled with -fmad=false) instantiates the builder
// ----------------------------------------------------------------------------
// small-range regimes, one thread per range
// ----------------------------------------------------------------------------
struct SplitResult {
  uint32_t split;  // absolute position in idx[]
  Box3 lbox, rbox;
};

// n == 2 (:5351-5356) and n <= 6 exhaustive bipartition (:5363-5453)
DEVFN void split_tiny(const TriBox *boxes, uint32_t *idx, uint32_t begin, uint32_t n, SplitResult &out) {
  if (n == 2) {
    out.lbox = load_box(boxes, idx[begin]);
    out.rbox = load_box(boxes, idx[begin + 1]);
    out.split = begin + 1;
    return;
  }
  Box3 local[6];
  uint32_t lidx[6];
  for (uint32_t i = 0; i < n; ++i) { lidx[i] = idx[begin + i]; local[i] = load_box(boxes, lidx[i]); }
  const uint32_t limit = 1u << (n - 1);
  uint32_t best = NONE;
  float best_h = 0.f;
  for (uint32_t bits = 1; bits < limit; ++bits) {
    Box3 sb[2];
    sb[0] = local[0];
    box_empty(sb[1]);
    uint32_t cnt[2] = {1, 0};
    for (uint32_t bit = 0; bit < n - 1; ++bit) {
      const uint32_t dest = (bits >> bit) & 1u;
      box_combine(sb[dest], local[bit + 1]);
      ++cnt[dest];
    }
    const float h = half_area(sb[0]) * (float)cnt[0] + half_area(sb[1]) * (float)cnt[1];
    if (best == NONE || h < best_h) {
      best = bits; best_h = h; out.lbox = sb[0]; out.rbox = sb[1];
    }
  }
  uint32_t d = begin + 1;
  for (uint32_t bit = 0; bit < n - 1; ++bit)
    if (!((best >> bit) & 1u)) idx[d++] = lidx[bit + 1];
  out.split = d;
  for (uint32_t bit = 0; bit < n - 1; ++bit)
    if ((best >> bit) & 1u) idx[d++] = lidx[bit + 1];
}

// 6 < n <= 32: stable sort by centre along `axis`, sweep all n-1 planes (:5476-5585).
// The reference's chunked insertion sort + std::merge is a stable sort; any
// stable sort yields the same order.
DEVFN void split_mid(const TriBox *boxes, uint32_t *idx, uint32_t begin, uint32_t n, uint32_t axis, SplitResult &out) {
  float key[32];
  uint32_t lidx[32];
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t id = idx[begin + i];
    const float k = load_key(boxes, id, axis);
    uint32_t j = i;
    while (j > 0 && k < key[j - 1]) { key[j] = key[j - 1]; lidx[j] = lidx[j - 1]; --j; }
    key[j] = k; lidx[j] = id;
  }
  for (uint32_t i = 0; i < n; ++i) idx[begin + i] = lidx[i];
  const uint32_t nsplits = n - 1;
  Box3 right[31];
  // NB the reference seeds the suffix sweep with sorted element n-2, not n-1 (:5543-5549):
  // right[k] never contains the last element.  These boxes steer the children's
  // splits, so the quirk is reproduced, not corrected.
  Box3 acc = load_box(boxes, lidx[nsplits - 1]);
  right[nsplits - 1] = acc;
  for (uint32_t i = nsplits - 1; i > 0; --i) { box_combine(acc, load_box(boxes, lidx[i])); right[i - 1] = acc; }
  // left sweep + cost
  Box3 left = load_box(boxes, lidx[0]);
  uint32_t best_split = 0;
  float best_h = half_area(left) + half_area(right[0]) * (float)(n - 1);
  Box3 best_left = left;
  for (uint32_t s = 1; s < nsplits; ++s) {
    box_combine(left, load_box(boxes, lidx[s]));
    const float h = half_area(left) * (float)(s + 1) + half_area(right[s]) * (float)(n - (s + 1));
    if (h < best_h) { best_split = s; best_h = h; best_left = left; }
  }
  out.split = begin + best_split + 1;
  out.lbox = best_left;
  out.rbox = right[best_split];
}

DEVFN uint32_t longest_axis(const Box3 &b, float &len) {  // :5455-5463 first maximum
  uint32_t ax = 0;
  len = b.hi[0] - b.lo[0];
  for (uint32_t a = 1; a < 3; ++a) {
    const float l = b.hi[a] - b.lo[a];
    if (l > len) { ax = a; len = l; }
  }
  return ax;
}

// ----------------------------------------------------------------------------
// closed form of partitionByCentre (:5900-5945)
//
// The reference scans i = 0..n-1 keeping [0,ps) "<pivot", [ps,pe) "==pivot",
// [pe,i) ">pivot" and swaps at most three entries per step.  Reading the swaps
// as moves of single entries:
//   * a "<" entry ends at its rank among the "<" entries (stable);
//   * a ">" entry at position p jumps, when the p-th non-">" entry arrives at
//     position NG[p], to that position; repeat until p >= nL+nE;
//   * a "==" entry is first placed at its rank among non-">" entries, and,
//     when the p-th "<" entry arrives at position L[p], moves to
//     p + (#"==" before L[p]); repeat until p >= nL.
// Positions are relative to the range.  Every jump strictly increases p.
// ----------------------------------------------------------------------------
DEVFN uint32_t scatter_dest(uint32_t cls, uint32_t rel, uint32_t rankL, uint32_t rankNG, uint32_t nL, uint32_t nNG,
                            const uint32_t *Lpos, const uint32_t *NGpos, const uint32_t *ePre) {
  if (cls == 0) return rankL;
  if (cls == 2) {
    uint32_t p = rel;
    while (p < nNG) p = NGpos[p];
    return p;
  }
  uint32_t p = rankNG;
  while (p < nL) p = p + ePre[Lpos[p]];
  return p;
}

// sequential three-way partition, used only inside the quick-select fallback
DEVFN void partition_seq(const TriBox *boxes, uint32_t *a, uint32_t n, uint32_t axis, float pivot, uint32_t &ps_out, uint32_t &pe_out) {
  uint32_t ps = 0, pe = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const float v = load_key(boxes, a[i], axis);
    if (v < pivot) {
      if (i != ps) {
        if (ps == pe) { const uint32_t t = a[i]; a[i] = a[ps]; a[ps] = t; }
        else { const uint32_t t = a[i]; a[i] = a[pe]; a[pe] = a[ps]; a[ps] = t; }
      }
      ++ps; ++pe;
    } else if (v == pivot) {
      if (i != pe) { const uint32_t t = a[i]; a[i] = a[pe]; a[pe] = t; }
      ++pe;
    }
  }
  ps_out = ps; pe_out = pe;
}

// nthElement (:5845-5898): median-of-three quick-select on the centre keys
DEVFN void nth_element_seq(const TriBox *boxes, uint32_t *a, uint32_t n, uint32_t axis, uint32_t nth) {
  uint32_t lo = 0, hi = n;
  while (true) {
    float p0 = load_key(boxes, a[lo], axis);
    float p1 = load_key(boxes, a[lo + (hi - lo) / 2], axis);
    float p2 = load_key(boxes, a[hi - 1], axis);
    if (p0 < p1) { const float t = p0; p0 = p1; p1 = t; }
    if (p0 < p2) { const float t = p0; p0 = p2; p2 = t; }
    if (p1 < p2) { const float t = p1; p1 = p2; p2 = t; }
    uint32_t ps, pe;
    partition_seq(boxes, a + lo, hi - lo, axis, p1, ps, pe);
    ps += lo; pe += lo;
    if (nth < ps) hi = ps;
    else if (nth < pe) return;
    else lo = pe;
    if (hi <= lo + 1) return;
  }
}

// ----------------------------------------------------------------------------
// kernels (functors; see platform.h for the launchers)
// ----------------------------------------------------------------------------
struct TriBoxes {  // per-triangle fp32 boxes (:6566-6590); also flags NaN/Inf boxes (:4488-4500)
  const float4 *tri;  // 3 float4 per face: a,b,c (fp32 vertices, .w unused)
  TriBox *boxes;
  uint32_t *bad;
  DEVFN void operator()(int64_t f) const {
    const float4 a = tri[3 * f], b = tri[3 * f + 1], c = tri[3 * f + 2];
    // SYSmin/SYSmax keep the FIRST argument unless the second is strictly
    // smaller/larger (:253-258), so a NaN in vertex a sticks and one in b or c is dropped
    #define IGLB_MN(x, y) (((x) > (y)) ? (y) : (x))
    #define IGLB_MX(x, y) (((x) < (y)) ? (y) : (x))
    TriBox o;
    o.lo.x = IGLB_MN(IGLB_MN(a.x, b.x), c.x); o.lo.y = IGLB_MN(IGLB_MN(a.y, b.y), c.y); o.lo.z = IGLB_MN(IGLB_MN(a.z, b.z), c.z);
    o.hi.x = IGLB_MX(IGLB_MX(a.x, b.x), c.x); o.hi.y = IGLB_MX(IGLB_MX(a.y, b.y), c.y); o.hi.z = IGLB_MX(IGLB_MX(a.z, b.z), c.z);
    #undef IGLB_MN
    #undef IGLB_MX
    o.lo.w = 0.f; o.hi.w = 0.f;
    const float vals[6] = {o.lo.x, o.lo.y, o.lo.z, o.hi.x, o.hi.y, o.hi.z};
    uint32_t isbad = 0;
    for (int k = 0; k < 6; ++k) isbad |= ((plat::f2u(vals[k]) & 0x7F800000u) == 0x7F800000u) ? 1u : 0u;
    boxes[f] = o;
    bad[f] = isbad;
  }
};

struct CompactGood {  // utExcludeNaNInfBoxIndices (:4535-4577): stable removal
  const uint32_t *bad, *badscan;
  uint32_t *idx;
  DEVFN void operator()(int64_t f) const {
    if (!bad[f]) idx[f - badscan[f]] = (uint32_t)f;
  }
};

struct RootBox {  // computeFullBoundingBox (:4858-4912) via integer atomics on encoded floats
  const TriBox *boxes;
  const uint32_t *idx;
  uint32_t *acc;  // [6]: lo xyz (min), hi xyz (max)
  DEVFN void operator()(int64_t i) const {
    const Box3 b = load_box(boxes, idx[i]);
    for (int a = 0; a < 3; ++a) {
      plat::atomic_min(&acc[a], plat::f2ord(b.lo[a]));
      plat::atomic_max(&acc[3 + a], plat::f2ord(b.hi[a]));
    }
  }
};

struct InitRoot {
  const uint32_t *acc;
  uint32_t n;
  Task *tasks;
  uint32_t *nodes;  // uint4 per node
  uint32_t *idx;
  uint32_t *counters;  // [0]=ntasks(next) [1]=nnodes
  DEVFN void operator()(int64_t) const {
    counters[1] = 1;
    if (n <= 4) {  // initNode :4919-4928
      for (uint32_t i = 0; i < 4; ++i) nodes[i] = (i < n) ? idx[i] : EMPTY;
      counters[0] = 0;
      return;
    }
    Task t{};
    t.node = 0; t.nsub = 1; t.bnd[0] = 0; t.bnd[1] = n; t.choice = 0; t.sarea[0] = 0.f;
    for (int a = 0; a < 3; ++a) { t.sbox[0].lo[a] = plat::ord2f(acc[a]); t.sbox[0].hi[a] = plat::ord2f(acc[3 + a]); }
    tasks[0] = t;
    counters[0] = 1;
  }
};

DEVFN void commit_split(Task &t, uint32_t split, const Box3 &lb, const Box3 &rb) {
  const uint32_t c = t.choice;
  t.bnd[c + 1] = split;
  t.sbox[c] = lb; t.sbox[c + 1] = rb;
  t.sarea[c] = half_area(lb); t.sarea[c + 1] = half_area(rb);
  t.nsub += 1;
}

struct PhasePrepare {  // multiSplit's choice (:5307-5344) + the small regimes of split()
  Task *tasks;
  const TriBox *boxes;
  uint32_t *idx;
  uint32_t *is_large, *tiles_of;  // [ntasks+1]
  uint32_t ntasks;
  DEVFN void operator()(int64_t ti) const {
    if (ti == ntasks) { is_large[ti] = 0; tiles_of[ti] = 0; return; }
    Task t = tasks[ti];
    uint32_t choice = NONE;
    float max_h = 0.f;
    for (uint32_t i = 0; i < t.nsub; ++i) {
      const uint32_t cnt = t.bnd[i + 1] - t.bnd[i];
      if (cnt > 1) {
        const float h = t.sarea[i] * (float)cnt;
        if (choice == NONE || h > max_h) { choice = i; max_h = h; }
      }
    }
    for (uint32_t i = t.nsub; i > choice; --i) t.bnd[i + 1] = t.bnd[i];
    for (uint32_t i = t.nsub - 1; i > choice; --i) { t.sbox[i + 1] = t.sbox[i]; t.sarea[i + 1] = t.sarea[i]; }
    t.choice = choice;
    const uint32_t b = t.bnd[choice], n = t.bnd[choice + 1] - b;   // NB bnd[choice+1] is the range end both before and after the shift
    const Box3 box = t.sbox[choice];
    uint32_t large = 0;
    SplitResult r;
    if (n <= 6) {
      split_tiny(boxes, idx, b, n, r);
      commit_split(t, r.split, r.lbox, r.rbox);
    } else {
      float len;
      const uint32_t axis = longest_axis(box, len);
      if (!(len > 0.f)) {  // :5465-5472
        commit_split(t, b + n / 2, box, box);
      } else if (n <= 2 * NSPANS) {
        split_mid(boxes, idx, b, n, axis, r);
        commit_split(t, r.split, r.lbox, r.rbox);
      } else {
        large = 1;
      }
    }
    tasks[ti] = t;
    is_large[ti] = large;
    tiles_of[ti] = large ? (n + TILE - 1) / TILE : 0;
  }
};

struct JobSetup {
  const Task *tasks;
  const uint32_t *is_large, *job_of, *tile_base_of;
  Job *jobs;
  Bins *bins;
  ReboxAcc *rebox;
  DEVFN void operator()(int64_t ti) const {
    if (!is_large[ti]) return;
    const Task &t = tasks[ti];
    const uint32_t j = job_of[ti];
    Job jb;
    jb.task = (uint32_t)ti;
    jb.begin = t.bnd[t.choice];
    jb.count = t.bnd[t.choice + 1] - jb.begin;
    jb.box = t.sbox[t.choice];
    float len;
    jb.axis = longest_axis(jb.box, len);
    jb.axis_len = len;
    jb.axis_min_x2 = 2.0f * jb.box.lo[jb.axis];           // :5596
    jb.axis_scale = (0.5f * (float)NSPANS) / len;         // :5598
    jb.tile_base = tile_base_of[ti];
    jb.ntiles = (jb.count + TILE - 1) / TILE;
    jb.split_index = NONE; jb.pivot = 0.f; jb.left_count = 0; jb.nL = jb.nE = 0; jb.split = 0; jb.rebox = 0;
    jb.lbox = jb.box; jb.rbox = jb.box;
    jobs[j] = jb;
    Bins z;
    for (uint32_t s = 0; s < NSPANS; ++s) {
      z.cnt[s] = 0;
      for (int a = 0; a < 3; ++a) { z.lo[s][a] = 0xFFFFFFFFu; z.hi[s][a] = 0u; }
    }
    bins[j] = z;
    ReboxAcc ra;
    for (int s = 0; s < 2; ++s) for (int a = 0; a < 3; ++a) { ra.lo[s][a] = 0xFFFFFFFFu; ra.hi[s][a] = 0u; }
    rebox[j] = ra;
  }
};

DEVFN uint32_t find_job(const Job *jobs, uint32_t njobs, uint32_t tile) {  // last job with tile_base <= tile
  uint32_t lo = 0, hi = njobs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (jobs[mid].tile_base <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

DEVFN uint32_t span_of(float key, float axis_min_x2, float axis_scale) {  // :5610
  int s = (int)((key - axis_min_x2) * axis_scale);
  s = s < 0 ? 0 : (s > (int)NSPANS - 1 ? (int)NSPANS - 1 : s);
  return (uint32_t)s;
}

struct BinTiles {  // span counts and boxes (:5604-5660)
  const Job *jobs;
  const uint32_t *njobs;
  const TriBox *boxes;
  const uint32_t *idx;
  Bins *bins;
  DEVFN void tile(uint32_t t) const {
    const uint32_t j = find_job(jobs, *njobs, t);
    const Job &jb = jobs[j];
    const uint32_t off = (t - jb.tile_base) * TILE;
    const uint32_t end = (off + TILE < jb.count) ? off + TILE : jb.count;
#ifdef IGLB200_EMUL
    Bins &loc = bins[j];
    for (uint32_t e = off; e < end; ++e) {
      const Box3 b = load_box(boxes, idx[jb.begin + e]);
      const uint32_t s = span_of(b.lo[jb.axis] + b.hi[jb.axis], jb.axis_min_x2, jb.axis_scale);
      loc.cnt[s]++;
      for (int a = 0; a < 3; ++a) {
        plat::atomic_min(&loc.lo[s][a], plat::f2ord(b.lo[a]));
        plat::atomic_max(&loc.hi[s][a], plat::f2ord(b.hi[a]));
      }
    }
#else
    __shared__ Bins sh;
    uint32_t *shw = (uint32_t *)&sh;
    for (uint32_t k = threadIdx.x; k < sizeof(Bins) / 4; k += blockDim.x) {
      // cnt -> 0, lo -> 0xFFFFFFFF, hi -> 0
      shw[k] = (k < NSPANS) ? 0u : (((k - NSPANS) / 3) % 2 == 0 ? 0u : 0u);
    }
    __syncthreads();
    if (threadIdx.x < NSPANS) for (int a = 0; a < 3; ++a) { sh.lo[threadIdx.x][a] = 0xFFFFFFFFu; sh.hi[threadIdx.x][a] = 0u; }
    __syncthreads();
    const uint32_t axis = jb.axis;
    const float m2 = jb.axis_min_x2, sc = jb.axis_scale;
    for (uint32_t e = off + threadIdx.x; e < end; e += blockDim.x) {
      const Box3 b = load_box(boxes, idx[jb.begin + e]);
      const uint32_t s = span_of(b.lo[axis] + b.hi[axis], m2, sc);
      atomicAdd(&sh.cnt[s], 1u);
      for (int a = 0; a < 3; ++a) {
        atomicMin(&sh.lo[s][a], plat::f2ord(b.lo[a]));
        atomicMax(&sh.hi[s][a], plat::f2ord(b.hi[a]));
      }
    }
    __syncthreads();
    Bins &g = bins[j];
    if (threadIdx.x < NSPANS && sh.cnt[threadIdx.x]) {
      const uint32_t s = threadIdx.x;
      atomicAdd(&g.cnt[s], sh.cnt[s]);
      for (int a = 0; a < 3; ++a) { atomicMin(&g.lo[s][a], sh.lo[s][a]); atomicMax(&g.hi[s][a], sh.hi[s][a]); }
    }
#endif
  }
};

struct Choose {  // the 15 candidate planes (:5662-5704) and the pivot (:5756)
  Job *jobs;
  const Bins *bins;
  DEVFN void operator()(int64_t j) const {
    Job jb = jobs[j];
    const Bins &bn = bins[j];
    Box3 span[NSPANS];
    for (uint32_t s = 0; s < NSPANS; ++s) {
      if (bn.cnt[s]) for (int a = 0; a < 3; ++a) { span[s].lo[a] = plat::ord2f(bn.lo[s][a]); span[s].hi[a] = plat::ord2f(bn.hi[s][a]); }
      else box_empty(span[s]);
    }
    Box3 left[NSPANS - 1], right[NSPANS - 1];
    Box3 acc = span[0];
    left[0] = acc;
    for (uint32_t i = 1; i < NSPANS - 1; ++i) { box_combine(acc, span[i]); left[i] = acc; }
    acc = span[NSPANS - 1];
    right[NSPANS - 2] = acc;
    for (uint32_t i = NSPANS - 2; i > 0; --i) { box_combine(acc, span[i]); right[i - 1] = acc; }
    uint32_t left_counts[NSPANS - 1];
    uint32_t c = bn.cnt[0];
    left_counts[0] = c;
    for (uint32_t i = 1; i < NSPANS - 1; ++i) { c += bn.cnt[i]; left_counts[i] = c; }
    const uint32_t n = jb.count;
    const uint32_t min_count = n / 16;
    const uint32_t max_count = (uint32_t)((15ull * (uint64_t)n) / 16ull);
    float smallest = INFINITY;
    uint32_t split_index = NONE;
    for (uint32_t s = 0; s < NSPANS - 1; ++s) {
      const uint32_t lc = left_counts[s];
      if (lc < min_count || lc > max_count) continue;
      const uint32_t rc = n - lc;
      const float h = (float)lc * half_area(left[s]) + (float)rc * half_area(right[s]);
      if (h < smallest) { smallest = h; split_index = s; }
    }
    jb.split_index = split_index;
    if (split_index != NONE) {
      jb.pivot = jb.axis_min_x2 + ((float)(split_index + 1) * jb.axis_len) / 8.0f;  // :5756
      jb.left_count = left_counts[split_index];
      jb.lbox = left[split_index];
      jb.rbox = right[split_index];
    } else {
      // quick-select target (:5708-5741), executed sequentially in Finish
      if (left_counts[0] > max_count) jb.left_count = max_count;
      else if (left_counts[NSPANS - 2] < min_count) jb.left_count = min_count;
      else jb.left_count = n / 2;
    }
    jobs[j] = jb;
  }
};

DEVFN uint32_t classify(const Job &jb, const TriBox *boxes, uint32_t id) {
  if (jb.split_index == NONE) return 0;  // fallback jobs keep their order here
  const float k = load_key(boxes, id, jb.axis);
  return (k < jb.pivot) ? 0u : ((k == jb.pivot) ? 1u : 2u);
}

struct CountTiles {
  const Job *jobs;
  const uint32_t *njobs;
  const TriBox *boxes;
  const uint32_t *idx;
  uint32_t *tileL, *tileE;
  DEVFN void tile(uint32_t t) const {
    const uint32_t j = find_job(jobs, *njobs, t);
    const Job &jb = jobs[j];
    const uint32_t off = (t - jb.tile_base) * TILE;
    const uint32_t end = (off + TILE < jb.count) ? off + TILE : jb.count;
#ifdef IGLB200_EMUL
    uint32_t nl = 0, ne = 0;
    for (uint32_t e = off; e < end; ++e) {
      const uint32_t c = classify(jb, boxes, idx[jb.begin + e]);
      nl += (c == 0); ne += (c == 1);
    }
    tileL[t] = nl; tileE[t] = ne;
#else
    uint32_t nl = 0, ne = 0;
    for (uint32_t e = off + threadIdx.x; e < end; e += blockDim.x) {
      const uint32_t c = classify(jb, boxes, idx[jb.begin + e]);
      nl += (c == 0); ne += (c == 1);
    }
    typedef cub::BlockReduce<uint32_t, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const uint32_t sl = BR(tmp).Sum(nl);
    __syncthreads();
    const uint32_t se = BR(tmp).Sum(ne);
    if (threadIdx.x == 0) { tileL[t] = sl; tileE[t] = se; }
#endif
  }
};

struct JobTotals {  // nL, nE per job from the tile scans (arrays have ntiles+1 entries)
  Job *jobs;
  const uint32_t *scanL, *scanE;
  DEVFN void operator()(int64_t j) const {
    Job &jb = jobs[j];
    jb.nL = scanL[jb.tile_base + jb.ntiles] - scanL[jb.tile_base];
    jb.nE = scanE[jb.tile_base + jb.ntiles] - scanE[jb.tile_base];
  }
};

// per element: class and ranks; publishes Lpos / NGpos / ePre (indexed from job.begin)
struct RankTiles {
  const Job *jobs;
  const uint32_t *njobs;
  const TriBox *boxes;
  const uint32_t *idx;
  const uint32_t *scanL, *scanE;
  uint32_t *Lpos, *NGpos, *ePre;
  uint32_t *cls_rank;  // per element: cls<<30 | rankL (cls 0) / rankNG (cls 1) / 0
  DEVFN void tile(uint32_t t) const {
    const uint32_t j = find_job(jobs, *njobs, t);
    const Job &jb = jobs[j];
    const uint32_t off = (t - jb.tile_base) * TILE;
    const uint32_t end = (off + TILE < jb.count) ? off + TILE : jb.count;
    uint32_t baseL = scanL[t] - scanL[jb.tile_base];
    uint32_t baseE = scanE[t] - scanE[jb.tile_base];
#ifdef IGLB200_EMUL
    for (uint32_t e = off; e < end; ++e) {
      const uint32_t c = classify(jb, boxes, idx[jb.begin + e]);
      ePre[jb.begin + e] = baseE;
      if (c == 0) { Lpos[jb.begin + baseL] = e; NGpos[jb.begin + baseL + baseE] = e; cls_rank[jb.begin + e] = baseL; ++baseL; }
      else if (c == 1) { NGpos[jb.begin + baseL + baseE] = e; cls_rank[jb.begin + e] = (1u << 30) | (baseL + baseE); ++baseE; }
      else cls_rank[jb.begin + e] = (2u << 30);
    }
#else
    typedef cub::BlockScan<uint32_t, 256> BS;
    __shared__ typename BS::TempStorage tmp;
    for (uint32_t r = off; r < end; r += 256) {
      const uint32_t e = r + threadIdx.x;
      uint32_t c = 3;
      if (e < end) c = classify(jb, boxes, idx[jb.begin + e]);
      const uint32_t packed = (c == 0 ? 1u : 0u) | (c == 1 ? (1u << 16) : 0u);
      uint32_t ex, total;
      BS(tmp).ExclusiveSum(packed, ex, total);
      __syncthreads();
      const uint32_t myL = baseL + (ex & 0xFFFFu), myE = baseE + (ex >> 16);
      if (e < end) {
        ePre[jb.begin + e] = myE;
        if (c == 0) { Lpos[jb.begin + myL] = e; NGpos[jb.begin + myL + myE] = e; cls_rank[jb.begin + e] = myL; }
        else if (c == 1) { NGpos[jb.begin + myL + myE] = e; cls_rank[jb.begin + e] = (1u << 30) | (myL + myE); }
        else cls_rank[jb.begin + e] = (2u << 30);
      }
      baseL += total & 0xFFFFu; baseE += total >> 16;
    }
#endif
  }
};

struct ScatterTiles {
  const Job *jobs;
  const uint32_t *njobs;
  const uint32_t *idx;
  const uint32_t *Lpos, *NGpos, *ePre, *cls_rank;
  uint32_t *idx_tmp;
  DEVFN void one(const Job &jb, uint32_t e) const {
    const uint32_t cr = cls_rank[jb.begin + e];
    const uint32_t c = cr >> 30, rk = cr & 0x3FFFFFFFu;
    const uint32_t d = scatter_dest(c, e, rk, rk, jb.nL, jb.nL + jb.nE, Lpos + jb.begin, NGpos + jb.begin, ePre + jb.begin);
    idx_tmp[jb.begin + d] = idx[jb.begin + e];
  }
  DEVFN void tile(uint32_t t) const {
    const uint32_t j = find_job(jobs, *njobs, t);
    const Job &jb = jobs[j];
    const uint32_t off = (t - jb.tile_base) * TILE;
    const uint32_t end = (off + TILE < jb.count) ? off + TILE : jb.count;
#ifdef IGLB200_EMUL
    for (uint32_t e = off; e < end; ++e) one(jb, e);
#else
    for (uint32_t e = off + threadIdx.x; e < end; e += blockDim.x) one(jb, e);
#endif
  }
};

struct CopyBackTiles {
  const Job *jobs;
  const uint32_t *njobs;
  uint32_t *idx;
  const uint32_t *idx_tmp;
  DEVFN void tile(uint32_t t) const {
    const uint32_t j = find_job(jobs, *njobs, t);
    const Job &jb = jobs[j];
    const uint32_t off = (t - jb.tile_base) * TILE;
    const uint32_t end = (off + TILE < jb.count) ? off + TILE : jb.count;
#ifdef IGLB200_EMUL
    for (uint32_t e = off; e < end; ++e) idx[jb.begin + e] = idx_tmp[jb.begin + e];
#else
    for (uint32_t e = off + threadIdx.x; e < end; e += blockDim.x) idx[jb.begin + e] = idx_tmp[jb.begin + e];
#endif
  }
};

struct Finish {  // where the range is cut (:5760-5798) + the rare quick-select path (:5708-5754)
  Job *jobs;
  const TriBox *boxes;
  uint32_t *idx;
  DEVFN void operator()(int64_t j) const {
    Job &jb = jobs[j];
    if (jb.split_index == NONE) {
      nth_element_seq(boxes, idx + jb.begin, jb.count, jb.axis, jb.left_count);
      jb.split = jb.left_count;
      jb.rebox = 1;
      return;
    }
    uint32_t split = jb.left_count;
    const uint32_t ps = jb.nL, pe = jb.nL + jb.nE;
    if (split >= ps && split <= pe) { jb.split = split; jb.rebox = 0; return; }
    split = (split < ps) ? ps : pe;
    if (split == 0) split = 1;
    else if (split == jb.count) split = jb.count - 1;
    jb.split = split;
    jb.rebox = 1;
  }
};

struct ReboxTiles {  // recompute both side boxes when the binned ones do not apply
  const Job *jobs;
  const uint32_t *njobs;
  const TriBox *boxes;
  const uint32_t *idx;
  ReboxAcc *acc;
  DEVFN void tile(uint32_t t) const {
    const uint32_t j = find_job(jobs, *njobs, t);
    const Job &jb = jobs[j];
    if (!jb.rebox) return;
    const uint32_t off = (t - jb.tile_base) * TILE;
    const uint32_t end = (off + TILE < jb.count) ? off + TILE : jb.count;
    ReboxAcc &g = acc[j];
#ifdef IGLB200_EMUL
    for (uint32_t e = off; e < end; ++e) {
#else
    for (uint32_t e = off + threadIdx.x; e < end; e += blockDim.x) {
#endif
      const Box3 b = load_box(boxes, idx[jb.begin + e]);
      const int s = (e < jb.split) ? 0 : 1;
      for (int a = 0; a < 3; ++a) {
        plat::atomic_min(&g.lo[s][a], plat::f2ord(b.lo[a]));
        plat::atomic_max(&g.hi[s][a], plat::f2ord(b.hi[a]));
      }
    }
  }
};

struct CommitJobs {
  Job *jobs;
  const ReboxAcc *acc;
  Task *tasks;
  DEVFN void operator()(int64_t j) const {
    Job &jb = jobs[j];
    if (jb.rebox) {
      for (int a = 0; a < 3; ++a) {
        jb.lbox.lo[a] = plat::ord2f(acc[j].lo[0][a]); jb.lbox.hi[a] = plat::ord2f(acc[j].hi[0][a]);
        jb.rbox.lo[a] = plat::ord2f(acc[j].lo[1][a]); jb.rbox.hi[a] = plat::ord2f(acc[j].hi[1][a]);
      }
    }
    commit_split(tasks[jb.task], jb.begin + jb.split, jb.lbox, jb.rbox);
  }
};

// after the third split: how many child nodes does each task create?
struct CountChildren {
  const Task *tasks;
  uint32_t ntasks;
  uint32_t *nchild_nodes, *nchild_tasks;  // [ntasks+1]
  DEVFN void operator()(int64_t ti) const {
    if (ti == ntasks) { nchild_nodes[ti] = 0; nchild_tasks[ti] = 0; return; }
    const Task &t = tasks[ti];
    uint32_t nn = 0, nt = 0;
    for (uint32_t i = 0; i < 4; ++i) {
      const uint32_t c = t.bnd[i + 1] - t.bnd[i];
      if (c != 1) { ++nn; if (c > 4) ++nt; }
    }
    nchild_nodes[ti] = nn; nchild_tasks[ti] = nt;
  }
};

struct EmitChildren {  // initNode's tail (:4942-4949, :5036-5045)
  const Task *tasks;
  const uint32_t *node_off, *task_off;  // exclusive scans
  uint32_t ntasks, node_base;
  const uint32_t *idx;
  uint32_t *nodes;
  Task *next;
  uint32_t *counters;
  DEVFN void operator()(int64_t ti) const {
    const Task &t = tasks[ti];
    uint32_t nn = node_base + node_off[ti], nt = task_off[ti];
    for (uint32_t i = 0; i < 4; ++i) {
      const uint32_t b = t.bnd[i], c = t.bnd[i + 1] - b;
      if (c == 1) { nodes[4 * t.node + i] = idx[b]; continue; }
      const uint32_t id = nn++;
      nodes[4 * t.node + i] = id | INTERNAL_BIT;
      if (c <= 4) {
        for (uint32_t k = 0; k < 4; ++k) nodes[4 * id + k] = (k < c) ? idx[b + k] : EMPTY;
      } else {
        Task ch{};   // zero-filled: the whole record is read back by PhasePrepare
        ch.node = id; ch.nsub = 1; ch.bnd[0] = b; ch.bnd[1] = b + c; ch.sbox[0] = t.sbox[i]; ch.choice = 0; ch.sarea[0] = 0.f;
        next[nt++] = ch;
      }
    }
    if (ti + 1 == ntasks) { counters[0] = task_off[ntasks]; counters[1] = node_base + node_off[ntasks]; }
  }
};

// ----------------------------------------------------------------------------
// host driver
// ----------------------------------------------------------------------------
struct Tree {
  plat::DevBuf<uint32_t> nodes;  // 4 per node
  uint32_t nnodes = 0;
  uint32_t nexcluded = 0;        // faces dropped because their box holds NaN/Inf (:4606-4615)
  std::vector<uint32_t> level_start;  // breadth-first: level l = [level_start[l], level_start[l+1])
};

inline void build_tree(const float4 *tri32, uint32_t nF, plat::DevBuf<TriBox> &boxes, Tree &out, stream_t s) {
  using plat::DevBuf;
  out.level_start.clear();
  out.nnodes = 0;
  if (nF == 0) return;
  boxes.alloc(nF);
  DevBuf<uint32_t> bad(nF + 1), badscan(nF + 1), idx(nF), idx_tmp(nF);
  plat::dev_memset(bad.p + nF, 0, 4, s);
  plat::for_n(nF, TriBoxes{tri32, boxes.p, bad.p}, s);
  plat::exclusive_scan_u32(bad.p, badscan.p, nF + 1, s);
  uint32_t nbad = 0;
  plat::d2h(&nbad, badscan.p + nF, 4, s);
  plat::sync(s);
  const uint32_t n = nF - nbad;
  out.nexcluded = nbad;
  if (n == 0) return;
  plat::for_n(nF, CompactGood{bad.p, badscan.p, idx.p}, s);

  DevBuf<uint32_t> acc(6), counters(4);
  {
    const uint32_t init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    plat::h2d(acc.p, init, sizeof(init), s);
  }
  plat::for_n(n, RootBox{boxes.p, idx.p, acc.p}, s);
  const size_t max_tasks = (size_t)n / 5 + 2;
  const size_t max_jobs = (size_t)n / 33 + 2;
  const size_t max_tiles = (size_t)n / TILE + max_jobs + 2;
  out.nodes.alloc((size_t)4 * ((size_t)n + 1));
  DevBuf<Task> tasksA(max_tasks), tasksB(max_tasks);
  DevBuf<uint32_t> is_large(max_tasks + 1), tiles_of(max_tasks + 1), job_of(max_tasks + 1), tile_base_of(max_tasks + 1);
  DevBuf<Job> jobs(max_jobs);
  DevBuf<Bins> bins(max_jobs);
  DevBuf<ReboxAcc> rebox(max_jobs);
  DevBuf<uint32_t> tileL(max_tiles + 1), tileE(max_tiles + 1), scanL(max_tiles + 1), scanE(max_tiles + 1);
  DevBuf<uint32_t> Lpos(n), NGpos(n), ePre(n), cls_rank(n);
  DevBuf<uint32_t> njobs_d(2);

  plat::for_n(1, InitRoot{acc.p, n, tasksA.p, out.nodes.p, idx.p, counters.p}, s);
  uint32_t hc[2];
  plat::d2h(hc, counters.p, 8, s);
  plat::sync(s);
  uint32_t ntasks = hc[0], nnodes = hc[1];
  out.level_start.push_back(0);
  out.level_start.push_back(nnodes);
  Task *cur = tasksA.p, *nxt = tasksB.p;
  while (ntasks > 0) {
    for (int phase = 0; phase < 3; ++phase) {
      plat::for_n((int64_t)ntasks + 1, PhasePrepare{cur, boxes.p, idx.p, is_large.p, tiles_of.p, ntasks}, s);
      plat::exclusive_scan_u32(is_large.p, job_of.p, (int64_t)ntasks + 1, s);
      plat::exclusive_scan_u32(tiles_of.p, tile_base_of.p, (int64_t)ntasks + 1, s);
      uint32_t hj[2];
      plat::d2h(&hj[0], job_of.p + ntasks, 4, s);
      plat::d2h(&hj[1], tile_base_of.p + ntasks, 4, s);
      plat::sync(s);
      const uint32_t njobs = hj[0], ntiles = hj[1];
      if (njobs == 0) continue;
      plat::h2d(njobs_d.p, hj, 8, s);
      const uint32_t *nj = njobs_d.p, *ntl = njobs_d.p + 1;
      plat::for_n(ntasks, JobSetup{cur, is_large.p, job_of.p, tile_base_of.p, jobs.p, bins.p, rebox.p}, s);
      plat::tiles_dyn(ntl, BinTiles{jobs.p, nj, boxes.p, idx.p, bins.p}, s);
      plat::for_n(njobs, Choose{jobs.p, bins.p}, s);
      plat::tiles_dyn(ntl, CountTiles{jobs.p, nj, boxes.p, idx.p, tileL.p, tileE.p}, s);
      plat::dev_memset(tileL.p + ntiles, 0, 4, s);
      plat::dev_memset(tileE.p + ntiles, 0, 4, s);
      plat::exclusive_scan_u32(tileL.p, scanL.p, (int64_t)ntiles + 1, s);
      plat::exclusive_scan_u32(tileE.p, scanE.p, (int64_t)ntiles + 1, s);
      plat::for_n(njobs, JobTotals{jobs.p, scanL.p, scanE.p}, s);
      plat::tiles_dyn(ntl, RankTiles{jobs.p, nj, boxes.p, idx.p, scanL.p, scanE.p, Lpos.p, NGpos.p, ePre.p, cls_rank.p}, s);
      plat::tiles_dyn(ntl, ScatterTiles{jobs.p, nj, idx.p, Lpos.p, NGpos.p, ePre.p, cls_rank.p, idx_tmp.p}, s);
      plat::tiles_dyn(ntl, CopyBackTiles{jobs.p, nj, idx.p, idx_tmp.p}, s);
      plat::for_n(njobs, Finish{jobs.p, boxes.p, idx.p}, s);
      plat::tiles_dyn(ntl, ReboxTiles{jobs.p, nj, boxes.p, idx.p, rebox.p}, s);
      plat::for_n(njobs, CommitJobs{jobs.p, rebox.p, cur}, s);
    }
    // is_large/tiles_of double as scratch for the child counts
    plat::for_n((int64_t)ntasks + 1, CountChildren{cur, ntasks, is_large.p, tiles_of.p}, s);
    plat::exclusive_scan_u32(is_large.p, job_of.p, (int64_t)ntasks + 1, s);
    plat::exclusive_scan_u32(tiles_of.p, tile_base_of.p, (int64_t)ntasks + 1, s);
    plat::for_n(ntasks, EmitChildren{cur, job_of.p, tile_base_of.p, ntasks, nnodes, idx.p, out.nodes.p, nxt, counters.p}, s);
    plat::d2h(hc, counters.p, 8, s);
    plat::sync(s);
    ntasks = hc[0];
    nnodes = hc[1];
    out.level_start.push_back(nnodes);
    Task *t = cur; cur = nxt; nxt = t;
  }
  out.nnodes = nnodes;
}

#endif  // IGLB200_BUILDER_TU

}  // namespace fwn
