// Exact (generalised) winding number for SIGNED_DISTANCE_TYPE_WINDING_NUMBER / _DEFAULT and
// igl::winding_number:
//   igl::solid_angle            /root/reference/include/igl/solid_angle.cpp:10-55
//   igl::winding_number         /root/reference/include/igl/winding_number.cpp:17-89
//   igl::WindingNumberTree      /root/reference/include/igl/WindingNumberTree.h:262-335 (winding_number),
//                               :196-218 (cap from igl::exterior_edges), WindingNumberAABB.h (the hierarchy)
//   gluing                      /root/reference/include/igl/signed_distance.cpp:60-64, 107-113, 611-629
//
// Not a port.  The reference keeps a second binary hierarchy of `new`-ed nodes, each with its own copy
// of its faces and a triangulated "cap" closing its boundary; a query outside a node's box sums the cap
// instead of the node's faces (a closed surface has zero winding number about outside points).  Here
// the one 4-ary tree of the mesh handle is reused: for every internal child cluster the boundary CHAIN
// (grid edges whose orientations do not cancel inside the cluster, with their net multiplicity) is
// built on the device from the half-edges sorted by undirected vertex pair; a query that lies outside
// a cluster's box adds, for each chain edge a->b, the solid angle of the fan triangle (apex, a, b) —
// same identity, no cap mesh stored.  Everything is double precision; the value equals the plain sum
// over all faces up to rounding (~1e-13), which is also what the reference's hierarchy returns.
#pragma once
#include "mesh.h"
#include "fwn_build.h"

namespace wn {
using fwn::EMPTY; using fwn::INTERNAL_BIT;
static constexpr uint32_t NONE = 0xFFFFFFFFu;
static constexpr int STACK = 96;

struct View {   // device pointers into the arena
  const uint32_t *cnt, *start, *bstart, *leafface;   // per cluster (4*node+lane): faces, first leaf slot, first chain entry
  const uint64_t *bent;                              // chain entries: low 32 = 3*face+corner, high 32 = signed multiplicity
};

#ifdef IGLB200_BUILDER_TU   // the build lives in build.cu's translation unit, next to the tree builder
// ---------------------------------------------------------------- build
struct ParentClusters {
  const uint32_t *nodes; uint32_t *parent;
  DEVFN void operator()(int64_t i) const {
    if (i == 0) parent[0] = NONE;
    for (int l = 0; l < 4; ++l) {
      const uint32_t c = nodes[4 * i + l];
      if (c != EMPTY && (c & INTERNAL_BIT)) parent[c & ~INTERNAL_BIT] = (uint32_t)(4 * i + l);
    }
  }
};
struct CountLevel {   // bottom-up: faces below each child lane
  const uint32_t *nodes; uint32_t *cnt; uint32_t first;
  DEVFN void operator()(int64_t k) const {
    const int64_t i = first + k;
    for (int l = 0; l < 4; ++l) {
      const uint32_t c = nodes[4 * i + l];
      uint32_t n = 0;
      if (c == EMPTY) n = 0;
      else if (c & INTERNAL_BIT) { const uint32_t j = c & ~INTERNAL_BIT; n = cnt[4 * j] + cnt[4 * j + 1] + cnt[4 * j + 2] + cnt[4 * j + 3]; }
      else n = 1;
      cnt[4 * i + l] = n;
    }
  }
};
struct StartLevel {   // top-down: depth-first leaf slots, so that every cluster is one contiguous range
  const uint32_t *nodes, *cnt, *parent; uint32_t *start, *leafface, *leafcluster; uint32_t first;
  DEVFN void operator()(int64_t k) const {
    const int64_t i = first + k;
    uint32_t base = i == 0 ? 0u : start[parent[i]];
    for (int l = 0; l < 4; ++l) {
      const uint32_t c = nodes[4 * i + l];
      start[4 * i + l] = base;
      if (c != EMPTY && !(c & INTERNAL_BIT)) { leafface[base] = c; leafcluster[base] = (uint32_t)(4 * i + l); }
      base += cnt[4 * i + l];
    }
  }
};
struct HalfEdges {    // in leaf order: h = 3*slot + corner
  const int32_t *F; int64_t f_rs, f_cs; const uint32_t *leafface;
  uint64_t *key; uint32_t *val; unsigned char *dir;
  DEVFN void operator()(int64_t h) const {
    const uint32_t f = leafface[h / 3];
    const int k = (int)(h % 3);
    const uint32_t a = (uint32_t)F[f * f_rs + k * f_cs], b = (uint32_t)F[f * f_rs + ((k + 1) % 3) * f_cs];
    key[h] = a == b ? ~0ull : (a < b ? ((uint64_t)a << 32 | b) : ((uint64_t)b << 32 | a));
    val[h] = (uint32_t)h;
    dir[h] = a < b ? 1 : 0;
  }
};
// One sorted half-edge walks the clusters above its face.  In each cluster it is the group's first
// member of, it sums the orientations of the group's members inside the cluster; a non-zero net is a
// chain edge of that cluster.  EMIT = false counts, EMIT = true writes at moff[i].
template <bool EMIT> struct ChainWalk {
  const uint64_t *skey; const uint32_t *sval; const unsigned char *dir; int64_t n;
  const uint32_t *leafface, *leafcluster, *parent, *start, *cnt;
  uint32_t *mcount; const uint32_t *moff; uint32_t *ecluster; uint64_t *edata;
  DEVFN void operator()(int64_t i) const {
    const uint64_t key = skey[i];
    uint32_t emitted = 0;
    if (key != ~0ull) {
      const uint32_t h = sval[i], slot = h / 3;
      const bool has_prev = i > 0 && skey[i - 1] == key;
      const uint32_t prev_slot = has_prev ? sval[i - 1] / 3 : 0;
      const int my = dir[h] ? 1 : -1;
      uint32_t c = parent[leafcluster[slot] >> 2];
      while (c != NONE) {
        const uint32_t s = start[c], e = s + cnt[c];
        if (has_prev && prev_slot >= s) break;   // an earlier member lies in this cluster, hence in every one above
        int net = 0;
        for (int64_t j = i; j < n && skey[j] == key && sval[j] / 3 < e; ++j) net += dir[sval[j]] ? 1 : -1;
        if (net != 0) {
          if (EMIT) {
            ecluster[moff[i] + emitted] = c;
            edata[moff[i] + emitted] = (uint64_t)(3 * leafface[slot] + h % 3) | ((uint64_t)(uint32_t)(net * my) << 32);
          }
          ++emitted;
        }
        c = parent[c >> 2];
      }
    }
    if (!EMIT) mcount[i] = emitted;
  }
};
struct GatherEntries { const uint32_t *order; const uint64_t *in; uint64_t *out; DEVFN void operator()(int64_t i) const { out[i] = in[order[i]]; } };
struct Iota { uint32_t *v; DEVFN void operator()(int64_t i) const { v[i] = (uint32_t)i; } };
struct ClusterBegin {   // first sorted entry of every cluster (lower bound); entry [nclusters] = total
  const uint32_t *sorted; uint32_t nent; uint32_t *bstart;
  DEVFN void operator()(int64_t c) const {
    uint32_t lo = 0, hi = nent;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sorted[mid] < (uint32_t)c) lo = mid + 1; else hi = mid; }
    bstart[c] = lo;
  }
};

struct Built {   // temporary device arrays, copied into the arena by create_mesh
  plat::DevBuf<uint32_t> cnt, start, bstart, leafface;
  plat::DevBuf<uint64_t> bent;
  uint64_t nent = 0, nleaf = 0;
};

// cluster = child lane (4 per node): number of faces below it and its contiguous range of depth-first leaf slots
struct Ranges { plat::DevBuf<uint32_t> parent, cnt, start, leafface, leafcluster; };
inline void cluster_ranges(const fwn::Tree &tree, int64_t nF, Ranges &r, stream_t s) {
  const int64_t nn = tree.nnodes, nc = 4 * nn;
  const int levels = (int)tree.level_start.size() - 1;
  r.parent.alloc((size_t)nn); r.leafcluster.alloc((size_t)nF + 1); r.cnt.alloc((size_t)nc); r.start.alloc((size_t)nc); r.leafface.alloc((size_t)nF + 1);
  plat::for_n(nn, ParentClusters{tree.nodes.p, r.parent.p}, s);
  for (int l = levels - 1; l >= 0; --l)
    plat::for_n(tree.level_start[l + 1] - tree.level_start[l], CountLevel{tree.nodes.p, r.cnt.p, tree.level_start[l]}, s);
  for (int l = 0; l < levels; ++l)
    plat::for_n(tree.level_start[l + 1] - tree.level_start[l],
                StartLevel{tree.nodes.p, r.cnt.p, r.parent.p, r.start.p, r.leafface.p, r.leafcluster.p, tree.level_start[l]}, s);
}

inline void build(const fwn::Tree &tree, const Ranges &rg, const int32_t *dF, int64_t f_rs, int64_t f_cs, int64_t nF, Built &out, stream_t s) {
  using plat::DevBuf;
  const int64_t nn = tree.nnodes, nc = 4 * nn;
  if (nn == 0) return;
  if (3 * nF > 0xFFFFFFFFll) throw std::invalid_argument("winding number: more than 2^32/3 faces");
  const uint32_t *parent_p = rg.parent.p, *leafcluster_p = rg.leafcluster.p;
  out.cnt.alloc((size_t)nc); out.start.alloc((size_t)nc); out.bstart.alloc((size_t)nc + 1); out.leafface.alloc((size_t)nF);
  plat::d2d(out.cnt.p, rg.cnt.p, (size_t)4 * nc, s);
  plat::d2d(out.start.p, rg.start.p, (size_t)4 * nc, s);
  plat::d2d(out.leafface.p, rg.leafface.p, (size_t)4 * nF, s);
  struct { const uint32_t *p; } parent{parent_p}, leafcluster{leafcluster_p};
  const int64_t nleaf = nF - tree.nexcluded, nh = 3 * nleaf;
  out.nleaf = (uint64_t)nleaf;
  DevBuf<uint64_t> key((size_t)nh), skey((size_t)nh);
  DevBuf<uint32_t> val((size_t)nh), sval((size_t)nh), mcount((size_t)nh + 1), moff((size_t)nh + 1);
  DevBuf<unsigned char> dir((size_t)nh);
  plat::for_n(nh, HalfEdges{dF, f_rs, f_cs, out.leafface.p, key.p, val.p, dir.p}, s);
  {
    const size_t tb = plat::sort_pairs64_temp_bytes(nh);
    DevBuf<char> tmp(tb);
    plat::sort_pairs_u64(tmp.p, tb, key.p, skey.p, val.p, sval.p, nh, s);   // stable: members of a group stay in leaf order
  }
  key.release(); val.release();
  ChainWalk<false> cw{skey.p, sval.p, dir.p, nh, out.leafface.p, leafcluster.p, parent.p, out.start.p, out.cnt.p, mcount.p, nullptr, nullptr, nullptr};
  plat::for_n(nh, cw, s);
  plat::dev_memset(mcount.p + nh, 0, 4, s);
  plat::exclusive_scan_u32(mcount.p, moff.p, nh + 1, s);
  uint32_t nent = 0;
  plat::d2h(&nent, moff.p + nh, 4, s);
  plat::sync(s);
  out.nent = nent;
  out.bent.alloc((size_t)nent + 1);
  if (nent) {
    DevBuf<uint32_t> ecl((size_t)nent), ecl_sorted((size_t)nent), idx((size_t)nent), idx_sorted((size_t)nent);
    DevBuf<uint64_t> edata((size_t)nent);
    ChainWalk<true> ce{skey.p, sval.p, dir.p, nh, out.leafface.p, leafcluster.p, parent.p, out.start.p, out.cnt.p, nullptr, moff.p, ecl.p, edata.p};
    plat::for_n(nh, ce, s);
    plat::for_n(nent, Iota{idx.p}, s);
    const size_t tb = plat::sort_pairs_temp_bytes(nent);
    DevBuf<char> tmp(tb);
    plat::sort_pairs_u32(tmp.p, tb, ecl.p, ecl_sorted.p, idx.p, idx_sorted.p, nent, 32, s);   // stable: deterministic summation order
    plat::for_n(nent, GatherEntries{idx_sorted.p, edata.p, out.bent.p}, s);
    plat::for_n(nc + 1, ClusterBegin{ecl_sorted.p, nent, out.bstart.p}, s);
    plat::sync(s);
  } else {
    plat::dev_memset(out.bstart.p, 0, (size_t)4 * (nc + 1), s);
    plat::sync(s);
  }
}

#endif  // IGLB200_BUILDER_TU

// ---------------------------------------------------------------- query
// solid_angle.cpp:17-54 in double: the triangle's solid angle over 4 pi, i.e. its winding number contribution
DEVFN double tri_winding(const double *A, const double *B, const double *C, double px, double py, double pz) {
  const double ax = A[0] - px, ay = A[1] - py, az = A[2] - pz;
  const double bx = B[0] - px, by = B[1] - py, bz = B[2] - pz;
  const double cx = C[0] - px, cy = C[1] - py, cz = C[2] - pz;
  const double la = sqrt(ax * ax + ay * ay + az * az), lb = sqrt(bx * bx + by * by + bz * bz), lc = sqrt(cx * cx + cy * cy + cz * cz);
  const double det = ax * by * cz + bx * cy * az + cx * ay * bz - cx * by * az - bx * ay * cz - ax * cy * bz;
  const double dp0 = bx * cx + by * cy + bz * cz, dp1 = cx * ax + cy * ay + cz * az, dp2 = ax * bx + ay * by + az * bz;
  return atan2(det, la * lb * lc + dp0 * la + dp1 * lb + dp2 * lc) / (2.0 * 3.1415926535897932384626433832795);
}

// w(q) = sum over all faces of tri_winding, evaluated hierarchically
DEVFN double winding_number(const fwn::CpNode *nodes, const double *tri64, const View &v, double px, double py, double pz,
                            uint32_t *stack, bool &overflow) {
  double w = 0.0;
  int sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const uint32_t ni = stack[--sp];
    const fwn::CpNode &n = nodes[ni];
    const uint32_t ch[4] = {n.child.x, n.child.y, n.child.z, n.child.w};
    const uint32_t imask = n.pad.x, tmask = n.pad.y;
    for (int l = 0; l < 4; ++l) {
      if ((tmask >> l) & 1u) {
        const double *t = tri64 + 9 * (size_t)ch[l];
        w += tri_winding(t, t + 3, t + 6, px, py, pz);
      } else if ((imask >> l) & 1u) {
        const double lx = ((const float *)&n.lox)[l], ly = ((const float *)&n.loy)[l], lz = ((const float *)&n.loz)[l];
        const double hx = ((const float *)&n.hix)[l], hy = ((const float *)&n.hiy)[l], hz = ((const float *)&n.hiz)[l];
        const bool inside = px >= lx && px <= hx && py >= ly && py <= hy && pz >= lz && pz <= hz;
        const uint32_t c = 4 * ni + l;
        const uint32_t b0 = v.bstart[c], b1 = v.bstart[c + 1], nf = v.cnt[c];
        if (inside) {
          if (sp < STACK) stack[sp++] = ch[l] & ~INTERNAL_BIT; else overflow = true;
        } else if (b1 - b0 < nf) {
          // outside the cluster's box: its faces and the fan over its boundary chain bound a closed surface
          const double apex[3] = {0.5 * (lx + hx), 0.5 * (ly + hy), 0.5 * (lz + hz)};
          double acc = 0.0;
          for (uint32_t k = b0; k < b1; ++k) {
            const uint64_t e = v.bent[k];
            const uint32_t slot = (uint32_t)e;
            const int mult = (int)(uint32_t)(e >> 32);
            const uint32_t f = slot / 3, k0 = slot % 3, k1 = (k0 + 1) % 3;
            const double *t = tri64 + 9 * (size_t)f;
            acc += (double)mult * tri_winding(apex, t + 3 * k0, t + 3 * k1, px, py, pz);
          }
          w += acc;
        } else {   // a chain as long as the face list: sum the faces (WindingNumberTree.h:292, 326-330)
          const uint32_t s0 = v.start[c];
          double acc = 0.0;
          for (uint32_t k = 0; k < nf; ++k) {
            const double *t = tri64 + 9 * (size_t)v.leafface[s0 + k];
            acc += tri_winding(t, t + 3, t + 6, px, py, pz);
          }
          w += acc;
        }
      }
    }
  }
  return w;
}

#ifndef IGLB200_EMUL
// The same evaluation for a PACKET: the 32 queries of a warp (neighbours in Morton order) walk the union of their nodes
// with one shared stack of (node, lanes inside its box).  Triangles, boxes and boundary-chain entries are then
// warp-uniform loads, and the chain / face loops of a cluster run once for all the lanes outside its box instead of once
// per thread at its own pace.  Per query the terms and their ORDER are exactly those of winding_number above, so the result
// is bit-identical to the per-thread walk.
static constexpr int PSTACK = 112;
DEVFN double winding_number_packet(const fwn::CpNode *nodes, const double *tri64, const View &v, bool alive, double px, double py, double pz,
                                   uint2 *stack, uint32_t *overflow) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  double w = 0.0;
  const unsigned m0 = __ballot_sync(FULL, alive);
  if (m0 == 0u) return 0.0;
  int sp = 1;
  if (lane == 0) stack[0] = make_uint2(0u, m0);
  __syncwarp();
  while (sp > 0) {
    const uint2 e = stack[--sp];
    __syncwarp();
    const bool me = (e.y >> lane) & 1u;
    const uint32_t ni = e.x;
    const fwn::CpNode &n = nodes[ni];
    const uint32_t ch[4] = {n.child.x, n.child.y, n.child.z, n.child.w};
    const uint32_t imask = n.pad.x, tmask = n.pad.y;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if ((tmask >> l) & 1u) {   // warp-uniform
        const double *t = tri64 + 9 * (size_t)ch[l];
        if (me) w += tri_winding(t, t + 3, t + 6, px, py, pz);
      } else if ((imask >> l) & 1u) {
        const double lx = ((const float *)&n.lox)[l], ly = ((const float *)&n.loy)[l], lz = ((const float *)&n.loz)[l];
        const double hx = ((const float *)&n.hix)[l], hy = ((const float *)&n.hiy)[l], hz = ((const float *)&n.hiz)[l];
        const bool inside = px >= lx && px <= hx && py >= ly && py <= hy && pz >= lz && pz <= hz;
        const unsigned min_ = __ballot_sync(FULL, me && inside), mout = __ballot_sync(FULL, me && !inside);
        if (min_) {
          if (sp < PSTACK) { if (lane == 0) stack[sp] = make_uint2(ch[l] & ~INTERNAL_BIT, min_); ++sp; }
          else if (lane == 0) *overflow = 1u;
        }
        if (mout) {
          const uint32_t c = 4 * ni + l;
          const uint32_t b0 = v.bstart[c], b1 = v.bstart[c + 1], nf = v.cnt[c];
          const bool mine = (mout >> lane) & 1u;
          double acc = 0.0;
          if (b1 - b0 < nf) {
            // outside the cluster's box: its faces and the fan over its boundary chain bound a closed surface
            const double apex[3] = {0.5 * (lx + hx), 0.5 * (ly + hy), 0.5 * (lz + hz)};
            for (uint32_t k = b0; k < b1; ++k) {
              const uint64_t en = v.bent[k];
              const uint32_t slot = (uint32_t)en;
              const int mult = (int)(uint32_t)(en >> 32);
              const uint32_t f = slot / 3, k0 = slot % 3, k1 = (k0 + 1) % 3;
              const double *t = tri64 + 9 * (size_t)f;
              if (mine) acc += (double)mult * tri_winding(apex, t + 3 * k0, t + 3 * k1, px, py, pz);
            }
          } else {   // a chain as long as the face list: sum the faces (WindingNumberTree.h:292, 326-330)
            const uint32_t s0 = v.start[c];
            for (uint32_t k = 0; k < nf; ++k) {
              const double *t = tri64 + 9 * (size_t)v.leafface[s0 + k];
              if (mine) acc += tri_winding(t, t + 3, t + 6, px, py, pz);
            }
          }
          if (mine) w += acc;
        }
      }
    }
    __syncwarp();
  }
  return w;
}
#endif

}  // namespace wn
