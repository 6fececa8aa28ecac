// The mesh handle: one position-independent device arena holding everything a
// query needs, so that replicating the structure to another GPU is a single
// broadcast of [arena, arena+bytes).
//
//   [ArenaHeader | tri64 (9 f64 / face) | tri32 (3 float4 / face) | triq (4 float4 / face) |
//    fwn nodes (384 B) | box nodes (128 B) | optional pseudonormals (21 f64 / face) |
//    optional exact-winding-number tables (wn.h) | pill nodes (160 B): the closest-triangle search]
//
// tri64/tri32 are the faces' vertices gathered once at build time (the
// reference dereferences V through F for every leaf test: AABB.cpp:1413-1432,
// FastWindingNumberForSoups.h:7260-7269).
#pragma once
#include "traverse.h"

static constexpr uint64_t IGLB200_ARENA_MAGIC = 0x32424C47494C4247ull;

struct ArenaHeader {
  uint64_t magic;
  uint64_t bytes;
  int64_t nF, nV, nnodes, nlevels, nexcluded;
  int32_t order, pad;
  double bbox_lo[3], bbox_hi[3];  // of the fp64 geometry
  float absmax, padf;             // max |coordinate| (slack of the fp32 box tests)
  uint64_t off_tri64, off_tri32, off_fnodes, off_cnodes, off_triq;
  uint64_t off_pn;  // 21 f64 / face: face normal, 3 corner (vertex) normals, 3 edge normals; 0 = absent
  // exact winding number (wn.h), 0 = absent: per cluster (4 per node) face count, first leaf slot, first chain
  // entry (4*nnodes+1 entries); chain entries; faces in depth-first leaf order
  uint64_t off_wn_cnt, off_wn_start, off_wn_bstart, off_wn_bent, off_wn_leafface, n_wn_bent;
  double sigma;          // power of two: the fp32 stage of the closest-triangle search (triq, pnodes) is stored in coordinates x sigma,
                         // chosen so that absmax * sigma lies in [0.5, 1)
  uint64_t off_pnodes;   // PillNode (160 B / node): the closest-triangle search walks these; off_cnodes (boxes) serve the exact winding number
};

struct iglb200_mesh {
  ArenaHeader h;
  char *arena = nullptr;  // device
  bool owns = true;
  double build_ms = 0.0;
  int device = 0;
  // copies of the arena on the other devices of the process's group (group.h), indexed by device; the entry of
  // `device` itself is null.  Empty for a single-device handle.
  std::vector<iglb200_mesh *> reps;
  // the handle whose arena lives on device `d`, or null
  const iglb200_mesh *on(int d) const {
    if (d == device) return this;
    return (d >= 0 && (size_t)d < reps.size()) ? reps[(size_t)d] : nullptr;
  }
  int n_devices() const { int k = 1; for (const iglb200_mesh *r : reps) k += r ? 1 : 0; return k; }
  const double *tri64() const { return (const double *)(arena + h.off_tri64); }
  const float4 *tri32() const { return (const float4 *)(arena + h.off_tri32); }
  const float4 *triq() const { return (const float4 *)(arena + h.off_triq); }
  const fwn::FwnNode *fnodes() const { return (const fwn::FwnNode *)(arena + h.off_fnodes); }
  const fwn::CpNode *cnodes() const { return (const fwn::CpNode *)(arena + h.off_cnodes); }
  const fwn::PillNode *pnodes() const { return (const fwn::PillNode *)(arena + h.off_pnodes); }
  const double *pn() const { return h.off_pn ? (const double *)(arena + h.off_pn) : nullptr; }
};

namespace fwn {
// build.cu (compiled with -fmad=false)
iglb200_mesh *create_mesh(const void *V, int v_is_f64, int64_t nV, int64_t v_rs, int64_t v_cs, const int32_t *F,
                          int64_t nF, int64_t f_rs, int64_t f_cs, int order, int flags, stream_t s);
}  // namespace fwn
