// Per-query device functions for the two traversals of the hot path.
//
//  closest_triangle : branch-and-bound over the 4-wide tree.  fp32 conservative
//      box tests prune; every triangle that survives is evaluated with
//      Ericson's routine in fp64 with un-fused arithmetic in the reference's
//      operation order (/root/reference/include/igl/
//      point_simplex_squared_distance.cpp:45-110,134), so squared distances are
//      bit-identical to the CPU restatement and ties can be compared exactly.
//      Tie rule: lowest face index (the reference keeps the first triangle its
//      own tree happens to visit, AABB.cpp:787-802 — order dependent).
//  winding_solid_angle : depth-first Barnes–Hut evaluation
//      (FastWindingNumberForSoups.h:7148-7284).  The descend test and the exact
//      triangle term use IEEE un-fused fp32 in the reference's order so that
//      every branch agrees with it; the far-field polynomial may use FMA.
#pragma once
#include "moments.h"

namespace fwn {

#ifdef IGLB200_EMUL
DEVFN double dmul(double a, double b) { return a * b; }
DEVFN double dadd(double a, double b) { return a + b; }
DEVFN double dsub(double a, double b) { return a - b; }
DEVFN double ddiv(double a, double b) { return a / b; }
DEVFN float fsub(float a, float b) { return a - b; }
DEVFN float f_up(double d) { float f = (float)d; if ((double)f < d) f = std::nextafterf(f, INFINITY); return f; }
DEVFN float rcp_fast(float x) { return 1.0f / x; }
DEVFN float sqrt_fast(float x) { return std::sqrt(x); }
DEVFN float rsqrt_fast(float x) { return 1.0f / std::sqrt(x); }
#else
// 1-2 ulp approximations (single MUFU op); their error is covered by CP_KAPPA_W
DEVFN float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DEVFN float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DEVFN float rsqrt_fast(float x) { float r; asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DEVFN double dmul(double a, double b) { return __dmul_rn(a, b); }
DEVFN double dadd(double a, double b) { return __dadd_rn(a, b); }
DEVFN double dsub(double a, double b) { return __dsub_rn(a, b); }
DEVFN double ddiv(double a, double b) { return __ddiv_rn(a, b); }
DEVFN float fsub(float a, float b) { return __fsub_rn(a, b); }
DEVFN float f_up(double d) { return __double2float_ru(d); }
#endif

struct D3 { double x, y, z; };
DEVFN D3 d3sub(D3 a, D3 b) { return {dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)}; }
DEVFN double d3dot(D3 a, D3 b) { return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z)); }
DEVFN D3 d3madd(D3 a, D3 d, double s) { return {dadd(a.x, dmul(d.x, s)), dadd(a.y, dmul(d.y, s)), dadd(a.z, dmul(d.z, s))}; }

// Ericson, Real-Time Collision Detection ch.5, as used by
// point_simplex_squared_distance.cpp:45-110
DEVFN D3 closest_on_triangle(D3 p, D3 a, D3 b, D3 c) {
  const D3 ab = d3sub(b, a), ac = d3sub(c, a), ap = d3sub(p, a);
  const double d1 = d3dot(ab, ap), d2 = d3dot(ac, ap);
  if (d1 <= 0.0 && d2 <= 0.0) return a;
  const D3 bp = d3sub(p, b);
  const double d3 = d3dot(ab, bp), d4 = d3dot(ac, bp);
  if (d3 >= 0.0 && d4 <= d3) return b;
  const double vc = dsub(dmul(d1, d4), dmul(d3, d2));
  if (!(a.x == b.x && a.y == b.y && a.z == b.z)) {
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) return d3madd(a, ab, ddiv(d1, dsub(d1, d3)));
  }
  const D3 cp = d3sub(p, c);
  const double d5 = d3dot(ab, cp), d6 = d3dot(ac, cp);
  if (d6 >= 0.0 && d5 <= d6) return c;
  const double vb = dsub(dmul(d5, d2), dmul(d1, d6));
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) return d3madd(a, ac, ddiv(d2, dsub(d2, d6)));
  const double va = dsub(dmul(d3, d6), dmul(d5, d4));
  const double e43 = dsub(d4, d3), e56 = dsub(d5, d6);
  if (va <= 0.0 && e43 >= 0.0 && e56 >= 0.0) return d3madd(b, d3sub(c, b), ddiv(e43, dadd(e43, e56)));
  const double denom = ddiv(1.0, dadd(dadd(va, vb), vc));
  const double v = dmul(vb, denom), w = dmul(vc, denom);
  return d3madd(d3madd(a, ab, v), ac, w);
}

struct CpResult {
  double sqr_d;
  int32_t face;
  D3 c;
};

// ---------------------------------------------------------------------------
// Closest triangle, two stages.
//
// Stage 1 (fp32, inside the tree walk): a robust point-triangle distance with a
// two-sided error bound.  It is NOT Ericson's routine: the barycentric formula
// loses all accuracy on thin triangles in fp32, and a one-sided error would
// make pruning unsafe.  Instead
//     inside the prism over the face ?  |plane distance|
//                                    :  min over the three edges of the
//                                       point-segment distance,
// every term of which is a difference/dot of quantities of magnitude <= |p-v|,
// so |d32 - d| <= kappa * 2^-24 * max(|p|,|v|)  (kappa stored per face; faces too
// thin for a plane test use the boundary only and carry their height as extra
// error).  A face is a CANDIDATE while d32 - err <= (smallest d32 + err seen).
// Stage 2 (fp64, after the walk): the few surviving candidates — the winner and
// whatever ties with it inside the fp32 error band, e.g. the faces around a
// shared vertex — are evaluated with Ericson's routine in the reference's
// un-fused operation order; minimum squared distance wins, lowest face index
// on exact ties.  The result is therefore exactly what an all-fp64 search
// returns.
// ---------------------------------------------------------------------------
#ifndef CP_KAPPA_W
#define CP_KAPPA_W 24.0
#endif
#ifndef CP_KAPPA_IN
#define CP_KAPPA_IN 4.0
#endif
#ifndef CP_NCAND_N
#define CP_NCAND_N 3
#endif
static constexpr int CP_NCAND = CP_NCAND_N;

// Error bound |d32 - d| <= err, shape independent (no aspect-ratio amplification):
//   input rounding   (float)p and (float)origin move the configuration by <= 2*sqrt(3)*u, u = 2^-24*max(|p|,|mesh|)
//   everything else  is relative to w = p - origin: rounding of w (1.7), of the frame vectors (1.7), of the
//                    three dot products (5.2), of the 2-D edge distances (~6), approximate rcp/sqrt (~4)
//                    => < 19 * 2^-24 * |w|
// err = CP_KAPPA_W * 2^-24 * (|x|+|y|+|z|) + CP_KAPPA_IN * u   with (|x|+|y|+|z|) >= |w|.
DEVFN float tri_dist_f32(const float4 Q0, const float4 Q1, const float4 Q2, const float4 Q3, float px, float py, float pz,
                         float u, float &err) {
  const float wx = px - Q0.x, wy = py - Q0.y, wz = pz - Q0.z;
  const float x = wx * Q1.x + wy * Q1.y + wz * Q1.z;
  const float y = wx * Q2.x + wy * Q2.y + wz * Q2.z;
  const float z = wx * Q3.x + wy * Q3.y + wz * Q3.z;
  const float L = Q0.w, cx = Q1.w, cy = Q2.w;
  // edge (0,0)-(L,0)
  const float dx0 = x - fminf(fmaxf(x, 0.f), L);
  float d2 = dx0 * dx0 + y * y;
  // edge (0,0)-(cx,cy)
  float t = fminf(fmaxf((x * cx + y * cy) * rcp_fast(cx * cx + cy * cy), 0.f), 1.f);
  float qx = x - t * cx, qy = y - t * cy;
  d2 = fminf(d2, qx * qx + qy * qy);
  // edge (L,0)-(cx,cy)
  const float ex = cx - L, xl = x - L;
  t = fminf(fmaxf((xl * ex + y * cy) * rcp_fast(ex * ex + cy * cy), 0.f), 1.f);
  qx = xl - t * ex; qy = y - t * cy;
  d2 = fminf(d2, qx * qx + qy * qy);
  // inside the face (proper triangles only)
  const bool inside = (cy > 0.f) && (y >= 0.f) && (cy * x - cx * y >= 0.f) && (ex * y - cy * xl >= 0.f);
  d2 = inside ? 0.f : d2;
  err = (float)(CP_KAPPA_W * 5.9604644775390625e-08) * (fabsf(x) + fabsf(y) + fabsf(z)) + (float)CP_KAPPA_IN * u;
#ifdef CP_ERR_SCALE   // test hook: shrink the band to probe how much margin the bound has
  err *= (float)CP_ERR_SCALE;
#endif
  return sqrt_fast(d2 + z * z);
}

// running state of one query during the walk (registers)
struct CpLane {
  float px, py, pz;
  float slack3;   // rounding allowance of one fp32 bound test, folded into thr2
  float u;        // 2^-24 * max(|p|, |mesh|)
  float best_ub;  // fp32 upper bound of the distance to the nearest face seen so far
  float thr2;     // prune a box when its squared lower bound exceeds this
};
// exact best among the candidates already evaluated in fp64 (touched only by the rare fp64
// stage: lives in shared memory on the device so it does not cost registers in the walk)
struct CpExact {
  double d2;
  int32_t face;
  int32_t ncand;
  float min_lb;             // smallest lower bound among the listed candidates
  uint32_t cf[CP_NCAND];    // candidate faces of the fp32 stage (touched by the few lanes a face is close to)
  float cl[CP_NCAND];       // their lower bounds
};
// where the query itself can be re-read by the fp64 stage
struct CpQueryRef { const double *P; int64_t n, qi; double sigma; };
DEVFN D3 load_query(const CpQueryRef &q) { return D3{q.P[q.qi], q.P[q.n + q.qi], q.P[2 * q.n + q.qi]}; }

DEVFN void lane_set_ub(CpLane &L, float ub) {
  L.best_ub = ub;
  const float t = ub + L.slack3;
  L.thr2 = t * t * 1.000001f;  // rounded up; absorbs the rounding of a box test (slack3) and of its sum of squares
}

// `absmax` and the lane's whole fp32 state are in the mesh's SCALED coordinates (x sigma, a power of two: exact)
DEVFN void lane_init(CpLane &L, CpExact &X, D3 p, double up_sqr_d, float absmax, double sigma) {
  L.px = (float)(p.x * sigma); L.py = (float)(p.y * sigma); L.pz = (float)(p.z * sigma);
  const float M = fmaxf(fmaxf(fabsf(L.px), fabsf(L.py)), fmaxf(fabsf(L.pz), absmax));
  // rounding allowance of one bound test (pill_lb2: < 64 * 2^-24 M); twice as much is added to the pruning threshold
  // instead of widening the query
  L.slack3 = M * 7.6293945e-6f;   // 128 * 2^-24
  L.u = M * 5.9604645e-8f * 1.0000002f;
  // a query astronomically far from the mesh (its squares overflow fp32): the fp32 stage is blind, every face is a
  // candidate of the fp64 stage (brute force: what the exact answer takes when all faces tie to within 1e-16)
  if (!(M < 1.0e18f)) L.u = INFINITY;
  X.ncand = 0; X.min_lb = INFINITY;
  X.d2 = up_sqr_d; X.face = -1;
  const double ups = sqrt(up_sqr_d) * sigma;
  const bool bounded = ups < 1.0e18;   // its square must stay finite in fp32
  const float ub = bounded ? f_up(ups) * 1.0000002f : INFINITY;
  L.best_ub = ub;
  if (bounded) lane_set_ub(L, ub); else L.thr2 = INFINITY;
}

DEVFN double face_sqr_distance(const double *tri64, D3 p, uint32_t f, D3 &cc) {
  const double *t = tri64 + 9 * (size_t)f;
  const D3 a = {t[0], t[1], t[2]}, b = {t[3], t[4], t[5]}, c = {t[6], t[7], t[8]};
  cc = closest_on_triangle(p, a, b, c);
  const D3 df = d3sub(p, cc);
  return d3dot(df, df);
}

// fp64 evaluation of one face; strict improvement or exact tie with a lower index replaces the best
DEVFN void lane_exact(CpExact &X, const double *tri64, D3 p, uint32_t f) {
  D3 cc;
  const double dd = face_sqr_distance(tri64, p, f, cc);
  if (dd < X.d2 || (dd == X.d2 && X.face >= 0 && (int32_t)f < X.face)) { X.d2 = dd; X.face = (int32_t)f; }
}

// evaluate every live candidate in fp64 and empty the list (rare: kept out of line so the
// fp64 code does not sit in the walk's instruction footprint).  Takes and returns the upper bound BY VALUE: the lane
// state must never have its address taken, or it lives in local memory and every node visit reloads it.
#ifdef IGLB200_EMUL
inline
#else
static __device__ __noinline__
#endif
float lane_flush_ub(float best_ub, CpExact &X, const double *tri64, const double *P, int64_t n, int64_t qi, double sigma) {
  const CpQueryRef q{P, n, qi, sigma};
  const D3 p = load_query(q);
  const int nc = X.ncand;
  for (int k = 0; k < nc; ++k)
    if (!(X.cl[k] > best_ub)) lane_exact(X, tri64, p, X.cf[k]);
  X.ncand = 0; X.min_lb = INFINITY;
  if (X.face >= 0) {
    const float ub = f_up(sqrt(X.d2) * sigma) * 1.0000002f;   // the exact distance in the scaled units of the fp32 stage
    if (ub < best_ub) return ub;
  }
  return best_ub;
}
DEVFN void lane_flush(CpLane &L, CpExact &X, const double *tri64, const CpQueryRef &q) {
  const float ub = lane_flush_ub(L.best_ub, X, tri64, q.P, q.n, q.qi, q.sigma);
  if (ub < L.best_ub) lane_set_ub(L, ub);
}

// stage-1 result for face f (caller has already checked d32 - err <= best_ub)
DEVFN void lane_offer(CpLane &L, CpExact &X, const double *tri64, const CpQueryRef &q, uint32_t f, float d32, float err) {
  float lb = d32 - err;
  if (lb > L.best_ub) return;
  if (!(lb == lb)) lb = 0.f;   // a non-finite fp32 distance (a query astronomically far from the mesh) proves nothing: keep the face
  const float ub = d32 + err;
  int nc = X.ncand;
  float mlb = X.min_lb;
  if (ub < mlb) { nc = 0; mlb = INFINITY; }   // every listed candidate is now out of the band
  if (nc == CP_NCAND) {
    X.ncand = nc;
    lane_flush(L, X, tri64, q);
    if (lb > L.best_ub) return;
    nc = 0; mlb = INFINITY;
  }
  X.cf[nc] = f; X.cl[nc] = lb;
  X.ncand = nc + 1;
  X.min_lb = fminf(mlb, lb);
  if (ub < L.best_ub) lane_set_ub(L, ub);
}

DEVFN void lane_finish(CpLane &L, CpExact &X, const double *tri64, const CpQueryRef &q, double up_sqr_d, CpResult &res) {
  lane_flush(L, X, tri64, q);
  res.sqr_d = X.d2; res.face = X.face; res.c = {0.0, 0.0, 0.0};
  if (!(res.sqr_d < up_sqr_d) || res.face < 0) { res.sqr_d = up_sqr_d; res.face = -1; return; }
  face_sqr_distance(tri64, load_query(q), (uint32_t)res.face, res.c);   // closest point of the winner
}

// Squared lower bound of the distance from the query to the oriented cylinder of child lane i (PillNode, moments.h) in
// plain fp32.  With w = p - c:  z = w.n,  l = w - z n,  bound^2 = max(|z| - t, 0)^2 + max(|l| - R, 0)^2.
// Rounding: the query is rounded to fp32 (<= u per coordinate, u = 2^-24 M, M = max |coordinate| of query and mesh), the
// stored axis is unit to 2^-23, the three dot products / differences each lose <= 3 ulp of |w| <= 2 sqrt(3) M, the
// approximate square root 2 ulp of |l|: |z|, |l| and hence the bound are off by < 64 * 2^-24 M in total.  lane_init folds
// 128 * 2^-24 M into the pruning threshold, so a cylinder the exact distance would keep is never pruned.
DEVFN float pill_lb2(float cx, float cy, float cz, float nx, float ny, float nz, float R, float t, const CpLane &L) {
  const float wx = L.px - cx, wy = L.py - cy, wz = L.pz - cz;
  const float z = wx * nx + wy * ny + wz * nz;
  const float lx = wx - z * nx, ly = wy - z * ny, lz = wz - z * nz;
  const float rho = sqrt_fast(lx * lx + ly * ly + lz * lz);
  const float a = fmaxf(fabsf(z) - t, 0.f), b = fmaxf(rho - R, 0.f);
  return a * a + b * b;
}
DEVFN float box_lb2(float lx, float ly, float lz, float hx, float hy, float hz, const CpLane &L) {
  const float dx = fmaxf(fmaxf(lx - L.px, L.px - hx), 0.f);
  const float dy = fmaxf(fmaxf(ly - L.py, L.py - hy), 0.f);
  const float dz = fmaxf(fmaxf(lz - L.pz, L.pz - hz), 0.f);
  return dx * dx + dy * dy + dz * dz;
}
// kind: the node's box-lane mask (warp-uniform)
DEVFN float lane_pill_d2(const PillNode &n, int i, const CpLane &L, uint32_t kind) {
  if ((kind >> i) & 1u)
    return box_lb2(((const float *)&n.cx)[i], ((const float *)&n.cy)[i], ((const float *)&n.cz)[i],
                   ((const float *)&n.nx)[i], ((const float *)&n.ny)[i], ((const float *)&n.nz)[i], L);
  return pill_lb2(((const float *)&n.cx)[i], ((const float *)&n.cy)[i], ((const float *)&n.cz)[i],
                  ((const float *)&n.nx)[i], ((const float *)&n.ny)[i], ((const float *)&n.nz)[i],
                  ((const float *)&n.R)[i], ((const float *)&n.t)[i], L);
}

// The walk itself is a PACKET traversal: the 32 queries of a warp (neighbours in
// Morton order) share one stack and visit a node when ANY of them still needs it,
// so every box and face test runs on all 32 lanes at once (iglb200.cu:
// closest_packet_kernel; host emulation: closest_packet_emul).
// node ids per warp (3 per level at most: 42 levels).  128 entries keep a block at 2 KB of shared memory, so that 20
// blocks fit the 64 KB carve-out and L1 keeps 192 KB (192 entries -> 100 KB carve-out, 156 KB L1: +1.5 % time; the
// search is sensitive to L1: 124 KB +4.5 %, 92 KB +14 %).  Overflow raises an error, never a wrong answer.
static constexpr int CP_STACK = 128;

// ---------------------------------------------------------------------------
// UTsignedSolidAngleTri (FastWindingNumberForSoups.h:6070-6110), un-fused fp32
// ---------------------------------------------------------------------------
DEVFN float f3dot(const float *a, const float *b) {
  return plat::fadd(plat::fadd(plat::fmul(a[0], b[0]), plat::fmul(a[1], b[1])), plat::fmul(a[2], b[2]));
}
DEVFN float tri_solid_angle(const float4 *tri32, uint32_t f, float qx, float qy, float qz) {
  const float4 A = tri32[TRI32_STRIDE * (size_t)f], B = tri32[TRI32_STRIDE * (size_t)f + 1], C = tri32[TRI32_STRIDE * (size_t)f + 2];
  float qa[3] = {fsub(A.x, qx), fsub(A.y, qy), fsub(A.z, qz)};
  float qb[3] = {fsub(B.x, qx), fsub(B.y, qy), fsub(B.z, qz)};
  float qc[3] = {fsub(C.x, qx), fsub(C.y, qy), fsub(C.z, qz)};
  const float al = plat::fsqrt(f3dot(qa, qa)), bl = plat::fsqrt(f3dot(qb, qb)), cl = plat::fsqrt(f3dot(qc, qc));
  if (al == 0.f || bl == 0.f || cl == 0.f) return 0.f;
  const float ia = plat::fdiv(1.f, al), ib = plat::fdiv(1.f, bl), ic = plat::fdiv(1.f, cl);
  for (int d = 0; d < 3; ++d) { qa[d] = plat::fmul(qa[d], ia); qb[d] = plat::fmul(qb[d], ib); qc[d] = plat::fmul(qc[d], ic); }
  const float u[3] = {fsub(qb[0], qa[0]), fsub(qb[1], qa[1]), fsub(qb[2], qa[2])};
  const float v[3] = {fsub(qc[0], qa[0]), fsub(qc[1], qa[1]), fsub(qc[2], qa[2])};
  const float cr[3] = {fsub(plat::fmul(u[1], v[2]), plat::fmul(u[2], v[1])),
                       fsub(plat::fmul(u[2], v[0]), plat::fmul(u[0], v[2])),
                       fsub(plat::fmul(u[0], v[1]), plat::fmul(u[1], v[0]))};
  const float num = f3dot(qa, cr);
  if (num == 0.f) return 0.f;
  const float den = plat::fadd(plat::fadd(plat::fadd(1.f, f3dot(qa, qb)), f3dot(qa, qc)), f3dot(qb, qc));
  return 2.f * atan2f(num, den);
}

static constexpr int FWN_STACK = 96;

// Far-field order-<=2 expansion (:7200-7241) of all four child lanes of a node at once, the way
// the reference's SSE code does it: every coefficient row is one 16-byte load whose four
// components belong to the four children.  q (already relative to each child's centre) and
// ql2 = |q|^2 come from the caller's descend test.
struct Lanes4 { float v[4]; };
DEVFN Lanes4 row4(const float4 *p) { const float4 t = *p; return Lanes4{{t.x, t.y, t.z, t.w}}; }
#ifndef FWN_LD256
#define FWN_LD256 1   // winding kernel 15.9 -> 13.6 ms on c2
#endif
// two consecutive rows (32-byte aligned pair) by one 256-bit load on the device
DEVFN void rows2(const float4 *p, Lanes4 &a, Lanes4 &b) {
#if defined(__CUDA_ARCH__) && FWN_LD256
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.v[0]), "=f"(a.v[1]), "=f"(a.v[2]), "=f"(a.v[3]), "=f"(b.v[0]), "=f"(b.v[1]), "=f"(b.v[2]), "=f"(b.v[3])
               : "l"(p));
#else
  a = row4(p); b = row4(p + 1);
#endif
}

DEVFN void node_omega4(const FwnNode &n, const Lanes4 &NX, Lanes4 q0, Lanes4 q1, Lanes4 q2, const Lanes4 &ql2, int order, Lanes4 &omega) {
  Lanes4 m1, m2;
  // 1/|q| by one MUFU op (<= 2 ulp): the polynomial below is an approximation with ~1e-3 truncation
  // error anyway, and no branch depends on these values except the is-finite test
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m1.v[i] = rsqrt_fast(ql2.v[i]);
    m2.v[i] = m1.v[i] * m1.v[i];
    q0.v[i] *= m1.v[i]; q1.v[i] *= m1.v[i]; q2.v[i] *= m1.v[i];
  }
  {
    Lanes4 NY, NZ; rows2(&n.d[R_NY], NY, NZ);
#pragma unroll
    for (int i = 0; i < 4; ++i) omega.v[i] = -m2.v[i] * (q0.v[i] * NX.v[i] + q1.v[i] * NY.v[i] + q2.v[i] * NZ.v[i]);
  }
  if (order < 1) return;
  {
    Lanes4 NXX, NYY, NZZ, NXY, NYZ, NZX;
    rows2(&n.d[R_NXX], NXX, NYY); rows2(&n.d[R_NZZ], NZZ, NXY); rows2(&n.d[R_NYZ_NZY], NYZ, NZX);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a0 = q0.v[i], a1 = q1.v[i], a2 = q2.v[i];
      const float o1 = (m2.v[i] * m1.v[i]) * (NXX.v[i] + NYY.v[i] + NZZ.v[i] -
                                             3.0f * ((a0 * a0 * NXX.v[i] + a1 * a1 * NYY.v[i] + a2 * a2 * NZZ.v[i]) +
                                                     a0 * a1 * NXY.v[i] + a0 * a2 * NZX.v[i] + a1 * a2 * NYZ.v[i]));
      omega.v[i] += o1;
    }
  }
  if (order < 2) return;
  {
    Lanes4 NXXX, NYYY, NZZZ, SP, XXY, XXZ, YYZ, YYX, ZZX, ZZY;
    rows2(&n.d[R_NXXX], NXXX, NYYY); rows2(&n.d[R_NZZZ], NZZZ, SP); rows2(&n.d[R_2NXXY_NYXX], XXY, XXZ);
    rows2(&n.d[R_2NYYZ_NZYY], YYZ, YYX); rows2(&n.d[R_2NZZX_NXZZ], ZZX, ZZY);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a0 = q0.v[i], a1 = q1.v[i], a2 = q2.v[i];
      const float s0 = a0 * a0, s1 = a1 * a1, s2 = a2 * a2;
      const float t00 = YYX.v[i] + ZZX.v[i], t01 = ZZY.v[i] + XXY.v[i], t02 = XXZ.v[i] + YYZ.v[i];
      const float t10 = a1 * XXY.v[i] + a2 * XXZ.v[i];
      const float t11 = a2 * YYZ.v[i] + a0 * YYX.v[i];
      const float t12 = a0 * ZZX.v[i] + a1 * ZZY.v[i];
      const float o2 = (m2.v[i] * m2.v[i]) *
                       (1.5f * (a0 * (3.f * NXXX.v[i] + t00) + a1 * (3.f * NYYY.v[i] + t01) + a2 * (3.f * NZZZ.v[i] + t02)) -
                        7.5f * ((s0 * a0 * NXXX.v[i] + s1 * a1 * NYYY.v[i] + s2 * a2 * NZZZ.v[i]) + a0 * a1 * a2 * SP.v[i] +
                                (s0 * t10 + s1 * t11 + s2 * t12)));
      omega.v[i] += o2;
    }
  }
}

DEVFN bool finite_f(float x) { return (plat::f2u(x) & 0x7F800000u) != 0x7F800000u; }

// Solid angle (not yet divided by 4 pi) of the whole soup seen from q.
// `overflow` (nullable) is raised when the stack cannot hold a pushed child: the caller turns it into an error, the
// truncated sum is never returned as a result.
DEVFN float winding_solid_angle(const FwnNode *nodes, const float4 *tri32, float qx, float qy, float qz, float acc2,
                                int order, uint32_t *stack, uint32_t *visits /*nullable: [nodes, tris]*/, uint32_t *overflow = nullptr) {
  float sum = 0.f;
  int sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const uint32_t ni = stack[--sp];
    const FwnNode &n = nodes[ni];
    if (visits) visits[0]++;
    // rows pair up on 32-byte boundaries: (child, R2) (PX, PY) (PZ, NX) ...
    Lanes4 CH, R2, PX, PY, PZ, NX;
    rows2((const float4 *)&n.child, CH, R2); rows2(&n.d[R_PX], PX, PY); rows2(&n.d[R_PZ], PZ, NX);
    const uint32_t ch[4] = {plat::f2u(CH.v[0]), plat::f2u(CH.v[1]), plat::f2u(CH.v[2]), plat::f2u(CH.v[3])};
    uint32_t descend = 0;
    float node_sum = 0.f;
    Lanes4 q0, q1, q2, ql2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q0.v[i] = fsub(qx, PX.v[i]); q1.v[i] = fsub(qy, PY.v[i]); q2.v[i] = fsub(qz, PZ.v[i]);
      ql2.v[i] = plat::fadd(plat::fadd(plat::fmul(q0.v[i], q0.v[i]), plat::fmul(q1.v[i], q1.v[i])), plat::fmul(q2.v[i], q2.v[i]));
      if (ql2.v[i] <= plat::fmul(R2.v[i], acc2)) descend |= 1u << i;   // absent lanes have R2 = +inf
    }
    if (descend != 0xFu) {   // :7194-7198: nothing to approximate when all four descend
      Lanes4 om;
      node_omega4(n, NX, q0, q1, q2, ql2, order, om);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if ((descend >> i) & 1u) continue;
        if (!finite_f(om.v[i])) descend |= 1u << i;   // :7249-7251
        else node_sum += om.v[i];
      }
    }
    sum += node_sum;
    // children in lane order; push in reverse so that lane 0 pops first
    for (int i = 3; i >= 0; --i) {
      if (!((descend >> i) & 1u)) continue;
      const uint32_t c = ch[i];
      if (c == EMPTY) continue;
      if (c & INTERNAL_BIT) {
        if (sp < FWN_STACK) stack[sp++] = c & ~INTERNAL_BIT;
        else if (overflow) *overflow = 1u;   // plain store of the same value by every thread that overflows
      }
      else { sum += tri_solid_angle(tri32, c, qx, qy, qz); if (visits) visits[1]++; }
    }
  }
  return sum;
}

#ifndef IGLB200_EMUL
// The same evaluation for a PACKET: the 32 queries of a warp (neighbours in Morton order) walk the union of their
// nodes with one shared stack of (node, lanes that descend into it).  Every node row is then one warp-uniform load —
// the per-thread walk is bound by the L1 data pipe (each thread fetching its own copy of a 384-byte node: 92 % of the
// LSU wavefront peak on c2) — and no thread keeps a stack in local memory.  Per query the arithmetic and its ORDER are
// exactly those of winding_solid_angle (the far-field sum of a node, then its leaf triangles from lane 3 down to 0,
// children visited from lane 0 up), so the result is bit-identical to the per-thread walk.
static constexpr int FWN_PSTACK = 112;   // (node, mask) entries per warp: 3 pushes per level
DEVFN float winding_solid_angle_packet(const FwnNode *nodes, const float4 *tri32, bool alive, float qx, float qy, float qz, float acc2,
                                       int order, uint2 *stack, uint32_t *overflow) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  float sum = 0.f;
  int sp = 0;
  const unsigned m0 = __ballot_sync(FULL, alive);
  if (m0 == 0u) return 0.f;
  if (lane == 0) stack[0] = make_uint2(0u, m0);
  sp = 1;
  __syncwarp();
  while (sp > 0) {
    const uint2 e = stack[--sp];
    __syncwarp();
    const bool me = (e.y >> lane) & 1u;
    const FwnNode &n = nodes[e.x];
    Lanes4 CH, R2, PX, PY, PZ, NX;
    rows2((const float4 *)&n.child, CH, R2); rows2(&n.d[R_PX], PX, PY); rows2(&n.d[R_PZ], PZ, NX);
    const uint32_t ch[4] = {plat::f2u(CH.v[0]), plat::f2u(CH.v[1]), plat::f2u(CH.v[2]), plat::f2u(CH.v[3])};
    uint32_t descend = 0;
    Lanes4 q0, q1, q2, ql2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q0.v[i] = fsub(qx, PX.v[i]); q1.v[i] = fsub(qy, PY.v[i]); q2.v[i] = fsub(qz, PZ.v[i]);
      ql2.v[i] = plat::fadd(plat::fadd(plat::fmul(q0.v[i], q0.v[i]), plat::fmul(q1.v[i], q1.v[i])), plat::fmul(q2.v[i], q2.v[i]));
      if (ql2.v[i] <= plat::fmul(R2.v[i], acc2)) descend |= 1u << i;   // absent lanes have R2 = +inf
    }
    if (!me) descend = 0u;
    if (__any_sync(FULL, me && descend != 0xFu)) {   // :7194-7198: nothing to approximate when all four descend
      Lanes4 om;
      node_omega4(n, NX, q0, q1, q2, ql2, order, om);
      if (me && descend != 0xFu) {
        float node_sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((descend >> i) & 1u) continue;
          if (!finite_f(om.v[i])) descend |= 1u << i;   // :7249-7251
          else node_sum += om.v[i];
        }
        sum += node_sum;
      }
    }
    // children in lane order; push in reverse so that lane 0 pops first
#pragma unroll
    for (int i = 3; i >= 0; --i) {
      const unsigned dm = __ballot_sync(FULL, (descend >> i) & 1u);
      if (dm == 0u) continue;
      const uint32_t c = ch[i];
      if (c == EMPTY) continue;
      if (c & INTERNAL_BIT) {
        if (sp < FWN_PSTACK) { if (lane == 0) stack[sp] = make_uint2(c & ~INTERNAL_BIT, dm); ++sp; }
        else if (lane == 0) *overflow = 1u;
      } else if ((dm >> lane) & 1u) {
        sum += tri_solid_angle(tri32, c, qx, qy, qz);
      }
    }
    __syncwarp();
  }
  return sum;
}
#endif

}  // namespace fwn
