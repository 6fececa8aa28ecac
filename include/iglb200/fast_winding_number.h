// igl::fast_winding_number with the reference's signatures, evaluated on the GPU.
//   soup:   /root/reference/include/igl/fast_winding_number.h:160-206 (FastWindingNumberBVH :140-147)
//   cloud:  /root/reference/include/igl/fast_winding_number.h:44-53, 88-131
#pragma once
#include "common.h"

namespace iglb200 {

// Replaces igl::FastWindingNumberBVH.  Owns a device mesh handle; V and F are copied at build time
// (the reference keeps raw pointers into U/F, FastWindingNumberForSoups.h:6232-6234).
struct FastWindingNumberBVH {
  iglb200_mesh *mesh = nullptr;
  FastWindingNumberBVH() = default;
  FastWindingNumberBVH(const FastWindingNumberBVH &) = delete;
  FastWindingNumberBVH &operator=(const FastWindingNumberBVH &) = delete;
  ~FastWindingNumberBVH() { clear(); }
  void clear() { if (mesh) iglb200_mesh_destroy(mesh); mesh = nullptr; }
};

namespace detail {
template <class DerivedV, class DerivedF>
inline iglb200_mesh *make_mesh(const DerivedV &V, const DerivedF &F, int order, int flags = 0) {
  static_assert(std::is_same<typename DerivedF::Scalar, int>::value || std::is_same<typename DerivedF::Scalar, int32_t>::value,
                "iglb200: F must hold 32-bit ints");
  if (V.cols() != 3 || (F.rows() && (F.cols() < 1 || F.cols() > 3))) throw Error(IGLB200_ERR_ARG, "iglb200: V must be #Vx3 and F #Fx3 (or #Fx2/#Fx1)");
  auto v = view(V);
  iglb200_mesh *m = nullptr;
  if (F.rows() && F.cols() != 3) {
    // segments / points: corner k is Ele(i, k % cols) (point_simplex_squared_distance.cpp:126-130)
    const int64_t n = (int64_t)F.rows(), nc = (int64_t)F.cols();
    int32_t *t = new int32_t[(size_t)3 * n];
    for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) t[3 * i + k] = (int32_t)F(i, k % nc);
    const int rc = iglb200_mesh_create_ex(v.ptr, v.is_f64, v.rows, v.rs, v.cs, t, n, 3, 1, order, flags, &m);
    delete[] t;
    check(rc);
    return m;
  }
  check(iglb200_mesh_create_ex(v.ptr, v.is_f64, v.rows, v.rs, v.cs, (const int32_t *)F.data(), (int64_t)F.rows(),
                               F.rows() ? row_stride(F) : 3, F.rows() ? col_stride(F) : 1, order, flags, &m));
  return m;
}
}  // namespace detail

// fast_winding_number(V,F,order,fwn_bvh)  — precompute (fast_winding_number.cpp:383-416)
template <class DerivedV, class DerivedF>
inline void fast_winding_number(const DerivedV &V, const DerivedF &F, const int order, FastWindingNumberBVH &fwn_bvh) {
  fwn_bvh.clear();
  fwn_bvh.mesh = detail::make_mesh(V, F, order);
}

// fast_winding_number(fwn_bvh,accuracy_scale,Q,W)  (fast_winding_number.cpp:421-437)
template <class DerivedQ, class DerivedW>
inline void fast_winding_number(const FastWindingNumberBVH &fwn_bvh, const float accuracy_scale, const DerivedQ &Q, DerivedW &W) {
  if (!fwn_bvh.mesh) throw Error(IGLB200_ERR_ARG, "iglb200: FastWindingNumberBVH is empty");
  W.resize(Q.rows(), 1);
  auto q = detail::view(Q);
  check(iglb200_fwn_eval(fwn_bvh.mesh, q.ptr, q.is_f64, q.rows, q.rs, q.cs, accuracy_scale, (void *)W.data(),
                         std::is_same<typename DerivedW::Scalar, double>::value ? 1 : 0));
}

// fast_winding_number(fwn_bvh,accuracy_scale,p) -> w, one query (fast_winding_number.h:196-206, fast_winding_number.cpp:439-455).
// Functional but a whole device round trip per point: batch the queries wherever possible.  p is a 1x3 (or 3x1) vector.
template <class Derivedp>
inline typename Derivedp::Scalar fast_winding_number(const FastWindingNumberBVH &fwn_bvh, const float accuracy_scale, const Derivedp &p) {
  if (!fwn_bvh.mesh) throw Error(IGLB200_ERR_ARG, "iglb200: FastWindingNumberBVH is empty");
  if ((int64_t)p.rows() * (int64_t)p.cols() != 3) throw Error(IGLB200_ERR_ARG, "iglb200: p must have 3 entries");
  const double q[3] = {(double)p(0), (double)p(1), (double)p(2)};
  double w = 0.0;
  check(iglb200_fwn_eval(fwn_bvh.mesh, q, 1, 1, 3, 1, accuracy_scale, &w, 1));
  return (typename Derivedp::Scalar)w;
}

// fast_winding_number(V,F,Q,W)  (fast_winding_number.cpp:367-378: order 2, accuracy_scale 2)
template <class DerivedV, class DerivedF, class DerivedQ, class DerivedW>
inline typename std::enable_if<!std::is_arithmetic<DerivedF>::value && std::is_integral<typename DerivedF::Scalar>::value>::type
fast_winding_number(const DerivedV &V, const DerivedF &F, const DerivedQ &Q, DerivedW &W) {
  FastWindingNumberBVH bvh;
  fast_winding_number(V, F, 2, bvh);
  fast_winding_number(bvh, 2.0f, Q, W);
}

// fast_winding_number(P,N,A,Q,expansion_order,beta,WN)  — oriented point cloud
// (fast_winding_number.cpp:320-344: octree + precompute + evaluation in one call)
template <class DerivedP, class DerivedN, class DerivedA, class DerivedQ, class BetaType, class DerivedWN>
inline void fast_winding_number(const DerivedP &P, const DerivedN &N, const DerivedA &A, const DerivedQ &Q,
                                const int expansion_order, const BetaType beta, DerivedWN &WN) {
  using S = typename DerivedP::Scalar;
  static_assert(std::is_same<S, typename DerivedN::Scalar>::value && std::is_same<S, typename DerivedA::Scalar>::value,
                "iglb200: P, N and A must share a scalar type");
  if (P.cols() != 3 || N.cols() != 3 || N.rows() != P.rows() || A.rows() * A.cols() != P.rows())
    throw Error(IGLB200_ERR_ARG, "iglb200: P,N must be #Px3 and A #P");
  // the C ABI takes P, N row-major and A contiguous: repack (3 small host copies; the reference
  // makes per-cell index lists of the same total size)
  const int64_t n = (int64_t)P.rows();
  S *buf = new S[(size_t)7 * (n ? n : 1)];
  S *p = buf, *nn = buf + 3 * n, *a = buf + 6 * n;
  for (int64_t i = 0; i < n; ++i) {
    for (int d = 0; d < 3; ++d) { p[3 * i + d] = P(i, d); nn[3 * i + d] = N(i, d); }
    a[i] = A.data()[i * (A.cols() == 1 ? detail::row_stride(A) : detail::col_stride(A))];
  }
  iglb200_cloud *c = nullptr;
  int rc = iglb200_cloud_create(p, nn, a, std::is_same<S, double>::value ? 1 : 0, n, expansion_order, &c);
  delete[] buf;
  check(rc);
  WN.resize(Q.rows(), 1);
  auto q = detail::view(Q);
  rc = iglb200_cloud_eval(c, q.ptr, q.is_f64, q.rows, q.rs, q.cs, (double)beta, (void *)WN.data(),
                          std::is_same<typename DerivedWN::Scalar, double>::value ? 1 : 0);
  iglb200_cloud_destroy(c);
  check(rc);
}

// The reference's cached point-cloud overloads split the call above in two, handing the octree (point_indices, CH) and
// the expansion (CM, R, EC) back to the caller so that one precompute serves many evaluations:
//   fast_winding_number(P,N,A,point_indices,CH,expansion_order,CM,R,EC)     fast_winding_number.h:44-53
//   fast_winding_number(P,N,A,point_indices,CH,CM,R,EC,Q,beta,WN)           fast_winding_number.h:88-100
// Here the whole of it — octree, CM, R, EC — stays on the device behind ONE opaque handle:
struct FastWindingNumberCloud {
  iglb200_cloud *cloud = nullptr;
  FastWindingNumberCloud() = default;
  FastWindingNumberCloud(const FastWindingNumberCloud &) = delete;
  FastWindingNumberCloud &operator=(const FastWindingNumberCloud &) = delete;
  ~FastWindingNumberCloud() { clear(); }
  void clear() { if (cloud) iglb200_cloud_destroy(cloud); cloud = nullptr; }
};
// precompute: igl::octree(P,...) + fast_winding_number(P,N,A,point_indices,CH,expansion_order,CM,R,EC)
template <class DerivedP, class DerivedN, class DerivedA>
inline void fast_winding_number(const DerivedP &P, const DerivedN &N, const DerivedA &A, const int expansion_order,
                                FastWindingNumberCloud &fwn_cloud) {
  using S = typename DerivedP::Scalar;
  static_assert(std::is_same<S, typename DerivedN::Scalar>::value && std::is_same<S, typename DerivedA::Scalar>::value,
                "iglb200: P, N and A must share a scalar type");
  if (P.cols() != 3 || N.cols() != 3 || N.rows() != P.rows() || A.rows() * A.cols() != P.rows())
    throw Error(IGLB200_ERR_ARG, "iglb200: P,N must be #Px3 and A #P");
  fwn_cloud.clear();
  const int64_t n = (int64_t)P.rows();
  S *buf = new S[(size_t)7 * (n ? n : 1)];
  S *p = buf, *nn = buf + 3 * n, *a = buf + 6 * n;
  for (int64_t i = 0; i < n; ++i) {
    for (int d = 0; d < 3; ++d) { p[3 * i + d] = P(i, d); nn[3 * i + d] = N(i, d); }
    a[i] = A.data()[i * (A.cols() == 1 ? detail::row_stride(A) : detail::col_stride(A))];
  }
  const int rc = iglb200_cloud_create(p, nn, a, std::is_same<S, double>::value ? 1 : 0, n, expansion_order, &fwn_cloud.cloud);
  delete[] buf;
  check(rc);
}
// evaluate: fast_winding_number(P,N,A,point_indices,CH,CM,R,EC,Q,beta,WN)
template <class DerivedQ, class BetaType, class DerivedWN>
inline void fast_winding_number(const FastWindingNumberCloud &fwn_cloud, const DerivedQ &Q, const BetaType beta, DerivedWN &WN) {
  if (!fwn_cloud.cloud) throw Error(IGLB200_ERR_ARG, "iglb200: FastWindingNumberCloud is empty");
  WN.resize(Q.rows(), 1);
  auto q = detail::view(Q);
  check(iglb200_cloud_eval(fwn_cloud.cloud, q.ptr, q.is_f64, q.rows, q.rs, q.cs, (double)beta, (void *)WN.data(),
                           std::is_same<typename DerivedWN::Scalar, double>::value ? 1 : 0));
}

// fast_winding_number(P,N,A,Q,WN)  (fast_winding_number.cpp:346-365: order 2, beta 2)
template <class DerivedP, class DerivedN, class DerivedA, class DerivedQ, class DerivedWN>
inline typename std::enable_if<std::is_floating_point<typename DerivedA::Scalar>::value && !std::is_arithmetic<DerivedQ>::value>::type
fast_winding_number(const DerivedP &P, const DerivedN &N, const DerivedA &A, const DerivedQ &Q, DerivedWN &WN) {
  fast_winding_number(P, N, A, Q, 2, 2.0, WN);
}

}  // namespace iglb200
