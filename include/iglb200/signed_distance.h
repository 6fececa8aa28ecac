// igl::signed_distance with the reference's Eigen signatures, computed on the GPU.
//   /root/reference/include/igl/signed_distance.h:20-34 (enum), :61-89 and :91-110 (signed_distance),
//   :283-312 (signed_distance_fast_winding_number); semantics signed_distance.cpp:28-139, 372-440.
// All five sign types run on the device (there is no CPU fallback): FAST_WINDING_NUMBER, UNSIGNED, PSEUDONORMAL,
// and WINDING_NUMBER = DEFAULT (exact hierarchical winding number, csrc/wn.h).
#pragma once
#include "AABB.h"

namespace iglb200 {

enum SignedDistanceType {
  SIGNED_DISTANCE_TYPE_PSEUDONORMAL = 0,
  SIGNED_DISTANCE_TYPE_WINDING_NUMBER = 1,
  SIGNED_DISTANCE_TYPE_DEFAULT = 2,
  SIGNED_DISTANCE_TYPE_UNSIGNED = 3,
  SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER = 4,
  NUM_SIGNED_DISTANCE_TYPE = 5
};

namespace detail {
// throw-away I / C outputs for the S-only overload
struct ScratchI {
  using Scalar = int32_t;
  int32_t *p = nullptr;
  ~ScratchI() { delete[] p; }
  void resize(int64_t r, int64_t) { delete[] p; p = new int32_t[r ? r : 1]; }
  int32_t *data() { return p; }
};
struct ScratchC {
  using Scalar = double;
  static constexpr bool IsRowMajor = false;
  double *p = nullptr; int64_t n = 0;
  ~ScratchC() { delete[] p; }
  void resize(int64_t r, int64_t c) { delete[] p; p = new double[r * c ? r * c : 1]; n = r; }
  double *data() { return p; }
  const double *data() const { return p; }
  int64_t rows() const { return n; }
  int64_t cols() const { return 3; }
  int64_t innerStride() const { return 1; }
  int64_t outerStride() const { return n; }
};
template <class DerivedP, class DerivedS, class DerivedI, class DerivedC, class DerivedW, class DerivedN = ScratchC>
inline void sd_query(const iglb200_mesh *mesh, const DerivedP &P, int sign_type, double lower, double upper, DerivedS &S, DerivedI &I,
                     DerivedC &C, DerivedW *W, DerivedN *N = nullptr) {
  static_assert(sizeof(typename DerivedI::Scalar) == 4, "iglb200: I must hold 32-bit ints");
  S.resize(P.rows(), 1);
  I.resize(P.rows(), 1);
  C.resize(P.rows(), 3);
  if (W) W->resize(P.rows(), 1);
  if (N) N->resize(P.rows(), 3);
  auto p = view(P);
  auto c = view(C);
  int64_t nrs = 1, ncs = 1; int nf64 = 1; void *nptr = nullptr;
  if (N) { auto nv = view(*N); nrs = nv.rs; ncs = nv.cs; nf64 = nv.is_f64; nptr = (void *)N->data(); }
  check(iglb200_signed_distance(mesh, p.ptr, p.is_f64, p.rows, p.rs, p.cs, sign_type, lower, upper, (void *)S.data(),
                                std::is_same<typename DerivedS::Scalar, double>::value ? 1 : 0, (int32_t *)I.data(), (void *)C.data(),
                                c.is_f64, c.rs, c.cs, W ? (void *)W->data() : nullptr,
                                W ? (std::is_same<typename DerivedW::Scalar, double>::value ? 1 : 0) : 1, nptr, nf64, nrs, ncs));
}
}  // namespace detail

// signed_distance(P,V,F,sign_type,lower_bound,upper_bound,S,I,C,N).  N is written only for
// PSEUDONORMAL and left untouched otherwise, exactly as the reference does (signed_distance.cpp:75).
template <class DerivedP, class DerivedV, class DerivedF, class DerivedS, class DerivedI, class DerivedC, class DerivedN>
inline void signed_distance(const DerivedP &P, const DerivedV &V, const DerivedF &F, const SignedDistanceType sign_type,
                            const typename DerivedV::Scalar lower_bound, const typename DerivedV::Scalar upper_bound, DerivedS &S,
                            DerivedI &I, DerivedC &C, DerivedN &N) {
  if (F.rows() == 0) {  // signed_distance.cpp:398-405
    S.resize(P.rows(), 1); I.resize(P.rows(), 1); C.resize(P.rows(), 3); N.resize(P.rows(), 3);   // N too: :403
    for (int64_t i = 0; i < (int64_t)P.rows(); ++i) {
      S(i, 0) = std::numeric_limits<typename DerivedS::Scalar>::quiet_NaN();
      I(i, 0) = -1;
      for (int d = 0; d < 3; ++d) C(i, d) = std::numeric_limits<typename DerivedC::Scalar>::quiet_NaN();
      for (int d = 0; d < 3; ++d) N(i, d) = std::numeric_limits<typename DerivedN::Scalar>::quiet_NaN();
    }
    return;
  }
  const bool pseudo = sign_type == SIGNED_DISTANCE_TYPE_PSEUDONORMAL;
  const bool exact = sign_type == SIGNED_DISTANCE_TYPE_WINDING_NUMBER || sign_type == SIGNED_DISTANCE_TYPE_DEFAULT;
  iglb200_mesh *m = detail::make_mesh(V, F, 2, (pseudo ? IGLB200_MESH_PSEUDONORMALS : 0) | (exact ? IGLB200_MESH_WINDING_NUMBER : 0));
  try {
    detail::sd_query<DerivedP, DerivedS, DerivedI, DerivedC, DerivedS, DerivedN>(m, P, (int)sign_type, (double)lower_bound,
                                                                               (double)upper_bound, S, I, C, nullptr, pseudo ? &N : nullptr);
  } catch (...) { iglb200_mesh_destroy(m); throw; }
  iglb200_mesh_destroy(m);
}

// signed_distance(P,V,F,sign_type,S,I,C,N): default bounds (signed_distance.cpp:426-440)
template <class DerivedP, class DerivedV, class DerivedF, class DerivedS, class DerivedI, class DerivedC, class DerivedN>
inline void signed_distance(const DerivedP &P, const DerivedV &V, const DerivedF &F, const SignedDistanceType sign_type, DerivedS &S,
                            DerivedI &I, DerivedC &C, DerivedN &N) {
  using Scalar = typename DerivedV::Scalar;
  signed_distance(P, V, F, sign_type, std::numeric_limits<Scalar>::min(), std::numeric_limits<Scalar>::max(), S, I, C, N);
}

// extension: also returns the winding number W the sign was made from
template <class DerivedP, class DerivedV, class DerivedF, class DerivedS, class DerivedI, class DerivedC, class DerivedW>
inline void signed_distance_with_winding_number(const DerivedP &P, const DerivedV &V, const DerivedF &F, DerivedS &S, DerivedI &I,
                                                DerivedC &C, DerivedW &W) {
  using Scalar = typename DerivedV::Scalar;
  iglb200_mesh *m = detail::make_mesh(V, F, 2);
  try {
    detail::sd_query(m, P, (int)SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, (double)std::numeric_limits<Scalar>::min(),
                     (double)std::numeric_limits<Scalar>::max(), S, I, C, &W);
  } catch (...) { iglb200_mesh_destroy(m); throw; }
  iglb200_mesh_destroy(m);
}

// signed_distance_fast_winding_number(P,V,F,tree,fwn_bvh,S): prebuilt structures, S only
// (signed_distance.cpp:665-687).  Either handle holds both trees; fwn_bvh's is used.
template <class DerivedP, class DerivedV, class DerivedF, class DerivedS>
inline void signed_distance_fast_winding_number(const DerivedP &P, const DerivedV & /*V*/, const DerivedF & /*F*/,
                                                const AABB<DerivedV, 3> &tree, const FastWindingNumberBVH &fwn_bvh, DerivedS &S) {
  const iglb200_mesh *m = fwn_bvh.mesh ? fwn_bvh.mesh : tree.mesh;
  if (!m) throw Error(IGLB200_ERR_ARG, "iglb200: neither tree nor fwn_bvh has been built");
  using Scalar = typename DerivedV::Scalar;
  detail::ScratchI I; detail::ScratchC C;
  detail::sd_query<DerivedP, DerivedS, detail::ScratchI, detail::ScratchC, DerivedS>(m, P, (int)SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER,
                                                         (double)std::numeric_limits<Scalar>::min(),
                                                         (double)std::numeric_limits<Scalar>::max(), S, I, C, nullptr);
}

// signed_distance_fast_winding_number(q,V,F,tree,fwn_bvh) -> s, one query (signed_distance.h:300-312,
// signed_distance.cpp:689-713).  Functional but a whole device round trip per point: batch wherever possible.
template <class Derivedq, class DerivedV, class DerivedF>
inline typename DerivedV::Scalar signed_distance_fast_winding_number(const Derivedq &q, const DerivedV & /*V*/, const DerivedF & /*F*/,
                                                                     const AABB<DerivedV, 3> &tree, const FastWindingNumberBVH &fwn_bvh) {
  const iglb200_mesh *m = fwn_bvh.mesh ? fwn_bvh.mesh : tree.mesh;
  if (!m) throw Error(IGLB200_ERR_ARG, "iglb200: neither tree nor fwn_bvh has been built");
  if ((int64_t)q.rows() * (int64_t)q.cols() != 3) throw Error(IGLB200_ERR_ARG, "iglb200: q must have 3 entries");
  const double p[3] = {(double)q(0), (double)q(1), (double)q(2)};
  double s = 0.0, c[3];
  int32_t i = -1;
  // unbounded, as the reference's single-query form (tree.squared_distance(V,F,q,i,c))
  check(iglb200_signed_distance(m, p, 1, 1, 3, 1, (int)SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, 0.0, std::numeric_limits<double>::max(),
                                &s, 1, &i, c, 1, 3, 1, nullptr, 1, nullptr, 1, 3, 1));
  return (typename DerivedV::Scalar)s;
}

// signed_distance_pseudonormal(P,V,F,tree,FN,VN,EN,EMAP,S,I,C,N): batch overload with prebuilt structures
// (signed_distance.h:135-147, signed_distance.cpp:468-512).  The device keeps its own copy of the angle-weighted
// vertex, uniform edge and face normals, computed from V and F exactly as per_face_normals / per_vertex_normals /
// per_edge_normals do (face and edge normals bit-identical, vertex normals to one ulp), so FN, VN, EN and EMAP are
// accepted for signature compatibility and not read: callers who pass normals of their own making get the standard
// ones.
template <class DerivedP, class DerivedV, class DerivedF, class DerivedFN, class DerivedVN, class DerivedEN, class DerivedEMAP,
          class DerivedS, class DerivedI, class DerivedC, class DerivedN>
inline void signed_distance_pseudonormal(const DerivedP &P, const DerivedV &V, const DerivedF &F, const AABB<DerivedV, 3> & /*tree*/,
                                         const DerivedFN & /*FN*/, const DerivedVN & /*VN*/, const DerivedEN & /*EN*/,
                                         const DerivedEMAP & /*EMAP*/, DerivedS &S, DerivedI &I, DerivedC &C, DerivedN &N) {
  using Scalar = typename DerivedV::Scalar;
  // the reference searches without bounds here (:500-511)
  signed_distance(P, V, F, SIGNED_DISTANCE_TYPE_PSEUDONORMAL, std::numeric_limits<Scalar>::min(), std::numeric_limits<Scalar>::max(), S, I,
                  C, N);
}

}  // namespace iglb200
