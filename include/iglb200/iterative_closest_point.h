// igl::iterative_closest_point with the reference's signatures, on the GPU.
//   /root/reference/include/igl/iterative_closest_point.h:36-53 (build the tree, then iterate), :55-82 (prebuilt tree + NY)
//   semantics: iterative_closest_point.cpp:40-130 (sampling, closest points, rigid_alignment, R/t update)
#pragma once
#include "AABB.h"

namespace iglb200 {

namespace detail {
template <class DerivedVX, class DerivedFX, class DerivedR, class Derivedt>
inline void icp_run(const iglb200_mesh *meshY, const DerivedVX &VX, const DerivedFX &FX, int num_samples, int max_iters, DerivedR &R, Derivedt &t) {
  static_assert(sizeof(typename DerivedFX::Scalar) == 4, "iglb200: FX must hold 32-bit ints");
  if (VX.cols() != 3 || FX.cols() != 3) throw Error(IGLB200_ERR_ARG, "iglb200: VX must be #VXx3 and FX #FXx3");
  auto v = view(VX);
  double Rr[9], tt[3];
  // the reference's generate_default_urbg() is std::rand()-seeded: draw the seed the same way
  const uint64_t seed = (uint64_t)std::rand();
  check(iglb200_iterative_closest_point(meshY, v.ptr, v.is_f64, v.rows, v.rs, v.cs, (const int32_t *)FX.data(), (int64_t)FX.rows(),
                                        row_stride(FX), col_stride(FX), num_samples, max_iters, seed, Rr, tt));
  R.resize(3, 3);
  t.resize(1, 3);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R(i, j) = (typename DerivedR::Scalar)Rr[3 * i + j]; t(0, i) = (typename Derivedt::Scalar)tt[i]; }
}
}  // namespace detail

// iterative_closest_point(VX,FX,VY,FY,num_samples,max_iters,R,t)
template <class DerivedVX, class DerivedFX, class DerivedVY, class DerivedFY, class DerivedR, class Derivedt>
inline void iterative_closest_point(const DerivedVX &VX, const DerivedFX &FX, const DerivedVY &VY, const DerivedFY &FY, const int num_samples,
                                    const int max_iters, DerivedR &R, Derivedt &t) {
  AABB<DerivedVY, 3> Ytree;
  Ytree.init(VY, FY);
  detail::icp_run(Ytree.mesh, VX, FX, num_samples, max_iters, R, t);
}

// iterative_closest_point(VX,FX,VY,FY,Ytree,NY,num_samples,max_iters,R,t): prebuilt tree.  NY (per_face_normals(VY,FY)) is
// accepted for signature compatibility; the device computes the same normals from its own copy of Y.
template <class DerivedVX, class DerivedFX, class DerivedVY, class DerivedFY, class DerivedNY, class DerivedR, class Derivedt>
inline void iterative_closest_point(const DerivedVX &VX, const DerivedFX &FX, const DerivedVY & /*VY*/, const DerivedFY & /*FY*/,
                                    const AABB<DerivedVY, 3> &Ytree, const DerivedNY & /*NY*/, const int num_samples, const int max_iters,
                                    DerivedR &R, Derivedt &t) {
  if (!Ytree.mesh) throw Error(IGLB200_ERR_ARG, "iglb200::AABB: init() was not called");
  detail::icp_run(Ytree.mesh, VX, FX, num_samples, max_iters, R, t);
}

}  // namespace iglb200
