// igl::winding_number(V,F,O,W) for triangle meshes, computed on the GPU.
//   /root/reference/include/igl/winding_number.h:35-43 (batch), :53-60 (single point);
//   semantics winding_number.cpp:17-89 + solid_angle.cpp:10-55: the exact generalised winding number in the
//   matrices' own precision — here always evaluated in double (csrc/wn.h) and cast to W's scalar.
// 2-D inputs (F.cols() == 2, signed_angle) are not covered.
#pragma once
#include "fast_winding_number.h"

namespace iglb200 {
template <class DerivedV, class DerivedF, class DerivedO, class DerivedW>
inline void winding_number(const DerivedV &V, const DerivedF &F, const DerivedO &O, DerivedW &W) {
  if (F.rows() && F.cols() != 3) throw Error(IGLB200_ERR_UNSUPPORTED, "iglb200: winding_number needs triangles (F with 3 columns)");
  W.resize(O.rows(), 1);
  if (F.rows() == 0) { for (int64_t i = 0; i < (int64_t)O.rows(); ++i) W(i, 0) = 0; return; }
  iglb200_mesh *m = detail::make_mesh(V, F, 2, IGLB200_MESH_WINDING_NUMBER);
  auto o = detail::view(O);
  const int rc = iglb200_winding_number(m, o.ptr, o.is_f64, o.rows, o.rs, o.cs, (void *)W.data(),
                                        std::is_same<typename DerivedW::Scalar, double>::value ? 1 : 0);
  iglb200_mesh_destroy(m);
  check(rc);
}
}  // namespace iglb200
