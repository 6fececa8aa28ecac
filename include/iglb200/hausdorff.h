// igl::hausdorff(VA,FA,VB,FB,d)   /root/reference/include/igl/hausdorff.cpp:20-39
// (a caller of the batch closest-point path: SURVEY 8f rank 3)
#pragma once
#include <cmath>
#include "point_mesh_squared_distance.h"
#include "signed_distance.h"

namespace iglb200 {
namespace detail {
struct ScratchD {
  using Scalar = double;
  double *p = nullptr; int64_t n = 0;
  ~ScratchD() { delete[] p; }
  void resize(int64_t r, int64_t) { delete[] p; p = new double[r ? r : 1]; n = r; }
  double *data() { return p; }
};
}  // namespace detail

template <class DerivedVA, class DerivedFA, class DerivedVB, class DerivedFB, class ScalarD>
inline void hausdorff(const DerivedVA &VA, const DerivedFA &FA, const DerivedVB &VB, const DerivedFB &FB, ScalarD &d) {
  detail::ScratchD dBA, dAB;
  detail::ScratchI I; detail::ScratchC C;
  point_mesh_squared_distance(VA, VB, FB, dBA, I, C);
  point_mesh_squared_distance(VB, VA, FA, dAB, I, C);
  double mx = 0.0;
  for (int64_t i = 0; i < dBA.n; ++i) mx = dBA.p[i] > mx ? dBA.p[i] : mx;
  for (int64_t i = 0; i < dAB.n; ++i) mx = dAB.p[i] > mx ? dAB.p[i] : mx;
  d = (ScalarD)std::sqrt(mx);
}
}  // namespace iglb200
