// igl::AABB<DerivedV,3>, the subset on the hot path: init + batch squared_distance
//   /root/reference/include/igl/AABB.h:494-496 (init), :722-728 (squared_distance batch)
// Ray casts, dynamic insertion/rotation and find() are other queries and are not provided.
#pragma once
#include "fast_winding_number.h"

namespace iglb200 {

template <class DerivedV, int DIM> class AABB {
  static_assert(DIM == 3, "iglb200::AABB: only 3-D trees are implemented on the device");
 public:
  iglb200_mesh *mesh = nullptr;
  AABB() = default;
  AABB(const AABB &) = delete;
  AABB &operator=(const AABB &) = delete;
  ~AABB() { clear(); }
  void clear() { if (mesh) iglb200_mesh_destroy(mesh); mesh = nullptr; }

  // AABB::init(V,Ele): builds the tree on the device.  The handle also carries the winding-number
  // moments, so the same object serves signed_distance_fast_winding_number.
  template <class DerivedEle> void init(const DerivedV &V, const DerivedEle &Ele) {
    clear();
    mesh = detail::make_mesh(V, Ele, 2);
  }

  // AABB::squared_distance(V,Ele,P,sqrD,I,C)  (AABB.cpp:1241-1265).  V and Ele are accepted for
  // signature compatibility; the handle already holds its own copy.
  template <class DerivedEle, class DerivedP, class DerivedsqrD, class DerivedI, class DerivedC>
  void squared_distance(const DerivedV & /*V*/, const DerivedEle & /*Ele*/, const DerivedP &P, DerivedsqrD &sqrD, DerivedI &I,
                        DerivedC &C) const {
    if (!mesh) throw Error(IGLB200_ERR_ARG, "iglb200::AABB: init() was not called");
    static_assert(sizeof(typename DerivedI::Scalar) == 4, "iglb200: I must hold 32-bit ints");
    sqrD.resize(P.rows(), 1);
    I.resize(P.rows(), 1);
    C.resize(P.rows(), 3);
    auto p = detail::view(P);
    auto c = detail::view(C);
    check(iglb200_squared_distance(mesh, p.ptr, p.is_f64, p.rows, p.rs, p.cs, (void *)sqrD.data(),
                                   std::is_same<typename DerivedsqrD::Scalar, double>::value ? 1 : 0, (int32_t *)I.data(),
                                   (void *)C.data(), c.is_f64, c.rs, c.cs));
  }
};

}  // namespace iglb200
