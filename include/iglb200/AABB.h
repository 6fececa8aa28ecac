// igl::AABB<DerivedV,3>, the subset on the hot path: init + batch squared_distance
//   /root/reference/include/igl/AABB.h:494-496 (init), :722-728 (squared_distance batch)
//   :561-566 (one query, unbounded), :575-614 (one query within [low_sqr_d, up_sqr_d)).  Ray casts, dynamic
// insertion/rotation and find() are other queries and are not provided.
#pragma once
#include <limits>
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

  // AABB::squared_distance(V,Ele,p,i,c) -> sqr_d, one query (AABB.h:561-566, AABB.cpp:1107-1116).  Functional but a
  // whole device round trip per point: batch the queries wherever possible.  p and c are 1x3 (or 3x1) vectors.
  template <class DerivedEle, class Derivedp, class Derivedc>
  typename DerivedV::Scalar squared_distance(const DerivedV &V, const DerivedEle &Ele, const Derivedp &p, int &i, Derivedc &c) const {
    return squared_distance(V, Ele, p, (typename DerivedV::Scalar)0, std::numeric_limits<typename DerivedV::Scalar>::infinity(), i, c);
  }
  // AABB::squared_distance(V,Ele,p,up_sqr_d,i,c) (AABB.h:603-614, AABB.cpp:1107-1118)
  template <class DerivedEle, class Derivedp, class Derivedc>
  typename DerivedV::Scalar squared_distance(const DerivedV &V, const DerivedEle &Ele, const Derivedp &p, typename DerivedV::Scalar up_sqr_d,
                                             int &i, Derivedc &c) const {
    return squared_distance(V, Ele, p, (typename DerivedV::Scalar)0, up_sqr_d, i, c);
  }
  // AABB::squared_distance(V,Ele,p,low_sqr_d,up_sqr_d,i,c) (AABB.h:575-593, AABB.cpp:1135-1218): the closest primitive
  // nearer than up_sqr_d; if there is none the result is up_sqr_d and i, c are left untouched, as in the reference.
  // (low_sqr_d lets the reference stop at ANY primitive below it; here the closest one is returned.)
  template <class DerivedEle, class Derivedp, class Derivedc>
  typename DerivedV::Scalar squared_distance(const DerivedV & /*V*/, const DerivedEle & /*Ele*/, const Derivedp &p,
                                             typename DerivedV::Scalar low_sqr_d, typename DerivedV::Scalar up_sqr_d, int &i, Derivedc &c) const {
    if (!mesh) throw Error(IGLB200_ERR_ARG, "iglb200::AABB: init() was not called");
    if ((int64_t)p.rows() * (int64_t)p.cols() != 3) throw Error(IGLB200_ERR_ARG, "iglb200::AABB: p must have 3 entries");
    const double q[3] = {(double)p(0), (double)p(1), (double)p(2)};
    double d2 = 0.0, cc[3] = {0.0, 0.0, 0.0};
    int32_t ii = -1;
    check(iglb200_squared_distance_bounded(mesh, q, 1, 1, 3, 1, (double)low_sqr_d, (double)up_sqr_d, &d2, 1, &ii, cc, 1, 3, 1));
    if (ii >= 0) {
      i = (int)ii;
      if ((int64_t)c.rows() * (int64_t)c.cols() != 3) c.resize(1, 3);
      for (int d = 0; d < 3; ++d) c(d) = (typename Derivedc::Scalar)cc[d];
    }
    return (typename DerivedV::Scalar)d2;
  }
};

}  // namespace iglb200
