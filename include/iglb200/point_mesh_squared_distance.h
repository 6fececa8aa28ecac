// igl::point_mesh_squared_distance(P,V,Ele,sqrD,I,C)
//   /root/reference/include/igl/point_mesh_squared_distance.h:54-60, .cpp:44-46 (AABB::init + squared_distance)
// Triangles only; degenerate rows [i j j] / [i i i] act as segments / points (header :23-30).
#pragma once
#include "AABB.h"

namespace iglb200 {
template <class DerivedP, class DerivedV, class DerivedEle, class DerivedsqrD, class DerivedI, class DerivedC>
inline void point_mesh_squared_distance(const DerivedP &P, const DerivedV &V, const DerivedEle &Ele, DerivedsqrD &sqrD, DerivedI &I,
                                        DerivedC &C) {
  AABB<DerivedV, 3> tree;
  tree.init(V, Ele);
  tree.squared_distance(V, Ele, P, sqrD, I, C);
}
}  // namespace iglb200
