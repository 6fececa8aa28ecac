// The grid functions on either side of signed_distance, with the reference's Eigen signatures,
// computed on the GPU (the grid stays in device memory between the stages of offset_surface):
//   igl::voxel_grid      /root/reference/include/igl/voxel_grid.h:30-37, 46-52  (voxel_grid.cpp:14-88)
//   igl::flood_fill      /root/reference/include/igl/flood_fill.h              (flood_fill.cpp:11-98)
//   igl::marching_cubes  /root/reference/include/igl/marching_cubes.h:39-66    (marching_cubes.cpp:19-112)
//   igl::offset_surface  /root/reference/include/igl/offset_surface.h          (offset_surface.cpp:18-48)
// Values are double precision (the reference's templates also instantiate float; those are not offered:
// results would not be bit-identical).
#pragma once
#include <cmath>
#include <cstring>
#include <unordered_map>
#include "signed_distance.h"

namespace iglb200 {
namespace detail {
struct GridHandle {
  iglb200_grid *g = nullptr;
  ~GridHandle() { if (g) iglb200_grid_destroy(g); }
};
struct SurfaceHandle {
  iglb200_surface *s = nullptr;
  ~SurfaceHandle() { if (s) iglb200_surface_destroy(s); }
};
template <class M> inline void require_double() {
  static_assert(std::is_same<typename M::Scalar, double>::value, "iglb200: grid functions take double matrices");
}
template <class DerivedGV> inline void fetch_gv(const iglb200_grid *g, int64_t n, DerivedGV &GV) {
  require_double<DerivedGV>();
  GV.resize(n, 3);
  check(iglb200_grid_get_gv(g, (double *)GV.data(), row_stride(GV), col_stride(GV)));
}
template <class DerivedV, class DerivedF>
inline void fetch_surface(const iglb200_surface *s, DerivedV &V, DerivedF &F, std::unordered_map<std::int64_t, int> *E2V) {
  require_double<DerivedV>();
  static_assert(sizeof(typename DerivedF::Scalar) == 4, "iglb200: F must hold 32-bit ints");
  int64_t nv = 0, nf = 0;
  check(iglb200_surface_info(s, &nv, &nf));
  double *v = new double[(size_t)(3 * nv + 1)];
  int32_t *f = new int32_t[(size_t)(3 * nf + 1)];
  int32_t *e = E2V ? new int32_t[(size_t)(2 * nv + 1)] : nullptr;
  const int rc = iglb200_surface_get(s, v, f, e);
  if (rc == IGLB200_OK) {
    V.resize(nv, 3); F.resize(nf, 3);
    for (int64_t i = 0; i < nv; ++i) for (int d = 0; d < 3; ++d) V(i, d) = v[3 * i + d];
    for (int64_t i = 0; i < nf; ++i) for (int d = 0; d < 3; ++d) F(i, d) = f[3 * i + d];
    if (E2V) {   // march_cube.cpp:52-60: key = i | j << 32 with i < j
      E2V->clear();
      for (int64_t i = 0; i < nv; ++i) (*E2V)[(std::int64_t)e[2 * i] | ((std::int64_t)e[2 * i + 1] << 32)] = (int)i;
    }
  }
  delete[] v; delete[] f; delete[] e;
  check(rc);
}
template <class Derivedside> inline void set_side(const iglb200_grid *g, Derivedside &side, int64_t &n) {
  int sd[3];
  check(iglb200_grid_info(g, sd, &n, nullptr));
  side.resize(3);
  for (int c = 0; c < 3; ++c) side(c) = sd[c];
}
}  // namespace detail

// voxel_grid(box, s, pad_count, GV, side): Box is anything with min()(i) / max()(i), e.g. Eigen::AlignedBox<double,3>
template <class Box, class DerivedGV, class Derivedside>
inline void voxel_grid(const Box &box, const int s, const int pad_count, DerivedGV &GV, Derivedside &side) {
  const double bmin[3] = {(double)box.min()(0), (double)box.min()(1), (double)box.min()(2)};
  const double bmax[3] = {(double)box.max()(0), (double)box.max()(1), (double)box.max()(2)};
  detail::GridHandle h;
  check(iglb200_grid_create(bmin, bmax, s, pad_count, &h.g));
  int64_t n = 0;
  detail::set_side(h.g, side, n);
  detail::fetch_gv(h.g, n, GV);
}
// voxel_grid(V, offset, s, pad_count, GV, side)  (voxel_grid.cpp:68-88)
template <class DerivedV, class DerivedGV, class Derivedside>
inline void voxel_grid(const DerivedV &V, const typename DerivedV::Scalar offset, const int s, const int pad_count, DerivedGV &GV,
                       Derivedside &side) {
  if (V.cols() != 3 || V.rows() == 0) throw Error(IGLB200_ERR_ARG, "iglb200: V must contain positions in 3D");
  struct { double lo[3], hi[3]; const double *min() const { return lo; } const double *max() const { return hi; } } raw;
  for (int c = 0; c < 3; ++c) {
    double lo = (double)V(0, c), hi = lo;
    for (int64_t i = 1; i < (int64_t)V.rows(); ++i) { const double v = (double)V(i, c); lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
    raw.lo[c] = (double)(typename DerivedV::Scalar)(lo - offset);
    raw.hi[c] = (double)(typename DerivedV::Scalar)(hi + offset);
  }
  struct Idx { const double *p; double operator()(int i) const { return p[i]; } };
  struct BoxView { Idx a, b; Idx min() const { return a; } Idx max() const { return b; } } box{{raw.lo}, {raw.hi}};
  voxel_grid(box, s, pad_count, GV, side);
}

// flood_fill(res, S)
template <class Derivedres, class DerivedS>
inline void flood_fill(const Derivedres &res, DerivedS &S) {
  detail::require_double<DerivedS>();
  const int r[3] = {(int)res(0), (int)res(1), (int)res(2)};
  if ((int64_t)S.rows() * (int64_t)S.cols() != (int64_t)r[0] * r[1] * r[2]) throw Error(IGLB200_ERR_ARG, "iglb200: S must hold one value per grid vertex");
  check(iglb200_flood_fill(r, (double *)S.data()));
}

// marching_cubes(S, GV, nx, ny, nz, isovalue, V, F[, E2V])
template <class DerivedS, class DerivedGV, class DerivedV, class DerivedF>
inline void marching_cubes(const DerivedS &S, const DerivedGV &GV, const unsigned nx, const unsigned ny, const unsigned nz,
                           const typename DerivedS::Scalar isovalue, DerivedV &V, DerivedF &F,
                           std::unordered_map<std::int64_t, int> &E2V) {
  detail::require_double<DerivedS>(); detail::require_double<DerivedGV>();
  if ((int64_t)GV.rows() != (int64_t)nx * ny * nz || GV.cols() != 3 || (int64_t)S.rows() * (int64_t)S.cols() != (int64_t)GV.rows())
    throw Error(IGLB200_ERR_ARG, "iglb200: GV must be (nx*ny*nz) x 3 and S nx*ny*nz");
  detail::SurfaceHandle h;
  check(iglb200_marching_cubes((const double *)S.data(), (const double *)GV.data(), detail::row_stride(GV), detail::col_stride(GV),
                               (int)nx, (int)ny, (int)nz, (double)isovalue, &h.s));
  detail::fetch_surface(h.s, V, F, &E2V);
}
template <class DerivedS, class DerivedGV, class DerivedV, class DerivedF>
inline void marching_cubes(const DerivedS &S, const DerivedGV &GV, const unsigned nx, const unsigned ny, const unsigned nz,
                           const typename DerivedS::Scalar isovalue, DerivedV &V, DerivedF &F) {
  std::unordered_map<std::int64_t, int> E2V;
  marching_cubes(S, GV, nx, ny, nz, isovalue, V, F, E2V);
}

// marching_cubes(S, GV, GI, isovalue, V, F[, E2V]): cubes listed by their 8 corner indices (marching_cubes.h:68-92)
template <class DerivedS, class DerivedGV, class DerivedGI, class DerivedV, class DerivedF>
inline typename std::enable_if<!std::is_arithmetic<DerivedGI>::value>::type
marching_cubes(const DerivedS &S, const DerivedGV &GV, const DerivedGI &GI, const typename DerivedS::Scalar isovalue, DerivedV &V,
               DerivedF &F, std::unordered_map<std::int64_t, int> &E2V) {
  detail::require_double<DerivedS>(); detail::require_double<DerivedGV>();
  static_assert(sizeof(typename DerivedGI::Scalar) == 4, "iglb200: GI must hold 32-bit ints");
  if (GV.cols() != 3 || (int64_t)S.rows() * (int64_t)S.cols() != (int64_t)GV.rows() || (GI.rows() && GI.cols() != 8))
    throw Error(IGLB200_ERR_ARG, "iglb200: GV must be #GV x 3, S #GV, GI #cubes x 8");
  detail::SurfaceHandle h;
  check(iglb200_marching_cubes_sparse((const double *)S.data(), (const double *)GV.data(), (int64_t)GV.rows(),
                                      GV.rows() ? detail::row_stride(GV) : 3, GV.rows() ? detail::col_stride(GV) : 1,
                                      (const int32_t *)GI.data(), (int64_t)GI.rows(), GI.rows() ? detail::row_stride(GI) : 8,
                                      GI.rows() ? detail::col_stride(GI) : 1, (double)isovalue, &h.s));
  detail::fetch_surface(h.s, V, F, &E2V);
}
template <class DerivedS, class DerivedGV, class DerivedGI, class DerivedV, class DerivedF>
inline typename std::enable_if<!std::is_arithmetic<DerivedGI>::value>::type
marching_cubes(const DerivedS &S, const DerivedGV &GV, const DerivedGI &GI, const typename DerivedS::Scalar isovalue, DerivedV &V,
               DerivedF &F) {
  std::unordered_map<std::int64_t, int> E2V;
  marching_cubes(S, GV, GI, isovalue, V, F, E2V);
}

// offset_surface(V, F, isolevel, s, signed_distance_type, SV, SF, GV, side, S)
template <class DerivedV, class DerivedF, class isolevelType, class DerivedSV, class DerivedSF, class DerivedGV, class Derivedside,
          class DerivedS>
inline void offset_surface(const DerivedV &V, const DerivedF &F, const isolevelType isolevel, const int s,
                           const SignedDistanceType &signed_distance_type, DerivedSV &SV, DerivedSF &SF, DerivedGV &GV,
                           Derivedside &side, DerivedS &S) {
  detail::require_double<DerivedV>(); detail::require_double<DerivedS>();
  if (V.cols() != 3 || V.rows() == 0) throw Error(IGLB200_ERR_ARG, "iglb200: V must contain positions in 3D");
  double bmin[3], bmax[3];
  for (int c = 0; c < 3; ++c) {
    double lo = V(0, c), hi = lo;
    for (int64_t i = 1; i < (int64_t)V.rows(); ++i) { lo = V(i, c) < lo ? V(i, c) : lo; hi = V(i, c) > hi ? V(i, c) : hi; }
    bmin[c] = lo - (double)isolevel; bmax[c] = hi + (double)isolevel;   // voxel_grid(V, isolevel, s, 1, GV, side)
  }
  detail::GridHandle g;
  check(iglb200_grid_create(bmin, bmax, s, 1, &g.g));
  int64_t n = 0;
  detail::set_side(g.g, side, n);
  double corners[6];
  check(iglb200_grid_info(g.g, nullptr, nullptr, corners));
  const double h = (corners[3] - corners[0]) / ((double)(side(0) - 1));                    // offset_surface.cpp:31-32
  const double lower_bound = (double)isolevel - std::sqrt(3.0) * h, upper_bound = (double)isolevel + std::sqrt(3.0) * h;
  const bool pseudo = signed_distance_type == SIGNED_DISTANCE_TYPE_PSEUDONORMAL;
  const bool exact = signed_distance_type == SIGNED_DISTANCE_TYPE_WINDING_NUMBER || signed_distance_type == SIGNED_DISTANCE_TYPE_DEFAULT;
  iglb200_mesh *m = detail::make_mesh(V, F, 2, (pseudo ? IGLB200_MESH_PSEUDONORMALS : 0) | (exact ? IGLB200_MESH_WINDING_NUMBER : 0));
  int rc = iglb200_grid_signed_distance(g.g, m, (int)signed_distance_type, lower_bound, upper_bound);
  iglb200_mesh_destroy(m);
  check(rc);
  check(iglb200_grid_flood_fill(g.g));
  detail::SurfaceHandle sf;
  check(iglb200_grid_marching_cubes(g.g, (double)isolevel, 0.0, &sf.s));                   // SS = S - isolevel at 0 (:46-47)
  detail::fetch_surface(sf.s, SV, SF, nullptr);
  detail::fetch_gv(g.g, n, GV);
  S.resize(n, 1);
  check(iglb200_grid_get_s(g.g, (double *)S.data()));
}
}  // namespace iglb200
