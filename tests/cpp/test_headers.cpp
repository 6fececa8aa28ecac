// The reference's own known-answer tests for the hot path, re-created against include/iglb200/*.h
// (Catch2 and the libigl test data are not available here):
//   tests/include/igl/signed_distance.cpp:4-43      single_tet
//   tests/include/igl/fast_winding_number.cpp:17-53 one_point_cloud
//   tests/include/igl/AABB.cpp:100-113              barycentres of a soup
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "mini_eigen.h"
#include "iglb200/signed_distance.h"
#include "iglb200/point_mesh_squared_distance.h"
#include "iglb200/hausdorff.h"
#include "iglb200/offset_surface.h"
#include "iglb200/winding_number.h"
#include "iglb200/iterative_closest_point.h"

using namespace mini;
#define REQUIRE(x) do { if (!(x)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #x); std::exit(1); } } while (0)

static void tet(MatrixXd &V, MatrixXi &F) {
  V.resize(4, 3); F.resize(4, 3);
  const double v[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  const int f[4][3] = {{0, 1, 3}, {0, 2, 1}, {0, 3, 2}, {1, 2, 3}};
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) { V(i, j) = v[i][j]; F(i, j) = f[i][j]; }
}

int main() {
  MatrixXd V; MatrixXi F;
  tet(V, F);
  MatrixXd P(1, 3);
  P(0, 0) = P(0, 1) = P(0, 2) = 0.5;
  // signed_distance: single_tet
  // (the reference test loops over exactly these five: tests/include/igl/signed_distance.cpp:23-30)
  for (auto type : {iglb200::SIGNED_DISTANCE_TYPE_PSEUDONORMAL, iglb200::SIGNED_DISTANCE_TYPE_UNSIGNED,
                    iglb200::SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, iglb200::SIGNED_DISTANCE_TYPE_WINDING_NUMBER,
                    iglb200::SIGNED_DISTANCE_TYPE_DEFAULT}) {
    VectorXd S; VectorXi I; MatrixXd C, N;
    iglb200::signed_distance(P, V, F, type, S, I, C, N);
    REQUIRE(S.rows() == 1 && I.rows() == 1 && C.rows() == 1 && C.cols() == 3);
    const double tol = type == iglb200::SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER ? 1e-7 : 1e-15;
    REQUIRE(std::fabs(S(0) - std::sqrt(1. / 12.)) <= tol);
    REQUIRE(I(0) == 3);
    if (type == iglb200::SIGNED_DISTANCE_TYPE_PSEUDONORMAL) {   // the only sign type that fills N
      REQUIRE(N.rows() == 1 && N.cols() == 3);
      for (int d = 0; d < 3; ++d) REQUIRE(std::fabs(N(0, d) - std::sqrt(1. / 3.)) < 1e-15);
    } else {
      REQUIRE(N.rows() == 0);   // untouched, as in the reference
    }
  }
  // batch pseudonormal overload with prebuilt structures (normals arguments are accepted, not read)
  {
    iglb200::AABB<MatrixXd, 3> tree; tree.init(V, F);
    MatrixXd FN, VN, EN, C, N; VectorXi EMAP, I; VectorXd S;
    iglb200::signed_distance_pseudonormal(P, V, F, tree, FN, VN, EN, EMAP, S, I, C, N);
    REQUIRE(std::fabs(S(0) - std::sqrt(1. / 12.)) <= 1e-15 && I(0) == 3 && N.rows() == 1);
  }
  // an unknown sign type must fail loudly
  {
    VectorXd S; VectorXi I; MatrixXd C, N;
    bool threw = false;
    try { iglb200::signed_distance(P, V, F, (iglb200::SignedDistanceType)9, S, I, C, N); }
    catch (const iglb200::Error &e) { threw = (e.code == IGLB200_ERR_UNSUPPORTED); }
    REQUIRE(threw);
  }
  // row-major and float inputs give the same answer
  {
    RowMatrixXd Vr, Pr; RowMatrixXi Fr; MatrixXf Pf;
    Vr.assign_from(V); Fr.assign_from(F); Pr.assign_from(P); Pf.assign_from(P);
    VectorXd S, S2, W; VectorXi I; MatrixXd C, N; RowMatrixXd Cr;
    iglb200::signed_distance(Pr, Vr, Fr, iglb200::SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, S, I, Cr, N);
    iglb200::signed_distance_with_winding_number(Pf, V, F, S2, I, C, W);
    REQUIRE(S(0) == S2(0));
    REQUIRE(std::fabs(W(0)) < 1e-6);
    for (int d = 0; d < 3; ++d) REQUIRE(Cr(0, d) == C(0, d));
  }
  // empty F (signed_distance.cpp:398-405)
  {
    MatrixXi F0(0, 3);
    VectorXd S; VectorXi I; MatrixXd C, N;
    iglb200::signed_distance(P, V, F0, iglb200::SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, S, I, C, N);
    REQUIRE(std::isnan(S(0)) && I(0) == -1 && std::isnan(C(0, 0)));
  }
  // AABB: barycentres of a soup -> I == face, d == 0; plus the prebuilt-structure overloads
  {
    MatrixXd BC(4, 3), C;
    for (int f = 0; f < 4; ++f) for (int d = 0; d < 3; ++d) BC(f, d) = (V(F(f, 0), d) + V(F(f, 1), d) + V(F(f, 2), d)) / 3.0;
    iglb200::AABB<MatrixXd, 3> tree;
    tree.init(V, F);
    VectorXd sqrD; VectorXi I;
    tree.squared_distance(V, F, BC, sqrD, I, C);
    for (int f = 0; f < 4; ++f) { REQUIRE(I(f) == f); REQUIRE(sqrD(f) <= 1e-30); }
    {   // single-query overload
      MatrixXd q(1, 3), c1; q(0) = BC(3, 0); q(1) = BC(3, 1); q(2) = BC(3, 2);
      int i1 = -1;
      const double d1 = tree.squared_distance(V, F, q, i1, c1);
      REQUIRE(i1 == 3 && d1 == sqrD(3) && c1(0) == C(3, 0) && c1(1) == C(3, 1) && c1(2) == C(3, 2));
    }
    VectorXd sqrD2; VectorXi I2; MatrixXd C2;
    iglb200::point_mesh_squared_distance(BC, V, F, sqrD2, I2, C2);
    for (int f = 0; f < 4; ++f) REQUIRE(I2(f) == I(f) && sqrD2(f) == sqrD(f));
    iglb200::FastWindingNumberBVH bvh;
    iglb200::fast_winding_number(V, F, 2, bvh);
    VectorXd S;
    iglb200::signed_distance_fast_winding_number(P, V, F, tree, bvh, S);
    REQUIRE(std::fabs(S(0) - std::sqrt(1. / 12.)) <= 1e-7);
    MatrixXd Q(2, 3);
    Q(0, 0) = Q(0, 1) = Q(0, 2) = 0.1; Q(1, 0) = Q(1, 1) = Q(1, 2) = 2.0;
    VectorXd W, W2;
    iglb200::fast_winding_number(bvh, 2.0f, Q, W);
    iglb200::fast_winding_number(V, F, Q, W2);
    REQUIRE(std::fabs(W(0) - 1.0) < 1e-5 && std::fabs(W(1)) < 1e-3 && W(0) == W2(0) && W(1) == W2(1));
  }
  // segments (Ele with two columns) and hausdorff
  {
    MatrixXi E(1, 2); E(0, 0) = 0; E(0, 1) = 1;       // the edge (0,0,0)-(1,0,0)
    MatrixXd Q(1, 3), C; Q(0, 0) = 0.5; Q(0, 1) = 2.0; Q(0, 2) = 0.0;
    VectorXd sqrD; VectorXi I;
    iglb200::point_mesh_squared_distance(Q, V, E, sqrD, I, C);
    REQUIRE(sqrD(0) == 4.0 && I(0) == 0 && C(0, 0) == 0.5 && C(0, 1) == 0.0);
    MatrixXd V2; V2.assign_from(V);
    for (int i = 0; i < 4; ++i) V2(i, 0) += 0.25;
    double d = -1;
    iglb200::hausdorff(V, F, V2, F, d);
    REQUIRE(std::fabs(d - 0.25) < 1e-15);
  }
  // fast_winding_number: one_point_cloud
  {
    MatrixXd Pc(1, 3), Nc(1, 3), Q(4, 3); VectorXd A(1, 1), WN, WN2;
    Pc(0, 0) = 0.1; Pc(0, 1) = 0.2; Pc(0, 2) = 0.3;
    const double nrm = std::sqrt(16. + 25. + 36.);
    Nc(0, 0) = 4 / nrm; Nc(0, 1) = 5 / nrm; Nc(0, 2) = 6 / nrm;
    A(0) = 0.7;
    const double q[4][3] = {{0, 0, 0}, {4, -3, 2}, {-1, 1, 3}, {0, 5, -2}};
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) Q(i, j) = q[i][j];
    iglb200::fast_winding_number(Pc, Nc, A, Q, 2, 2.0, WN);
    iglb200::fast_winding_number(Pc, Nc, A, Q, WN2);
    const double gold[4] = {0.38779369004261133, -0.00041235296362485, -0.00362978253577090, -0.00041235296362485};
    for (int i = 0; i < 4; ++i) { REQUIRE(std::fabs(WN(i) - gold[i]) < 1e-15); REQUIRE(WN2(i) == WN(i)); }
  }
  // exact winding number (igl::winding_number): 1 inside the tet, 0 outside
  {
    MatrixXd O(3, 3); VectorXd W;
    O(0, 0) = O(0, 1) = O(0, 2) = 0.1; O(1, 0) = O(1, 1) = O(1, 2) = 2.0; O(2, 0) = O(2, 1) = O(2, 2) = -0.2;
    iglb200::winding_number(V, F, O, W);
    REQUIRE(std::fabs(W(0) - 1.0) < 1e-14 && std::fabs(W(1)) < 1e-14 && std::fabs(W(2)) < 1e-14);
    MatrixXi F1(1, 3); F1(0, 0) = 1; F1(0, 1) = 2; F1(0, 2) = 3;       // the slanted face alone, seen from the origin:
    MatrixXd O1(1, 3); O1(0, 0) = O1(0, 1) = O1(0, 2) = 0.0;            // one octant of the sphere = 1/8
    iglb200::winding_number(V, F1, O1, W);
    REQUIRE(std::fabs(std::fabs(W(0)) - 0.125) < 1e-14);
  }
  // voxel_grid / flood_fill / marching_cubes / offset_surface (tutorials 705, 909; offset_surface.cpp)
  {
    MatrixXd GV; VectorXi side;
    iglb200::voxel_grid(V, 0.25, 12, 1, GV, side);
    REQUIRE(side(0) == 12 && side(1) == 12 && side(2) == 12 && GV.rows() == 12 * 12 * 12 && GV.cols() == 3);
    RowMatrixXd GVr; VectorXi side2;
    iglb200::voxel_grid(V, 0.25, 12, 1, GVr, side2);
    for (int64_t i = 0; i < GV.rows(); i += 97) for (int d = 0; d < 3; ++d) REQUIRE(GV(i, d) == GVr(i, d));
    // sphere of radius 0.45 about the grid centre, NaN away from a band: fill, then polygonise
    VectorXd S(GV.rows(), 1);
    double c[3];
    for (int d = 0; d < 3; ++d) c[d] = 0.5 * (GV(0, d) + GV(GV.rows() - 1, d));
    for (int64_t i = 0; i < GV.rows(); ++i) {
      const double r = std::sqrt((GV(i, 0) - c[0]) * (GV(i, 0) - c[0]) + (GV(i, 1) - c[1]) * (GV(i, 1) - c[1]) + (GV(i, 2) - c[2]) * (GV(i, 2) - c[2]));
      S(i) = std::fabs(r - 0.45) < 0.3 ? r - 0.45 : std::nan("");
    }
    iglb200::flood_fill(side, S);
    for (int64_t i = 0; i < S.rows(); ++i) REQUIRE(S(i) == S(i));
    REQUIRE(S(0) > 0);
    MatrixXd SV; MatrixXi SF;
    std::unordered_map<std::int64_t, int> E2V;
    iglb200::marching_cubes(S, GV, side(0), side(1), side(2), 0.0, SV, SF, E2V);
    REQUIRE(SV.rows() > 100 && SF.rows() == 2 * SV.rows() - 4);          // closed genus-0 surface
    REQUIRE((int64_t)E2V.size() == SV.rows());
    for (int64_t i = 0; i < SV.rows(); ++i) {
      const double r = std::sqrt((SV(i, 0) - c[0]) * (SV(i, 0) - c[0]) + (SV(i, 1) - c[1]) * (SV(i, 1) - c[1]) + (SV(i, 2) - c[2]) * (SV(i, 2) - c[2]));
      REQUIRE(std::fabs(r - 0.45) < 0.02);
    }
    {   // the same surface from an explicit list of all cubes (marching_cubes(S,GV,GI,...))
      const int nx = side(0), ny = side(1), nz = side(2);
      RowMatrixXi GI((nx - 1) * (ny - 1) * (nz - 1), 8);
      const int off[8] = {0, 1, 1 + nx, nx, nx * ny, 1 + nx * ny, 1 + nx + nx * ny, nx + nx * ny};
      int64_t c = 0;
      for (int z = 0; z < nz - 1; ++z) for (int y = 0; y < ny - 1; ++y) for (int x = 0; x < nx - 1; ++x, ++c)
        for (int k = 0; k < 8; ++k) GI(c, k) = x + nx * (y + ny * z) + off[k];
      MatrixXd SVs; MatrixXi SFs;
      iglb200::marching_cubes(S, GV, GI, 0.0, SVs, SFs);
      REQUIRE(SVs.rows() == SV.rows() && SFs.rows() == SF.rows());
      for (int64_t i = 0; i < SV.rows(); ++i) for (int d = 0; d < 3; ++d) REQUIRE(SVs(i, d) == SV(i, d));
      for (int64_t i = 0; i < SF.rows(); ++i) for (int d = 0; d < 3; ++d) REQUIRE(SFs(i, d) == SF(i, d));
    }
    MatrixXd SV2, GV2; MatrixXi SF2; VectorXi side3; VectorXd S2;
    iglb200::offset_surface(V, F, 0.2, 16, iglb200::SIGNED_DISTANCE_TYPE_PSEUDONORMAL, SV2, SF2, GV2, side3, S2);
    REQUIRE(SV2.rows() > 50 && SF2.rows() == 2 * SV2.rows() - 4 && GV2.rows() == S2.rows());
    for (int64_t i = 0; i < S2.rows(); ++i) REQUIRE(S2(i) == S2(i));
  }
  // #F == 0: S = NaN, I = -1, C = NaN and N = NaN (signed_distance.cpp:398-405)
  {
    MatrixXi F0(0, 3); VectorXd S; VectorXi I; MatrixXd C, N;
    iglb200::signed_distance(P, V, F0, iglb200::SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, S, I, C, N);
    REQUIRE(S(0) != S(0) && I(0) == -1 && C(0, 1) != C(0, 1) && N.rows() == 1 && N.cols() == 3 && N(0, 2) != N(0, 2));
  }
  // single-query and bounded overloads (fast_winding_number.h:196-206, signed_distance.h:300-312, AABB.h:575-614)
  {
    iglb200::AABB<MatrixXd, 3> tree; tree.init(V, F);
    iglb200::FastWindingNumberBVH bvh;
    iglb200::fast_winding_number(V, F, 2, bvh);
    MatrixXd q(1, 3);
    q(0, 0) = q(0, 1) = q(0, 2) = 0.1;                       // inside the tet
    REQUIRE(std::fabs(iglb200::fast_winding_number(bvh, 2.0f, q) - 1.0) < 1e-6);
    q(0, 0) = q(0, 1) = q(0, 2) = 0.5;                       // outside, sqrt(1/12) from face 3
    REQUIRE(std::fabs(iglb200::fast_winding_number(bvh, 2.0f, q)) < 1e-6);
    REQUIRE(std::fabs(iglb200::signed_distance_fast_winding_number(q, V, F, tree, bvh) - std::sqrt(1. / 12.)) < 1e-7);
    int i = -7; MatrixXd c(1, 3); c(0, 0) = c(0, 1) = c(0, 2) = -7.0;
    REQUIRE(std::fabs(tree.squared_distance(V, F, q, 1.0, i, c) - 1. / 12.) < 1e-15 && i == 3);        // found below the bound
    i = -7; c(0, 0) = -7.0;
    REQUIRE(tree.squared_distance(V, F, q, 0.0, 0.05, i, c) == 0.05 && i == -7 && c(0, 0) == -7.0);       // nothing nearer than up_sqr_d: untouched
    REQUIRE(tree.squared_distance(V, F, q, 2.0, 1.0, i, c) == 2.0 && i == -7);                           // low > up returns low (AABB.cpp:1145-1148)
  }
  // cached point-cloud evaluation: one precompute, many evaluations (fast_winding_number.h:44-53, 88-100 as an opaque handle)
  {
    MatrixXd Pp(1, 3), Np(1, 3), Q(4, 3); VectorXd Ap(1, 1), W1, W2;
    Pp(0, 0) = 0.1; Pp(0, 1) = 0.2; Pp(0, 2) = 0.3;
    const double nl = std::sqrt(16. + 25. + 36.);
    Np(0, 0) = 4 / nl; Np(0, 1) = 5 / nl; Np(0, 2) = 6 / nl;
    Ap(0) = 0.7;
    const double qs[4][3] = {{0, 0, 0}, {4, -3, 2}, {-1, 1, 3}, {0, 5, -2}};
    for (int r = 0; r < 4; ++r) for (int d = 0; d < 3; ++d) Q(r, d) = qs[r][d];
    iglb200::FastWindingNumberCloud cloud;
    iglb200::fast_winding_number(Pp, Np, Ap, 2, cloud);
    iglb200::fast_winding_number(cloud, Q, 2.0, W1);
    iglb200::fast_winding_number(cloud, Q, 2.0, W2);
    const double gold[4] = {0.38779369004261133, -0.00041235296362485, -0.00362978253577090, -0.00041235296362485};
    for (int r = 0; r < 4; ++r) REQUIRE(std::fabs(W1(r) - gold[r]) < 1e-15 && W1(r) == W2(r));
  }
  // iterative_closest_point (iterative_closest_point.h:36-82): a shifted copy of the tet is pulled back onto it
  {
    MatrixXd VX(V.rows(), 3), R; MatrixXd t;
    for (int64_t i = 0; i < V.rows(); ++i) { VX(i, 0) = V(i, 0) + 0.02; VX(i, 1) = V(i, 1) - 0.01; VX(i, 2) = V(i, 2) + 0.015; }
    iglb200::iterative_closest_point(VX, F, V, F, 200, 8, R, t);
    REQUIRE(R.rows() == 3 && R.cols() == 3 && t.rows() == 1 && t.cols() == 3);
    double worst = 0;
    for (int64_t i = 0; i < V.rows(); ++i) for (int d = 0; d < 3; ++d) {
      const double y = VX(i, 0) * R(0, d) + VX(i, 1) * R(1, d) + VX(i, 2) * R(2, d) + t(0, d);
      worst = std::fmax(worst, std::fabs(y - V(i, d)));
    }
    REQUIRE(worst < 5e-3);
    iglb200::AABB<MatrixXd, 3> Ytree; Ytree.init(V, F);
    MatrixXd NY, R2, t2;
    iglb200::iterative_closest_point(VX, F, V, F, Ytree, NY, 200, 8, R2, t2);
    REQUIRE(R2.rows() == 3 && t2.cols() == 3);
  }
  std::printf("OK\n");
  return 0;
}
