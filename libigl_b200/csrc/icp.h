// igl::iterative_closest_point on the device (a caller of AABB::squared_distance)
//   /root/reference/include/igl/iterative_closest_point.cpp:40-130   the loop
//   /root/reference/include/igl/random_points_on_mesh.cpp:20-47, random_points_on_mesh_intrinsic.cpp:12-45   sampling
//   /root/reference/include/igl/doublearea.cpp:44-58                 areas of 3-D triangles
//   /root/reference/include/igl/rigid_alignment.cpp:24-108           point-to-plane fit, 5 linearised steps
//   /root/reference/include/igl/per_face_normals.cpp:14-37           NY
// Per outer iteration everything of size #samples or #VX stays on the device: transform, area-weighted sampling,
// closest points on Y (the hot path's own search), normals, and the 6x6 normal equations of every inner step (a
// fixed-order reduction).  Only the 42 numbers of a 6x6 system cross to the host, where they are solved by a
// rank-revealing symmetric eigen-decomposition (what Eigen's CompleteOrthogonalDecomposition does for the reference:
// the minimum-norm least-squares solution, with the same rank test that halves the distance of the regulariser w to 1).
// The reference draws its samples from std::rand()-seeded engines, so no two runs of it agree either; here the stream
// is a counter-based hash of (seed, iteration, sample), which tests restate in numpy.
#pragma once
#include "mesh.h"

namespace icp {

HD uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// uniform in [0,1): 53 random bits
HD double u01(uint64_t seed, uint64_t iter, uint64_t k, uint64_t j) {
  const uint64_t h = mix64(mix64(mix64(seed) ^ (iter * 0x632BE59BD9B4E019ull)) ^ (k * 3ull + j));
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

struct Transform {   // X <- X0 * R + t  (row vectors; R row-major)
  const double *X0; int64_t n; double R[9], t[3]; double *X;   // column-major n x 3
  DEVFN void operator()(int64_t i) const {
    const double x = X0[i], y = X0[n + i], z = X0[2 * n + i];
    X[i] = x * R[0] + y * R[3] + z * R[6] + t[0];
    X[n + i] = x * R[1] + y * R[4] + z * R[7] + t[1];
    X[2 * n + i] = x * R[2] + y * R[5] + z * R[8] + t[2];
  }
};
struct FaceAreas {   // doublearea.cpp:44-58
  const double *V; int64_t nV; const int32_t *F; int64_t nF; double *dblA;
  DEVFN void operator()(int64_t f) const {
    const int32_t a = F[3 * f], b = F[3 * f + 1], c = F[3 * f + 2];
    double s = 0.0;
    for (int d = 0; d < 3; ++d) {
      const int x = d, y = (d + 1) % 3;
      const double rx = V[x * nV + a] - V[x * nV + c], sx = V[x * nV + b] - V[x * nV + c];
      const double ry = V[y * nV + a] - V[y * nV + c], sy = V[y * nV + b] - V[y * nV + c];
      const double p = rx * sy - ry * sx;
      s += p * p;
    }
    dblA[f] = sqrt(s);
  }
};
struct Sample {   // random_points_on_mesh_intrinsic.cpp:21-45 + random_points_on_mesh.cpp:35-46; C = inclusive cumsum of dblA
  const double *V; int64_t nV; const int32_t *F; int64_t nF; const double *C; uint64_t seed, iter; int64_t n; double *X;
  DEVFN void operator()(int64_t k) const {
    const double Cmax = C[nF - 1];
    const double r = u01(seed, iter, (uint64_t)k, 0), s = u01(seed, iter, (uint64_t)k, 1), t = u01(seed, iter, (uint64_t)k, 2);
    // histc: the bin i with C0[i] <= r < C0[i+1], C0 = [0, C/Cmax]
    int64_t lo = 0, hi = nF - 1;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (C[mid] / Cmax > r) hi = mid; else lo = mid + 1; }
    const int64_t f = lo;
    const double st = sqrt(t), b0 = 1.0 - st, b1 = (1.0 - s) * st, b2 = s * st;
    const int32_t a = F[3 * f], b = F[3 * f + 1], c = F[3 * f + 2];
    for (int d = 0; d < 3; ++d) X[d * n + k] = (0.0 + b0 * V[d * nV + a]) + b1 * V[d * nV + b] + b2 * V[d * nV + c];
  }
};
struct Normals {   // N = NY(I, :)
  const double *tri64; const int32_t *I; int64_t n; double *N;
  DEVFN void operator()(int64_t k) const {
    const int32_t f = I[k];
    double nx = 0.0, ny = 0.0, nz = 0.0;
    if (f >= 0) {
      const double *t = tri64 + 9 * (size_t)f;
      const double v1[3] = {t[3] - t[0], t[4] - t[1], t[5] - t[2]}, v2[3] = {t[6] - t[0], t[7] - t[1], t[8] - t[2]};
      nx = v1[1] * v2[2] - v1[2] * v2[1]; ny = v1[2] * v2[0] - v1[0] * v2[2]; nz = v1[0] * v2[1] - v1[1] * v2[0];
      const double r = sqrt(nx * nx + ny * ny + nz * nz);
      if (r == 0) { nx = ny = nz = 0.0; } else { nx /= r; ny /= r; nz /= r; }
    }
    N[k] = nx; N[n + k] = ny; N[2 * n + k] = nz;
  }
};
// Normal equations of rigid_alignment.cpp:48-75 for the current X = X0*R + t:
//   out[0..20]  upper triangle of NA^T NA,  out[21..26] NA^T NB,  out[27..47] upper triangle of A^T A,  out[48..53] A^T B
static constexpr int NSYS = 54;
DEVFN void system_of(const double x[3], const double p[3], const double nn[3], double *acc) {
  // NA row = [x cross n, n], NB = n . (p - x)
  const double na[6] = {x[1] * nn[2] - x[2] * nn[1], x[2] * nn[0] - x[0] * nn[2], x[0] * nn[1] - x[1] * nn[0], nn[0], nn[1], nn[2]};
  const double b[3] = {p[0] - x[0], p[1] - x[1], p[2] - x[2]};
  const double nb = nn[0] * b[0] + nn[1] * b[1] + nn[2] * b[2];
  int q = 0;
  for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) acc[q++] += na[i] * na[j];
  for (int i = 0; i < 6; ++i) acc[q++] += na[i] * nb;
  // A rows: [0, x2, -x1, 1,0,0], [-x2, 0, x0, 0,1,0], [x1, -x0, 0, 0,0,1]
  const double A[3][6] = {{0.0, x[2], -x[1], 1.0, 0.0, 0.0}, {-x[2], 0.0, x[0], 0.0, 1.0, 0.0}, {x[1], -x[0], 0.0, 0.0, 0.0, 1.0}};
  for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) acc[q++] += A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j];
  for (int i = 0; i < 6; ++i) acc[q++] += A[0][i] * b[0] + A[1][i] * b[1] + A[2][i] * b[2];
}
static constexpr int RED_THREADS = 256;
struct Reduce {   // one block; thread j sums samples j, j+256, ... in order, then a fixed tree: deterministic
  const double *X0, *P, *N; int64_t n; double R[9], t[3]; double *out;
#ifdef IGLB200_EMUL
  DEVFN void tile(uint32_t) const {
    double part[RED_THREADS][NSYS];
    for (int j = 0; j < RED_THREADS; ++j) { for (int q = 0; q < NSYS; ++q) part[j][q] = 0.0; for (int64_t k = j; k < n; k += RED_THREADS) one(k, part[j]); }
    for (int o = RED_THREADS / 2; o > 0; o >>= 1) for (int j = 0; j < o; ++j) for (int q = 0; q < NSYS; ++q) part[j][q] += part[j + o][q];
    for (int q = 0; q < NSYS; ++q) out[q] = part[0][q];
  }
#else
  __device__ void tile(uint32_t) const {
    __shared__ double sh[RED_THREADS];
    double acc[NSYS];
    for (int q = 0; q < NSYS; ++q) acc[q] = 0.0;
    for (int64_t k = threadIdx.x; k < n; k += RED_THREADS) one(k, acc);
    for (int q = 0; q < NSYS; ++q) {
      sh[threadIdx.x] = acc[q];
      __syncthreads();
      for (int o = RED_THREADS / 2; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
      if (threadIdx.x == 0) out[q] = sh[0];
      __syncthreads();
    }
  }
#endif
  DEVFN void one(int64_t k, double *acc) const {
    const double x0 = X0[k], y0 = X0[n + k], z0 = X0[2 * n + k];
    const double x[3] = {x0 * R[0] + y0 * R[3] + z0 * R[6] + t[0], x0 * R[1] + y0 * R[4] + z0 * R[7] + t[1], x0 * R[2] + y0 * R[5] + z0 * R[8] + t[2]};
    const double p[3] = {P[k], P[n + k], P[2 * n + k]}, nn[3] = {N[k], N[n + k], N[2 * n + k]};
    system_of(x, p, nn, acc);
  }
};

// ---- host side: 6x6 symmetric systems
inline void mat3_mul(const double *A, const double *B, double *C) {   // row-major
  double T[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  for (int i = 0; i < 9; ++i) C[i] = T[i];
}
// minimum-norm least-squares solution of the symmetric positive semi-definite system Q u = f; returns the rank
inline int solve_sym6(const double Q[36], const double f[6], double u[6]) {
  double A[36], Vv[36];
  for (int i = 0; i < 36; ++i) { A[i] = Q[i]; Vv[i] = 0.0; }
  for (int i = 0; i < 6; ++i) Vv[7 * i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {   // cyclic Jacobi
    double off = 0.0;
    for (int i = 0; i < 6; ++i) for (int j = i + 1; j < 6; ++j) off += A[6 * i + j] * A[6 * i + j];
    if (off < 1e-300) break;
    for (int p = 0; p < 6; ++p) for (int q = p + 1; q < 6; ++q) {
      const double apq = A[6 * p + q];
      if (apq == 0.0) continue;
      const double theta = (A[6 * q + q] - A[6 * p + p]) / (2.0 * apq);
      const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
      for (int k = 0; k < 6; ++k) { const double akp = A[6 * k + p], akq = A[6 * k + q]; A[6 * k + p] = c * akp - s * akq; A[6 * k + q] = s * akp + c * akq; }
      for (int k = 0; k < 6; ++k) { const double apk = A[6 * p + k], aqk = A[6 * q + k]; A[6 * p + k] = c * apk - s * aqk; A[6 * q + k] = s * apk + c * aqk; }
      for (int k = 0; k < 6; ++k) { const double vkp = Vv[6 * k + p], vkq = Vv[6 * k + q]; Vv[6 * k + p] = c * vkp - s * vkq; Vv[6 * k + q] = s * vkp + c * vkq; }
    }
  }
  double lmax = 0.0;
  for (int i = 0; i < 6; ++i) lmax = std::max(lmax, std::fabs(A[7 * i]));
  const double thr = lmax * 6.0 * 2.220446049250313e-16;   // Eigen's default: epsilon * size, relative to the largest pivot
  int rank = 0;
  for (int i = 0; i < 6; ++i) u[i] = 0.0;
  for (int e = 0; e < 6; ++e) {
    const double lam = A[7 * e];
    if (!(std::fabs(lam) > thr)) continue;
    ++rank;
    double dotp = 0.0;
    for (int k = 0; k < 6; ++k) dotp += Vv[6 * k + e] * f[k];
    for (int k = 0; k < 6; ++k) u[k] += Vv[6 * k + e] * dotp / lam;
  }
  return rank;
}

// rigid_alignment(X,P,N,R,t): X0 = the samples, P their closest points, N the normals there (all column-major n x 3, device)
inline void rigid_alignment(const double *X0, const double *P, const double *N, int64_t n, double *d_sys, double Rout[9], double tout[3], stream_t s) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
  double w = 1e-5;
  for (int it = 0; it < 5; ++it) {
    Reduce red{X0, P, N, n, {R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8]}, {t[0], t[1], t[2]}, d_sys};
#ifdef IGLB200_EMUL
    red.tile(0);
#else
    plat::k_tiles_one<<<1, RED_THREADS, 0, s>>>(red);
    plat::launch_counter()++;
    PLAT_CHECK(cudaGetLastError());
#endif
    double h[NSYS];
    plat::d2h(h, d_sys, sizeof(h), s);
    plat::sync(s);
    double Q[36], f[6];
    int q = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { Q[6 * i + j] = Q[6 * j + i] = (1.0 - w) * h[q]; ++q; }
    for (int i = 0; i < 6; ++i) f[i] = (1.0 - w) * h[q++];
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { Q[6 * i + j] += w * h[q]; if (i != j) Q[6 * j + i] += w * h[q]; ++q; }
    for (int i = 0; i < 6; ++i) f[i] += w * h[q++];
    double u[6];
    if (solve_sym6(Q, f, u) < 6) w = 1.0 - (1.0 - w) / 2.0;
    const double W[9] = {0, u[2], -u[1], -u[2], 0, u[0], u[1], -u[0], 0};
    const double x = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double Ri[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (x != 0) {
      double WW[9];
      mat3_mul(W, W, WW);
      const double a = std::sin(x) / x, b = (1.0 - std::cos(x)) / (x * x);
      for (int k = 0; k < 9; ++k) Ri[k] += a * W[k] + b * WW[k];
    }
    // R = R*Ri; t = t*Ri + ti
    const double tn[3] = {t[0] * Ri[0] + t[1] * Ri[3] + t[2] * Ri[6] + u[3], t[0] * Ri[1] + t[1] * Ri[4] + t[2] * Ri[7] + u[4],
                          t[0] * Ri[2] + t[1] * Ri[5] + t[2] * Ri[8] + u[5]};
    mat3_mul(R, Ri, R);
    t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
  }
  for (int k = 0; k < 9; ++k) Rout[k] = R[k];
  for (int k = 0; k < 3; ++k) tout[k] = t[k];
}

}  // namespace icp
