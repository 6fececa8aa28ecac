// ~100-line stand-in for Eigen::Matrix — TESTS ONLY.  Eigen is not installed in the build image
// (libigl fetches 5.0.1 over the network, cmake/recipes/external/eigen.cmake:8-13); the wrappers in
// include/iglb200/ use only the members provided here, which real Eigen matrices have too.
#pragma once
#include <cstdint>
#include <vector>

namespace mini {
enum { ColMajor = 0, RowMajor = 1 };
template <class T, int Options = ColMajor> class Matrix {
 public:
  using Scalar = T;
  static constexpr bool IsRowMajor = (Options == RowMajor);
  Matrix() = default;
  Matrix(int64_t r, int64_t c) { resize(r, c); }
  void resize(int64_t r, int64_t c) { r_ = r; c_ = c; d_.assign((size_t)(r * c), T()); }
  void resize(int64_t n) { resize(n, 1); }   // vectors
  int64_t rows() const { return r_; }
  int64_t cols() const { return c_; }
  int64_t size() const { return r_ * c_; }
  T *data() { return d_.data(); }
  const T *data() const { return d_.data(); }
  int64_t innerStride() const { return 1; }
  int64_t outerStride() const { return IsRowMajor ? c_ : r_; }
  T &operator()(int64_t i, int64_t j) { return d_[(size_t)(IsRowMajor ? i * c_ + j : j * r_ + i)]; }
  const T &operator()(int64_t i, int64_t j) const { return d_[(size_t)(IsRowMajor ? i * c_ + j : j * r_ + i)]; }
  T &operator()(int64_t i) { return d_[(size_t)i]; }
  const T &operator()(int64_t i) const { return d_[(size_t)i]; }
  template <class U, int O2> void assign_from(const Matrix<U, O2> &o) {
    resize(o.rows(), o.cols());
    for (int64_t i = 0; i < r_; ++i) for (int64_t j = 0; j < c_; ++j) (*this)(i, j) = (T)o(i, j);
  }
 private:
  int64_t r_ = 0, c_ = 0;
  std::vector<T> d_;
};
using MatrixXd = Matrix<double>;
using MatrixXf = Matrix<float>;
using MatrixXi = Matrix<int>;
using VectorXd = Matrix<double>;
using VectorXi = Matrix<int>;
using RowMatrixXd = Matrix<double, RowMajor>;
using RowMatrixXi = Matrix<int, RowMajor>;
}  // namespace mini
