// Shared plumbing of the Eigen-signature wrappers: how a Derived matrix is handed to the C ABI.
//
// The templates touch only rows() cols() data() resize() innerStride() outerStride() and the
// compile-time flag IsRowMajor, so they compile against Eigen 3.4 / 5.x
// (Eigen::MatrixBase / PlainObjectBase) and against the ~100-line stand-in used by this
// repository's own tests (tests/cpp/mini_eigen.h) — Eigen is not available in the build image.
#pragma once
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "../iglb200.h"

namespace iglb200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != IGLB200_OK) throw Error(rc, std::string("iglb200: ") + iglb200_last_error());
}

namespace detail {
template <class M> struct View {
  const void *ptr;
  int is_f64;
  int64_t rows, cols, rs, cs;  // strides in elements
};
template <class M> inline int64_t row_stride(const M &m) { return M::IsRowMajor ? (int64_t)m.outerStride() : (int64_t)m.innerStride(); }
template <class M> inline int64_t col_stride(const M &m) { return M::IsRowMajor ? (int64_t)m.innerStride() : (int64_t)m.outerStride(); }
template <class S> struct is_real : std::integral_constant<bool, std::is_same<S, double>::value || std::is_same<S, float>::value> {};
template <class M> inline View<M> view(const M &m) {
  using S = typename M::Scalar;
  static_assert(is_real<S>::value, "iglb200: matrices must hold float or double");
  return View<M>{(const void *)m.data(), std::is_same<S, double>::value ? 1 : 0, (int64_t)m.rows(), (int64_t)m.cols(),
                 m.rows() ? row_stride(m) : 1, m.rows() ? col_stride(m) : 1};
}
}  // namespace detail
}  // namespace iglb200
