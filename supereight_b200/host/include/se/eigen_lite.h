// eigen_lite.h -- the handful of Eigen fixed-size types the DenseSLAMSystem interface uses
// (Vector2i, Vector3i, Vector3f, Vector4f, Matrix4f), for builds where Eigen3 is not installed
// (this image: Eigen is absent and there is no network).  With Eigen available, define
// SE_B200_USE_EIGEN (or just have <Eigen/Dense> on the include path) and the real types are used;
// the shim only relies on operator()(i), operator()(r,c), x()/y()/z()/w() and Identity(), which
// both provide, and never on the storage order.
#pragma once
#if defined(SE_B200_USE_EIGEN) || (defined(__has_include) && __has_include(<Eigen/Dense>))
#include <Eigen/Dense>
#else
#include <cmath>
#include <cstddef>
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
namespace Eigen {
template <typename T, int N> struct LiteVector {
  T v[N];
  LiteVector() { for (int i = 0; i < N; ++i) v[i] = T(0); }
  LiteVector(T a, T b) { static_assert(N == 2, "size"); v[0] = a; v[1] = b; }
  LiteVector(T a, T b, T c) { static_assert(N == 3, "size"); v[0] = a; v[1] = b; v[2] = c; }
  LiteVector(T a, T b, T c, T d) { static_assert(N == 4, "size"); v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
  static LiteVector Constant(T c) { LiteVector r; for (int i = 0; i < N; ++i) r.v[i] = c; return r; }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  T& x() { return v[0]; } const T& x() const { return v[0]; }
  T& y() { return v[1]; } const T& y() const { return v[1]; }
  T& z() { static_assert(N >= 3, "size"); return v[2]; } const T& z() const { static_assert(N >= 3, "size"); return v[2]; }
  T& w() { static_assert(N >= 4, "size"); return v[3]; } const T& w() const { static_assert(N >= 4, "size"); return v[3]; }
  T* data() { return v; } const T* data() const { return v; }
  LiteVector operator+(const LiteVector& o) const { LiteVector r; for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i]; return r; }
  LiteVector operator-(const LiteVector& o) const { LiteVector r; for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i]; return r; }
  LiteVector operator*(T s) const { LiteVector r; for (int i = 0; i < N; ++i) r.v[i] = v[i] * s; return r; }
  LiteVector operator/(T s) const { LiteVector r; for (int i = 0; i < N; ++i) r.v[i] = v[i] / s; return r; }
};
struct Matrix4f {
  float m[16];   // row-major here; real Eigen is column-major -- the shim never looks at data()
  Matrix4f() { for (float& f : m) f = 0.f; }
  static Matrix4f Identity() { Matrix4f r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
  float& operator()(int r, int c) { return m[4 * r + c]; }
  const float& operator()(int r, int c) const { return m[4 * r + c]; }
  // Eigen semantics: ||a - b||^2 <= prec^2 * min(||a||^2, ||b||^2)
  bool isApprox(const Matrix4f& o, float prec = 1e-5f) const {
    float d = 0.f, na = 0.f, nb = 0.f;
    for (int i = 0; i < 16; ++i) { d += (m[i] - o.m[i]) * (m[i] - o.m[i]); na += m[i] * m[i]; nb += o.m[i] * o.m[i]; }
    return d <= prec * prec * (na < nb ? na : nb);
  }
};
typedef LiteVector<int, 2> Vector2i;
typedef LiteVector<int, 3> Vector3i;
typedef LiteVector<float, 2> Vector2f;
typedef LiteVector<float, 3> Vector3f;
typedef LiteVector<float, 4> Vector4f;
}  // namespace Eigen
#endif
