// Stand-in for <sophus/se3.hpp> -- TEST INFRASTRUCTURE ONLY (oracle/_ref build; see ../Eigen/Dense).
// Only what the reference's per-frame path touches: SE3f from a 4x4 pose, inverse(), rotationMatrix(), matrix(), the action
// on a point (projective_functor.hpp:76-118, DenseSLAMSystem.cpp:237,249) and exp() (tracking.cpp:310).
// Stored as rotation matrix + translation (real Sophus keeps a unit quaternion: a few ulp apart).
#pragma once
#include <Eigen/Dense>
#include <cmath>

namespace Sophus {

template <class T> class SE3 {
  Eigen::Matrix<T, 3, 3> r_;
  Eigen::Matrix<T, 3, 1> t_;
 public:
  SE3() : r_(Eigen::Matrix<T, 3, 3>::Identity()) {}
  SE3(const Eigen::Matrix<T, 3, 3>& r, const Eigen::Matrix<T, 3, 1>& t) : r_(r), t_(t) {}
  explicit SE3(const Eigen::Matrix<T, 4, 4>& m) : r_(m.template topLeftCorner<3, 3>()), t_(m.template topRightCorner<3, 1>()) {}
  const Eigen::Matrix<T, 3, 3>& rotationMatrix() const { return r_; }
  const Eigen::Matrix<T, 3, 1>& translation() const { return t_; }
  Eigen::Matrix<T, 4, 4> matrix() const {
    Eigen::Matrix<T, 4, 4> m = Eigen::Matrix<T, 4, 4>::Identity();
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) m(i, j) = r_(i, j); m(i, 3) = t_(i); }
    return m;
  }
  SE3 inverse() const {                                         // [R^T | -(R^T t)]
    const Eigen::Matrix<T, 3, 3> rt = r_.transpose();
    return SE3(rt, -(rt * t_));
  }
  Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { return r_ * p + t_; }
  SE3 operator*(const SE3& o) const { return SE3(r_ * o.r_, r_ * o.t_ + t_); }

  // exp: se(3) -> SE(3), tangent = (upsilon, omega): Rodrigues rotation and the V matrix applied to upsilon
  static SE3 exp(const Eigen::Matrix<T, 6, 1>& x) {
    const T wx = x(3), wy = x(4), wz = x(5);
    const T theta2 = wx * wx + wy * wy + wz * wz, theta = std::sqrt(theta2);
    T A, B, C;
    if (theta < T(1e-4)) { A = T(1) - theta2 / T(6); B = T(0.5) - theta2 / T(24); C = T(1) / T(6) - theta2 / T(120); }
    else { A = std::sin(theta) / theta; B = (T(1) - std::cos(theta)) / theta2; C = (theta - std::sin(theta)) / (theta2 * theta); }
    Eigen::Matrix<T, 3, 3> W;
    W(0, 1) = -wz; W(0, 2) = wy; W(1, 0) = wz; W(1, 2) = -wx; W(2, 0) = -wy; W(2, 1) = wx;
    const Eigen::Matrix<T, 3, 3> W2 = W * W, I = Eigen::Matrix<T, 3, 3>::Identity();
    const Eigen::Matrix<T, 3, 3> R = I + W * A + W2 * B, V = I + W * B + W2 * C;
    return SE3(R, V * x.template head<3>());
  }
};
using SE3f = SE3<float>;
using SE3d = SE3<double>;

}  // namespace Sophus
