// Prints results of the stand-in Eigen / Sophus operations the reference's path relies on, as JSON, for
// tests/test_ref_standin.py to compare with numpy / scipy.  Test infrastructure only.
#include <Eigen/Dense>
#include <sophus/se3.hpp>

#include <cstdio>

template <class M> void dump(const char* name, const M& m, bool last = false) {
  std::printf("\"%s\": [", name);
  for (int r = 0; r < M::rows(); ++r) {
    std::printf("[");
    for (int c = 0; c < M::cols(); ++c) std::printf("%.9g%s", (double)m(r, c), c + 1 < M::cols() ? ", " : "");
    std::printf("]%s", r + 1 < M::rows() ? ", " : "");
  }
  std::printf("]%s\n", last ? "" : ",");
}

int main() {
  using namespace Eigen;
  std::printf("{\n");
  Matrix4f K;
  K << 481.2f, 0, 320.f, 0, 0, -480.f, 240.f, 0, 0, 0, 1, 0, 0, 0, 0, 1;
  dump("K", K); dump("K_inv", K.inverse());
  Matrix<float, 6, 1> xi; xi << 0.1f, -0.2f, 0.3f, 0.4f, -0.5f, 0.6f;
  const Matrix4f T = Sophus::SE3f::exp(xi).matrix();
  dump("xi", xi); dump("T", T); dump("T_inv", T.inverse()); dump("T_inv_sophus", Sophus::SE3f(T).inverse().matrix());
  Matrix4f G;
  G << 2, 1, 0, 3, -1, 4, 2, 0, 0.5f, -2, 5, 1, 1, 0, -1, 3;
  dump("G", G); dump("G_inv", G.inverse()); dump("KT", K * T);
  const Vector3f p(0.3f, -1.2f, 2.5f), q(-0.7f, 0.4f, 1.1f);
  dump("p", p); dump("q", q); dump("Tp", Sophus::SE3f(T) * p); dump("cross", p.cross(q)); dump("normalized", p.normalized());
  dump("hom", (T * p.homogeneous()).head<3>()); dump("cwise", p.cwiseProduct(q).cwiseMax(Vector3f::Constant(-0.3f)));
  dump("floor", Vector3f((p * 1.7f).array().floor()));
  Vector3f clamped = p * 3.f;
  {  // math_utils.h:112-116 clamp through MatrixBase
    MatrixBase<Vector3f>& res = clamped;
    res = res.array().max(Vector3f::Constant(-1.f).array());
    res = res.array().min(Vector3f::Constant(2.f).array());
  }
  dump("clamp", clamped);
  Matrix<float, 6, 6> A = Matrix<float, 6, 6>::Zero();
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A(i, j) = (i == j ? 10.f + i : 0.f) + 0.3f * (float)((i * 7 + j * 3) % 5) + 0.3f * (float)((j * 7 + i * 3) % 5);
  Matrix<float, 6, 1> b; b << 1, 2, 3, 4, 5, 6;
  LLT<Matrix<float, 6, 6>> llt; llt.compute(A);
  dump("A", A); dump("b", b); dump("llt_x", llt.solve(b));
  float buf[8 * 32];
  for (int i = 0; i < 8 * 32; ++i) buf[i] = (float)i;
  Map<Matrix<float, 8, 32, RowMajor>> rows(buf);
  for (int j = 1; j < 8; ++j) rows.row(0) += rows.row(j);
  Matrix<float, 1, 27> seg = rows.row(0).segment(1, 27);
  dump("rowsum_seg", seg);
  Matrix4f pose = T; pose.block<3, 1>(0, 3) += Vector3f(1, 2, 3);
  dump("block_add", pose);
  Vector3i vi(3, -4, 9);
  std::printf("\"bools\": [%d, %d, %d, %d],\n", (int)((vi.array() >= Vector3i(3, -5, 9).array()) && (vi.array() <= Vector3i(3, 0, 9).array())).all(),
              (int)((vi.array() >= Vector3i(4, -5, 9).array()) * (vi.array() <= Vector3i(3, 0, 9).array())).all(), (int)(p.array() == 0).all(), (int)T.isApprox(T));
  dump("cast", (p * 2.6f).cast<int>(), true);
  std::printf("}\n");
  return 0;
}
