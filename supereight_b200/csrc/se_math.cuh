// se_math.cuh -- small fixed-order float algebra and the octree key codec, host+device.
//
// Arithmetic contract of this library (DESIGN.md "Arithmetic contract"): IEEE binary32,
// round-to-nearest, no fused multiply-add (the translation unit is compiled with
// -fmad=false and host code with -ffp-contract=off), dot products and matrix rows summed
// left to right.  The reference leaves these orders to Eigen/Sophus; fixing them makes the
// allocation set and every float result reproducible bit for bit.
//
// Reference behaviour followed (paths relative to the supereight tree):
//   key codec      se_core/include/se/utils/morton_utils.hpp:37-72,
//                  se_core/include/se/octant_ops.hpp:41-57,107-113,
//                  se_core/include/se/octree_defines.h:38-80
//   camera matrix  se_denseslam/include/se/commons.h:255-271
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define SE_HD __host__ __device__ __forceinline__

namespace se_b200 {

constexpr int kBlockSide = 8;
constexpr int kBlockVoxels = 512;
constexpr int kMaxBits = 21;
constexpr int kCastStackDepth = 23;
constexpr unsigned long long kScaleMask = 0x1FFull;
constexpr float kNearPlane = 0.4f;     // constant_parameters.h:27
constexpr float kFarPlane = 4.0f;      // constant_parameters.h:32
constexpr float kMaxWeight = 100.f;    // DenseSLAMSystem.cpp:235
constexpr float kInvalid = -2.f;       // commons.h:71
constexpr float kAmbient = 0.1f;       // constant_parameters.h:37

struct V3 { float x, y, z; };
struct M4 { float m[16]; };            // row-major

SE_HD float mat(const M4& A, int r, int c) { return A.m[4 * r + c]; }
SE_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
SE_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
SE_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
SE_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
SE_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
SE_HD V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
SE_HD float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
SE_HD float norm3(V3 a) { return sqrtf(dot3(a, a)); }
SE_HD V3 normalized3(V3 a) { const float n2 = dot3(a, a); if (n2 > 0.f) return a / sqrtf(n2); return a; }

SE_HD M4 mul44(const M4& A, const M4& B) {
  M4 C;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      C.m[4 * i + j] = ((mat(A, i, 0) * mat(B, 0, j) + mat(A, i, 1) * mat(B, 1, j)) + mat(A, i, 2) * mat(B, 2, j)) + mat(A, i, 3) * mat(B, 3, j);
  return C;
}
// top-left 3x3 times vector
SE_HD V3 rot3(const M4& A, V3 v) {
  return v3((mat(A, 0, 0) * v.x + mat(A, 0, 1) * v.y) + mat(A, 0, 2) * v.z,
            (mat(A, 1, 0) * v.x + mat(A, 1, 1) * v.y) + mat(A, 1, 2) * v.z,
            (mat(A, 2, 0) * v.x + mat(A, 2, 1) * v.y) + mat(A, 2, 2) * v.z);
}
// top 3x4 times (v, 1)
SE_HD V3 xform3(const M4& A, V3 v) {
  return v3(((mat(A, 0, 0) * v.x + mat(A, 0, 1) * v.y) + mat(A, 0, 2) * v.z) + mat(A, 0, 3),
            ((mat(A, 1, 0) * v.x + mat(A, 1, 1) * v.y) + mat(A, 1, 2) * v.z) + mat(A, 1, 3),
            ((mat(A, 2, 0) * v.x + mat(A, 2, 1) * v.y) + mat(A, 2, 2) * v.z) + mat(A, 2, 3));
}
// inverse of a rigid camera-to-world pose (R^T, -R^T t); plays the role of
// Sophus::SE3f(pose).inverse() at DenseSLAMSystem.cpp:237
SE_HD M4 rigid_inverse(const M4& T) {
  M4 R;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[4 * i + j] = mat(T, j, i);
  for (int i = 0; i < 3; ++i)
    R.m[4 * i + 3] = -((mat(T, 0, i) * mat(T, 0, 3) + mat(T, 1, i) * mat(T, 1, 3)) + mat(T, 2, i) * mat(T, 2, 3));
  R.m[12] = 0.f; R.m[13] = 0.f; R.m[14] = 0.f; R.m[15] = 1.f;
  return R;
}
SE_HD M4 camera_matrix(const float k[4]) {
  M4 K; for (int i = 0; i < 16; ++i) K.m[i] = 0.f;
  K.m[0] = k[0]; K.m[2] = k[2]; K.m[5] = k[1]; K.m[6] = k[3]; K.m[10] = 1.f; K.m[15] = 1.f;
  return K;
}
SE_HD M4 inverse_camera_matrix(const float k[4]) {
  M4 K; for (int i = 0; i < 16; ++i) K.m[i] = 0.f;
  K.m[0] = 1.0f / k[0]; K.m[2] = -k[2] / k[0]; K.m[5] = 1.0f / k[1]; K.m[6] = -k[3] / k[1]; K.m[10] = 1.f; K.m[15] = 1.f;
  return K;
}

// ---- 63-bit Morton keys: code | level in the low 9 bits ------------------------------
SE_HD unsigned long long spread3(unsigned long long v) {
  unsigned long long x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8)  & 0x100f00f00f00f00full;
  x = (x | x << 4)  & 0x10c30c30c30c30c3ull;
  x = (x | x << 2)  & 0x1249249249249249ull;
  return x;
}
SE_HD unsigned long long squeeze3(unsigned long long v) {
  unsigned long long x = v & 0x1249249249249249ull;
  x = (x | x >> 2)  & 0x10c30c30c30c30c3ull;
  x = (x | x >> 4)  & 0x100f00f00f00f00full;
  x = (x | x >> 8)  & 0x1f0000ff0000ffull;
  x = (x | x >> 16) & 0x1f00000000ffffull;
  x = (x | x >> 32) & 0x1fffffull;
  return x;
}
SE_HD unsigned long long morton_encode(int x, int y, int z) {
  return spread3((unsigned long long)(long long)x) | (spread3((unsigned long long)(long long)y) << 1) | (spread3((unsigned long long)(long long)z) << 2);
}
SE_HD void morton_decode(unsigned long long c, int& x, int& y, int& z) {
  x = (int)squeeze3(c); y = (int)squeeze3(c >> 1); z = (int)squeeze3(c >> 2);
}
// MASK[i] of octree_defines.h:58-80: the top 3(i+1) of the 63 code bits
SE_HD unsigned long long level_mask(int i) {
  const int bits = 3 * (i + 1);
  return ((1ull << bits) - 1ull) << (63 - bits);
}
SE_HD unsigned long long key_code(unsigned long long k) { return k & ~kScaleMask; }
SE_HD int key_level(unsigned long long k) { return (int)(k & kScaleMask); }
SE_HD unsigned long long key_encode(int x, int y, int z, int level, int max_depth) {
  return (morton_encode(x, y, z) & level_mask(kMaxBits - max_depth + level - 1)) | (unsigned long long)level;
}
SE_HD int key_child_id(unsigned long long code, int level, int max_depth) {
  return (int)((key_code(code) >> (3 * (max_depth - level))) & 7ull);
}
SE_HD bool key_descendant(unsigned long long octant, unsigned long long ancestor, int max_depth) {
  const int level = key_level(ancestor);
  const unsigned long long a = key_code(ancestor);
  const unsigned long long o = key_code(octant) & level_mask(kMaxBits - max_depth + level - 1);
  return (a ^ o) == 0ull;
}

}  // namespace se_b200
