// se_tracking.cuh -- N1, the tracking front-end either side of the hot path:
//   k_bilateral       bilateralFilterKernel       se_denseslam/src/preprocessing.cpp:42-87
//   k_half_sample     halfSampleRobustImageKernel  preprocessing.cpp:190-226
//   k_depth2vertex    depth2vertexKernel           preprocessing.cpp:89-109
//   k_vertex2normal   vertex2normalKernel<NegY>    preprocessing.cpp:111-159
//   k_track           trackKernel                  se_denseslam/src/tracking.cpp:226-300
//   k_reduce_*        reduceKernel / new_reduce    tracking.cpp:66-224
// The 6x6 solve and the SE3 exponential of updatePoseKernel (tracking.cpp:302-318) run in the single CTA that
// finishes the reduction (k_icp_update), so the whole coarse-to-fine loop is enqueued without a host round trip
// per iteration; only checkPoseKernel (:320-336) is evaluated on the host, from the 32 sums copied back once.
// The reference sums its 32 reduction values with an OpenMP reduction (order undefined); here the
// order is fixed (warp tree -> CTA -> one final CTA), so runs are reproducible, and parity with the
// CPU oracle is to a tolerance (tests/test_gpu_tracking.py), not bit for bit.
#pragma once
#include "se_math.cuh"

namespace se_b200 {

struct TrackData { int result; float error; float J[6]; };      // commons.h:249-253

constexpr float kEDelta = 0.1f;              // constant_parameters.h:17
constexpr int   kFilterRadius = 2;           // :18
constexpr float kDistThreshold = 0.1f;       // :19
constexpr float kNormalThreshold = 0.8f;     // :20
constexpr float kTrackThreshold = 0.15f;     // :21
constexpr float kGaussDelta = 4.0f;          // :34

struct Gauss5 { float g[5]; };

__global__ void __launch_bounds__(256) k_bilateral(float* __restrict__ out, const float* __restrict__ in, int W, int H, Gauss5 gs) {
  pdl_prologue();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const int pos = x + y * W;
  const float center = in[pos];
  if (center == 0.f) { out[pos] = 0.f; return; }
  const float e_d_squared_2 = kEDelta * kEDelta * 2;
  float sum = 0.f, t = 0.f;
#pragma unroll
  for (int i = -kFilterRadius; i <= kFilterRadius; ++i)
#pragma unroll
    for (int j = -kFilterRadius; j <= kFilterRadius; ++j) {
      const int cx = max(0, min(x + i, W - 1)), cy = max(0, min(y + j, H - 1));
      const float curPix = __ldg(in + cx + cy * W);
      if (curPix > 0.f) {
        const float mod = (curPix - center) * (curPix - center);
        const float factor = gs.g[i + kFilterRadius] * gs.g[j + kFilterRadius] * expf(-mod / e_d_squared_2);
        t += factor * curPix;
        sum += factor;
      }
    }
  out[pos] = t / sum;
}

// inW: the parent level's real row stride (in.width(), preprocessing.cpp:209,217) -- 2 outW + 1 when the parent's width is
// odd; the clamp stays at 2 out - 1 as in the reference (:214-216)
__global__ void __launch_bounds__(256) k_half_sample(float* __restrict__ out, const float* __restrict__ in, int outW, int outH, int inW, float e_d, int r) {
  pdl_prologue();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= outW || y >= outH) return;
  const int cx = 2 * x, cy = 2 * y;
  float sum = 0.f, t = 0.f;
  const float center = in[cx + cy * inW];
  for (int i = -r + 1; i <= r; ++i)
    for (int j = -r + 1; j <= r; ++j) {
      const int px = min(max(cx + j, 0), 2 * outW - 1), py = min(max(cy + i, 0), 2 * outH - 1);
      const float current = __ldg(in + px + py * inW);
      if (fabsf(current - center) < e_d) { sum += 1.0f; t += current; }
    }
  out[x + y * outW] = t / sum;
}

__global__ void __launch_bounds__(256) k_depth2vertex(float* __restrict__ vertex, const float* __restrict__ depth, int W, int H, M4 invK) {
  pdl_prologue();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const float d = depth[x + y * W];
  float* o = vertex + 3 * (x + y * W);
  if (d > 0.f) {
    const float vx = (float)x, vy = (float)y;      // (depth * invK) * (x, y, 1, 0)
    o[0] = ((d * invK.m[0]) * vx + (d * invK.m[1]) * vy) + (d * invK.m[2]) * 1.f;
    o[1] = ((d * invK.m[4]) * vx + (d * invK.m[5]) * vy) + (d * invK.m[6]) * 1.f;
    o[2] = ((d * invK.m[8]) * vx + (d * invK.m[9]) * vy) + (d * invK.m[10]) * 1.f;
  } else { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
}

__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// invalid pixels get only .x = INVALID written, as in the reference (the rest keeps its previous value)
__global__ void __launch_bounds__(256) k_vertex2normal(float* __restrict__ out, const float* __restrict__ in, int W, int H, int negY) {
  pdl_prologue();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  float* o = out + 3 * (x + y * W);
  const V3 center = ld3(in + 3 * (x + W * y));
  if (center.z == 0.f) { o[0] = kInvalid; return; }
  const int xl = max(x - 1, 0), xr = min(x + 1, W - 1);
  int yu, yd;
  if (negY) { yu = max(y - 1, 0); yd = min(y + 1, H - 1); }
  else { yd = max(y - 1, 0); yu = min(y + 1, H - 1); }
  const V3 left = ld3(in + 3 * (xl + W * y)), right = ld3(in + 3 * (xr + W * y)), up = ld3(in + 3 * (x + W * yu)), down = ld3(in + 3 * (x + W * yd));
  if (left.z == 0.f || right.z == 0.f || up.z == 0.f || down.z == 0.f) { o[0] = kInvalid; return; }
  const V3 n = normalized3(cross3(right - left, up - down));
  o[0] = n.x; o[1] = n.y; o[2] = n.z;
}

struct TrackParams { M4 view; int inW, inH, refW, refH; float dist_threshold, normal_threshold; };

// device-resident ICP state: pose[16] (row-major, updated in place), then flags
struct IcpState { float pose[16]; int converged; int pad_[3]; };

constexpr int kTrackThreads = 256;

// trackKernel fused with the first level of reduceKernel: every thread builds its TrackData row (written out:
// renderTrack and the tests read it), then the CTA reduces the 32 sums of tracking.cpp:66-200 and writes one
// partial row.  Fixed order: lane tree (shfl_down), then warps in index order.
__global__ void __launch_bounds__(kTrackThreads) k_track(TrackData* __restrict__ output, const float* __restrict__ inVertex, const float* __restrict__ inNormal,
                                                         const float* __restrict__ refVertex, const float* __restrict__ refNormal, TrackParams p,
                                                         const IcpState* __restrict__ st, float* __restrict__ partial /* gridDim.x * 32 */) {
  pdl_prologue();
  __shared__ float s_part[kTrackThreads / 32][32];
  if (st->converged) return;              // updatePoseKernel already returned true at this level: the loop `break`s (DenseSLAMSystem.cpp:182-183)
  M4 Ttrack;
#pragma unroll
  for (int e = 0; e < 16; ++e) Ttrack.m[e] = st->pose[e];
  const int n = p.inW * p.inH;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float s[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) s[k] = 0.f;
  if (i < n) {
    const int px = i % p.inW, py = i / p.inW;
    TrackData row;
    row.result = 0; row.error = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) row.J[k] = 0.f;
    const V3 inN = ld3(inNormal + 3 * (px + py * p.inW));
    if (inN.x == kInvalid) row.result = -1;
    else {
      const V3 projectedVertex = xform3(Ttrack, ld3(inVertex + 3 * (px + py * p.inW)));
      const V3 projectedPos = xform3(p.view, projectedVertex);
      const float ppx = projectedPos.x / projectedPos.z + 0.5f, ppy = projectedPos.y / projectedPos.z + 0.5f;
      if (ppx < 0.f || ppx > (float)(p.refW - 1) || ppy < 0.f || ppy > (float)(p.refH - 1)) row.result = -2;
      else {
        const int rx = (int)ppx, ry = (int)ppy;
        const V3 referenceNormal = ld3(refNormal + 3 * (rx + ry * p.refW));
        if (referenceNormal.x == kInvalid) row.result = -3;
        else {
          const V3 diff = ld3(refVertex + 3 * (rx + ry * p.refW)) - projectedVertex;
          const V3 projectedNormal = rot3(Ttrack, inN);
          if (norm3(diff) > p.dist_threshold) row.result = -4;
          else if (dot3(projectedNormal, referenceNormal) < p.normal_threshold) row.result = -5;
          else {
            row.result = 1;
            row.error = dot3(referenceNormal, diff);
            row.J[0] = referenceNormal.x; row.J[1] = referenceNormal.y; row.J[2] = referenceNormal.z;
            const V3 c = cross3(projectedVertex, referenceNormal);
            row.J[3] = c.x; row.J[4] = c.y; row.J[5] = c.z;
          }
        }
      }
    }
    // the reference leaves error/J of rejected pixels untouched; only `result` is meaningful there
    TrackData* dst = output + (px + py * p.refW);
    if (row.result == 1) *dst = row; else dst->result = row.result;
    if (row.result < 1) {
      s[29] = row.result == -4 ? 1.f : 0.f;
      s[30] = row.result == -5 ? 1.f : 0.f;
      s[31] = row.result > -4 ? 1.f : 0.f;
    } else {
      s[0] = row.error * row.error;
#pragma unroll
      for (int a = 0; a < 6; ++a) s[1 + a] = row.error * row.J[a];
      int k = 7;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) s[k++] = row.J[a] * row.J[b];
      s[28] = 1.f;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    float v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = 0.f;
    for (int w = 0; w < kTrackThreads / 32; ++w) v += s_part[w][threadIdx.x];
    partial[blockIdx.x * 32 + threadIdx.x] = v;
  }
}

// 6x6 Cholesky solve of (J^T J) x = J^T e; vals = b[6] followed by the upper triangle (tracking.cpp:42-64, Eigen::LLT there)
SE_HD bool solve6(const float* vals, float x[6]) {
  float C[6][6], L[6][6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) L[i][j] = 0.f;
  int k = 6;
  for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { C[i][j] = vals[k]; C[j][i] = vals[k]; ++k; }
  for (int j = 0; j < 6; ++j) {
    float d = C[j][j];
    for (int q = 0; q < j; ++q) d -= L[j][q] * L[j][q];
    if (!(d > 0.f)) return false;
    L[j][j] = sqrtf(d);
    for (int i = j + 1; i < 6; ++i) {
      float v = C[i][j];
      for (int q = 0; q < j; ++q) v -= L[i][q] * L[j][q];
      L[i][j] = v / L[j][j];
    }
  }
  float y[6];
  for (int i = 0; i < 6; ++i) { float v = vals[i]; for (int q = 0; q < i; ++q) v -= L[i][q] * y[q]; y[i] = v / L[i][i]; }
  for (int i = 5; i >= 0; --i) { float v = y[i]; for (int q = i + 1; q < 6; ++q) v -= L[q][i] * x[q]; x[i] = v / L[i][i]; }
  return true;
}

// exp: se(3) -> SE(3), x = (upsilon, omega), Rodrigues + V matrix (Sophus::SE3f::exp at tracking.cpp:310)
SE_HD M4 se3_exp(const float x[6]) {
  const float wx = x[3], wy = x[4], wz = x[5];
  const float theta2 = wx * wx + wy * wy + wz * wz, theta = sqrtf(theta2);
  float A, B, Cc;
  if (theta < 1e-4f) { A = 1.f - theta2 / 6.f; B = 0.5f - theta2 / 24.f; Cc = 1.f / 6.f - theta2 / 120.f; }
  else { A = sinf(theta) / theta; B = (1.f - cosf(theta)) / theta2; Cc = (theta - sinf(theta)) / (theta2 * theta); }
  const float W[3][3] = {{0.f, -wz, wy}, {wz, 0.f, -wx}, {-wy, wx, 0.f}};
  float W2[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { W2[i][j] = 0.f; for (int q = 0; q < 3; ++q) W2[i][j] += W[i][q] * W[q][j]; }
  M4 T;
  for (int e = 0; e < 16; ++e) T.m[e] = 0.f;
  for (int i = 0; i < 3; ++i) {
    float tv = 0.f;
    for (int j = 0; j < 3; ++j) {
      const float I = i == j ? 1.f : 0.f;
      T.m[4 * i + j] = I + A * W[i][j] + B * W2[i][j];
      tv += (I + B * W[i][j] + Cc * W2[i][j]) * x[j];
    }
    T.m[4 * i + 3] = tv;
  }
  T.m[15] = 1.f;
  return T;
}

// Second level of reduceKernel + updatePoseKernel (tracking.cpp:208-224, 302-318) in one CTA of 32 x 8 threads:
// the partial rows are summed in a fixed order, then thread 0 solves the 6x6 system, applies exp(x) to the
// device-resident pose and raises `converged` when |x| < icp_threshold (the reference `break`s the level there).
__global__ void __launch_bounds__(256) k_icp_update(const float* __restrict__ partial, int rows, float* __restrict__ out /*32*/,
                                                    IcpState* __restrict__ st, float icp_threshold) {
  pdl_prologue();
  __shared__ float s[8][32];
  __shared__ float r[32];
  if (st->converged) return;
  const int k = threadIdx.x & 31, g = threadIdx.x >> 5;
  float v = 0.f;
  for (int row = g; row < rows; row += 8) v += partial[row * 32 + k];
  s[g][k] = v;
  __syncthreads();
  if (g == 0) { float acc = 0.f; for (int j = 0; j < 8; ++j) acc += s[j][k]; out[k] = acc; r[k] = acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x[6];
    if (!solve6(r + 1, x)) for (int i = 0; i < 6; ++i) x[i] = 0.f;
    M4 pose;
    for (int e = 0; e < 16; ++e) pose.m[e] = st->pose[e];
    pose = mul44(se3_exp(x), pose);
    for (int e = 0; e < 16; ++e) st->pose[e] = pose.m[e];
    float n2 = 0.f;
    for (int i = 0; i < 6; ++i) n2 += x[i] * x[i];
    if (sqrtf(n2) < icp_threshold) st->converged = 1;
  }
}

}  // namespace se_b200
