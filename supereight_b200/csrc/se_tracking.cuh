// se_tracking.cuh -- N1, the tracking front-end either side of the hot path:
//   k_bilateral       bilateralFilterKernel       se_denseslam/src/preprocessing.cpp:42-87
//   k_half_sample     halfSampleRobustImageKernel  preprocessing.cpp:190-226
//   k_depth2vertex    depth2vertexKernel           preprocessing.cpp:89-109
//   k_vertex2normal   vertex2normalKernel<NegY>    preprocessing.cpp:111-159
//   k_track           trackKernel                  se_denseslam/src/tracking.cpp:226-300
//   k_reduce_*        reduceKernel / new_reduce    tracking.cpp:66-224
// The 6x6 solve and the SE3 exponential of updatePoseKernel (tracking.cpp:302-318) and
// checkPoseKernel (:320-336) run on the host (se_b200.cu), as small as they are.
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

__global__ void __launch_bounds__(256) k_half_sample(float* __restrict__ out, const float* __restrict__ in, int outW, int outH, float e_d, int r) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= outW || y >= outH) return;
  const int inW = outW * 2, cx = 2 * x, cy = 2 * y;
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

struct TrackParams { M4 Ttrack, view; int inW, inH, refW, refH; float dist_threshold, normal_threshold; };

constexpr int kTrackThreads = 256;

// trackKernel fused with the first level of reduceKernel: every thread builds its TrackData row (written out:
// renderTrack and the tests read it), then the CTA reduces the 32 sums of tracking.cpp:66-200 and writes one
// partial row.  Fixed order: lane tree (shfl_down), then warps in index order.
__global__ void __launch_bounds__(kTrackThreads) k_track(TrackData* __restrict__ output, const float* __restrict__ inVertex, const float* __restrict__ inNormal,
                                                         const float* __restrict__ refVertex, const float* __restrict__ refNormal, TrackParams p,
                                                         float* __restrict__ partial /* gridDim.x * 32 */) {
  __shared__ float s_part[kTrackThreads / 32][32];
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
      const V3 projectedVertex = xform3(p.Ttrack, ld3(inVertex + 3 * (px + py * p.inW)));
      const V3 projectedPos = xform3(p.view, projectedVertex);
      const float ppx = projectedPos.x / projectedPos.z + 0.5f, ppy = projectedPos.y / projectedPos.z + 0.5f;
      if (ppx < 0.f || ppx > (float)(p.refW - 1) || ppy < 0.f || ppy > (float)(p.refH - 1)) row.result = -2;
      else {
        const int rx = (int)ppx, ry = (int)ppy;
        const V3 referenceNormal = ld3(refNormal + 3 * (rx + ry * p.refW));
        if (referenceNormal.x == kInvalid) row.result = -3;
        else {
          const V3 diff = ld3(refVertex + 3 * (rx + ry * p.refW)) - projectedVertex;
          const V3 projectedNormal = rot3(p.Ttrack, inN);
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

// second level: one CTA of 32 x 8 threads sums the partial rows in a fixed order
__global__ void __launch_bounds__(256) k_reduce_final(const float* __restrict__ partial, int rows, float* __restrict__ out /*32*/) {
  __shared__ float s[8][32];
  const int k = threadIdx.x & 31, g = threadIdx.x >> 5;
  float v = 0.f;
  for (int r = g; r < rows; r += 8) v += partial[r * 32 + k];
  s[g][k] = v;
  __syncthreads();
  if (g == 0) { float t = 0.f; for (int j = 0; j < 8; ++j) t += s[j][k]; out[k] = t; }
}

}  // namespace se_b200
