// se_integrate_staged.cuh -- EXPERIMENT, compiled only with -DSE_INT_STAGE_SLICES=<8|4|2> (not part of the default build).
//
// k_integrate_sdf with the pipeline stage size as a parameter: a stage is SE_INT_STAGE_SLICES z slices of a block instead
// of the whole 4 KiB payload.  With 4 slices a warp needs 2 x 2 KiB of shared memory instead of 2 x 4 KiB, and the kernel
// fits 4 CTAs (32 warps, 64 registers, no spills) per SM instead of 3 (24 warps): at 512^3 the ~4 600 active blocks of the
// headline workload then take one round of the 4 736 resident warps instead of 1.3 rounds of 3 552.
// State: bit-exact in the CPU test tier (all SDF parity tests, library built with the define); with 8 slices the
// SASS has the same opcode histogram as the default kernel.  NOT yet timed or run on the device: the round's GPU budget
// ran out (the A/B is scripts/build_variants.sh + scripts/ab_variants.sh).  Included by se_kernels.cuh in place of the default kernel when the macro is set.
#pragma once

constexpr int kIntegrateWarps = 8;                        // warps per CTA
// A pipeline stage is kStageSlices z slices of a block (8 = the whole 4 KiB payload, 4 = half of it).  Each warp owns two
// stage buffers, so the stage size sets the shared memory per warp and with it the resident warps per SM:
// 8 slices -> 8 KiB per warp, 3 CTAs (24 warps) per SM;  4 slices -> 4 KiB per warp, 4 CTAs (32 warps, 64 registers).
#ifndef SE_INT_STAGE_SLICES
#define SE_INT_STAGE_SLICES 8
#endif
constexpr int kStageSlices = SE_INT_STAGE_SLICES;
constexpr int kStagesPerBlock = kBlockSide / kStageSlices;
constexpr int kStageVoxels = kStageSlices * kBlockSide * kBlockSide;
constexpr unsigned kStageBytes = kStageVoxels * (unsigned)sizeof(SdfVoxel);
constexpr int kIntegrateSmem = kIntegrateWarps * 2 * (int)kStageBytes;
constexpr int kIntegrateMinCtas = kStageSlices == 8 ? 3 : 4;
static_assert(kStageSlices == 8 || kStageSlices == 4 || kStageSlices == 2, "a stage is a power-of-two number of z slices");

// One warp per active VoxelBlock, persistent (grid = SMs x resident CTAs, grid-stride over the active
// list).  Each warp runs a two-stage pipeline: while it fuses one stage out of one shared-memory buffer, the TMA
// engine streams the next stage -- the rest of the block, or the start of the warp's next block -- (cp.async.bulk,
// one elected lane, mbarrier completion) into the other, so the HBM/L2 latency of the payload never stalls the math.
// Lane l owns voxels x = 2(l&3), 2(l&3)+1 of row y = l>>2 in each z slice: one conflict-free LDS.128 per slice, and one
// fully coalesced 512 B STG.128 per warp for every slice that changed.
template <bool FAST>
__global__ void __launch_bounds__(kIntegrateWarps * 32, kIntegrateMinCtas) k_integrate_sdf(MapView<SdfVoxel> m, const float* __restrict__ depth, IntegrateParams p, const int* __restrict__ list, int parity) {
  pdl_prologue();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ unsigned long long bars[kIntegrateWarps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n = m.counters[kCntActive0 + parity];
  float4* const buf0 = reinterpret_cast<float4*>(smem_raw) + warp * (2 * kStageVoxels / 2);
  if (lane == 0) { mbar_init(&bars[warp][0], 1); mbar_init(&bars[warp][1], 1); }
  mbar_init_fence();
  __syncwarp();

  const int y = lane >> 2, x0 = (lane & 3) * 2;
  // per-lane constants: x * delta and x * cameraDelta for the lane's two voxel columns
  const float xf0 = (float)x0, xf1 = (float)(x0 + 1);
  const float d0x = xf0 * p.delta.x, d0y = xf0 * p.delta.y, d0z = xf0 * p.delta.z;
  const float d1x = xf1 * p.delta.x, d1y = xf1 * p.delta.y, d1z = xf1 * p.delta.z;
  const float c0x = xf0 * p.cameraDelta.x, c0y = xf0 * p.cameraDelta.y;
  const float c1x = xf1 * p.cameraDelta.x, c1y = xf1 * p.cameraDelta.y;
  const float K00 = p.K.m[0], K02 = p.K.m[2], K11 = p.K.m[5], K12 = p.K.m[6];
  const float rmu = rcp_rn<FAST>(p.mu);

  // item i of round k goes to warp (i - k * warps), warps numbered warp-major ACROSS the CTAs: a partial last round
  // (n is rarely a multiple of the warp count) then lands on every CTA / SM equally instead of on the first CTAs only
#ifndef SE_INT_CTAMAJOR
  int i = warp * gridDim.x + blockIdx.x;
#else
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
#endif
  int b = 0;
  int4 c = make_int4(0, 0, 0, 0);
  if (i < n) {
    b = list[i];
    c = m.block_coord[b];
    if (lane == 0) {
      mbar_expect_tx(&bars[warp][0], kStageBytes);
      bulk_copy_g2s(buf0, m.block_data + (size_t)b * kBlockVoxels, kStageBytes, &bars[warp][0]);
    }
  }
  // s counts this warp's stages: stage s lives in buffer s & 1 and completes phase (s >> 1) of barrier s & 1
  for (int s = 0; i < n; i += warps) {
    const int inext = i + warps;
    int bn = 0;
    int4 cn = make_int4(0, 0, 0, 0);
    if (inext < n) {
      bn = list[inext];
      cn = m.block_coord[bn];
    }
    float4* data = reinterpret_cast<float4*>(m.block_data + (size_t)b * kBlockVoxels);
    // start = Tcw * (px, py, pz): the x/y part of each row sum is the same for the 8 slices
    const float px = (float)c.x * p.voxelSize, py = (float)(c.y + y) * p.voxelSize;
    const float sx01 = p.Tcw.m[0] * px + p.Tcw.m[1] * py;
    const float sy01 = p.Tcw.m[4] * px + p.Tcw.m[5] * py;
    const float sz01 = p.Tcw.m[8] * px + p.Tcw.m[9] * py;
    bool visible = false;
#pragma unroll
    for (int part = 0; part < kStagesPerBlock; ++part, ++s) {
      // start the copy of the stage after this one into the other buffer: the next slices of this block, or the first
      // ones of the warp's next block.  That buffer was last read two stages ago; the __syncwarp (within a block) and the
      // __any_sync below (between blocks) order those reads before the refill.
      const bool last = part == kStagesPerBlock - 1;
      if (part > 0) __syncwarp();
      if (!last || inext < n) {
        if (lane == 0) {
          unsigned long long* bar = &bars[warp][(s + 1) & 1];
          const SdfVoxel* src = last ? m.block_data + (size_t)bn * kBlockVoxels : m.block_data + (size_t)b * kBlockVoxels + (part + 1) * kStageVoxels;
          mbar_expect_tx(bar, kStageBytes);
          bulk_copy_g2s(buf0 + ((s + 1) & 1) * (kStageVoxels / 2), src, kStageBytes, bar);
        }
      }
      // fuse the current stage out of shared memory
      const float4* sbuf = buf0 + (s & 1) * (kStageVoxels / 2);
      mbar_wait(&bars[warp][s & 1], (unsigned)((s >> 1) & 1));
#pragma unroll
      for (int zs = 0; zs < kStageSlices; ++zs) {
        const int z = part * kStageSlices + zs;
        const float pz = (float)(c.z + z) * p.voxelSize;
        const float sx = (sx01 + p.Tcw.m[2] * pz) + p.Tcw.m[3];
        const float sy = (sy01 + p.Tcw.m[6] * pz) + p.Tcw.m[7];
        const float sz = (sz01 + p.Tcw.m[10] * pz) + p.Tcw.m[11];
        // camerastart = K3 * start with K = [[fx,0,cx],[0,fy,cy],[0,0,1]]: the zero terms add exact zeros
        const float csx = K00 * sx + K02 * sz, csy = K11 * sy + K12 * sz;
        float4 v = sbuf[zs * 32 + lane];
        bool changed = false;
        if (FAST) {
          sdf_voxel_pair(v, visible, changed, sx, sy, sz, csx, csy, f2(d0x, d1x), f2(d0y, d1y), f2(d0z, d1z), f2(c0x, c1x), f2(c0y, c1y), depth, p, rmu);
        } else {
          sdf_voxel<FAST>(v.x, v.y, visible, changed, sx + d0x, sy + d0y, sz + d0z, csx + c0x, csy + c0y, depth, p, rmu);
          sdf_voxel<FAST>(v.z, v.w, visible, changed, sx + d1x, sy + d1y, sz + d1z, csx + c1x, csy + c1y, depth, p, rmu);
        }
        if (changed) data[z * 32 + lane] = v;
      }
    }
    const bool any = __any_sync(0xffffffffu, visible);      // also orders this block's smem reads before the buffer is refilled
    if (lane == 0) m.block_active[b] = any ? 1 : 0;           // projective_functor.hpp:110
    b = bn; c = cn;
  }
  update_nodes(m, depth, p);                                   // a12, projective_functor.hpp:152-155
}

