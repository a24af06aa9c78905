// se_ptx.cuh of the CPU test tier: the inline-PTX wrappers of supereight_b200/csrc/se_ptx.cuh restated in C++ for the
// fiber executor (tests/simt_emu/include/cuda_runtime.h).  TEST INFRASTRUCTURE ONLY.
//   * MUFU.RCP / MUFU.RSQ are replaced by correctly rounded values; the refinement sequences built on them
//     (rcp_rn / div_rn / sqrt_rn<true>) then return the correctly rounded results they return on the device.  That the
//     REAL approximations also do is a property of the hardware and is what the GPU tests establish bit for bit.
//   * packed fp32: the two halves computed one after the other with the same IEEE operations.
//   * TMA bulk copy: a memcpy by the issuing lane that completes one phase of the mbarrier; the wait polls and yields.
//   * programmatic dependent launch: launches are synchronous here, nothing to do.
#pragma once
#include <cuda_runtime.h>

namespace se_b200 {

__device__ __forceinline__ void pdl_prologue() {}

__device__ __forceinline__ int ld_relaxed(const int* p) { return *(const volatile int*)p; }
__device__ __forceinline__ void poll_backoff() { simt::yield(); }

__device__ __forceinline__ int ldg_if(bool pred, const int* p, int otherwise) { return pred ? *p : otherwise; }

__device__ __forceinline__ float mufu_rcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float mufu_rsq(float x) { return (float)(1.0 / std::sqrt((double)x)); }

__device__ __forceinline__ unsigned long long pk(float2 a) { return (unsigned long long)__float_as_uint(a.x) | ((unsigned long long)__float_as_uint(a.y) << 32); }
__device__ __forceinline__ float2 upk(unsigned long long v) { return make_float2(__uint_as_float((unsigned)v), __uint_as_float((unsigned)(v >> 32))); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float add_rz_emu(float a, float b) {
  const int old = std::fegetround();
  std::fesetround(FE_TOWARDZERO);
  volatile float va = a, vb = b;
  volatile float r = va + vb;
  std::fesetround(old);
  return r;
}
__device__ __forceinline__ float2 add2_rz(float2 a, float2 b) { return make_float2(add_rz_emu(a.x, b.x), add_rz_emu(a.y, b.y)); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }

// mbarrier word: number of completed phases (one producer, transaction-count completion only)
// shared-window address: the offset from the CTA's dynamic shared memory (the only thing the kernels address this way)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)((const unsigned char*)p - simt::dynamic_smem()); }
__device__ __forceinline__ float4 lds128(unsigned addr) { return *reinterpret_cast<const float4*>((simt::dynamic_smem() + addr)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long*, unsigned) {}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  if (((size_t)dst | (size_t)src | bytes) & 15) simt::die("cp.async.bulk needs 16-byte aligned addresses and size");
  std::memcpy(dst, src, bytes);
  ++*bar;
}
// try_wait.parity(P) succeeds once the phase with parity P has completed, i.e. when the current phase's parity is not P
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while ((*(volatile unsigned long long*)bar & 1ull) == (unsigned long long)parity) simt::yield();
}
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void timeline_mark(int) {}

}  // namespace se_b200
