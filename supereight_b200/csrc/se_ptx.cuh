// se_ptx.cuh -- every line of inline PTX in the library: programmatic dependent launch, the MUFU
// approximations behind the check-free division / square root, Blackwell's packed fp32 arithmetic,
// and the TMA bulk copy with its mbarrier.  The kernels (se_kernels.cuh, se_tracking.cuh) contain no
// `asm` of their own.
#pragma once
#include <cuda_runtime.h>

namespace se_b200 {
#ifdef __CUDACC__

// Programmatic dependent launch (sm_90+): the per-frame kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's CTAs may be scheduled while its predecessor in the
// stream is still draining.  Every such kernel starts with this: let the NEXT kernel start launching as soon as all of
// this grid's CTAs are resident, then block until the PREVIOUS grid has completed and its writes are visible.  Nothing is
// read or written before the wait, so the stream's ordering semantics are unchanged; launched without the attribute both
// instructions are no-ops.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- inter-warp hand-off through global memory (the integrate kernels' in-kernel active list) --------
// ld_relaxed: a coherent (L2) load for polling.  NOT ld.acquire: ptxas implements an acquire at gpu scope with CCTL.IVALL --
// an invalidation of the SM's whole L1 -- and a polling loop made of those wiped the depth image out of L1 twenty thousand
// times per launch (ncu, round 2: 20 % of the integrate kernel's stall samples sat on CCTL).  What a poll guards is read
// through L2 as well (block_coord via __ldcg, the payload by the TMA engine), so no acquire is needed.
__device__ __forceinline__ int ld_relaxed(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void poll_backoff() { asm volatile("nanosleep.u32 64;" ::: "memory"); }

// ---- predicated read-only load: `pred ? __ldg(p) : otherwise` as one predicated LDG, never a branch ----
__device__ __forceinline__ int ldg_if(bool pred, const int* p, int otherwise) {
  int v;
  asm("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n mov.s32 %0, %3;\n @q ld.global.nc.s32 %0, [%1];\n}" : "=r"(v) : "l"(p), "r"((int)pred), "r"(otherwise));
  return v;
}

// ---- MUFU approximations (the seeds of rcp_rn / sqrt_rn in se_kernels.cuh) ------------------------
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- packed fp32 (add/mul/fma.rn.f32x2: one instruction, two IEEE operations) ---------------------
__device__ __forceinline__ unsigned long long pk(float2 a) { return (unsigned long long)__float_as_uint(a.x) | ((unsigned long long)__float_as_uint(a.y) << 32); }
__device__ __forceinline__ float2 upk(unsigned long long v) { return make_float2(__uint_as_float((unsigned)v), __uint_as_float((unsigned)(v >> 32))); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
// (round toward zero: x + 2^23 then has trunc(x) in its low mantissa bits, 0 <= x < 2^22)
__device__ __forceinline__ float2 add2_rz(float2 a, float2 b) { unsigned long long d; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return upk(d); }

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a): global -> shared, completion counted in bytes ------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
// 16 bytes of shared memory by their 32-bit shared-window address (kept in one register: the generic-pointer form had its
// address re-derived from the CTA's window base in every slice of the integrate loop)
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// makes the mbarrier initialisation visible to the async proxy before the first bulk copy
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- SE_TIMELINE builds only (scripts/timeline.py): nanosecond time stamps of kernel phases, per CTA ----
#ifdef SE_TIMELINE
__device__ unsigned long long g_timeline[16 * 4096];
__device__ __forceinline__ void timeline_mark(int slot) {
  if (threadIdx.x == 0 && blockIdx.x < 4096) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_timeline[16 * blockIdx.x + slot] = t; }
}
#else
__device__ __forceinline__ void timeline_mark(int) {}
#endif

#endif  // __CUDACC__
}  // namespace se_b200
