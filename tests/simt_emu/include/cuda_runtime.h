// cuda_runtime.h -- a stand-in CUDA runtime for the CPU test tier.  TEST INFRASTRUCTURE ONLY.
//
// tests/simt_emu/build.py compiles the library's own sources (supereight_b200/csrc/*.cu, *.cuh, unmodified except that
// `kernel<<<grid, block, smem, stream>>>(args)` is rewritten to `simt::launch(kernel, grid, block, smem, stream, args)`)
// with g++ against this header, so that the kernels' LOGIC -- indexing, warp-collective protocols, the lock-free tree
// insert, the shared-memory pipeline's hand-shakes -- can be run and compared with the oracle in a container that has no
// GPU.  It is a checker, like the oracle: nothing under supereight_b200/ refers to it, build() does not build it,
// bench.py never loads it, and the product library still fails loudly when there is no CUDA device.
//
// Execution model ("SIMT on fibers"): a kernel launch runs synchronously, CTA after CTA; every CUDA thread of a CTA is a
// fiber (ucontext) on the calling OS thread.  A fiber runs until it finishes, reaches a warp collective (__shfl_sync,
// __ballot_sync, __any_sync, __match_any_sync, __reduce_max_sync, __syncwarp), reaches __syncthreads, or polls (the
// emulated mbarrier wait).  A collective completes when every live lane named in its mask has arrived; the results are then
// computed for all of them at once, as the hardware does.  `__shared__` variables are function-level statics (CTAs run one
// at a time), dynamic shared memory is a per-launch buffer.  Device memory is host memory; streams and events are no-ops
// (events keep a host time stamp).  Not modelled: memory-ordering races between truly concurrent threads, timing.
#pragma once
#include <algorithm>
#include <chrono>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include <sys/mman.h>
#include <ucontext.h>

#ifndef __CUDACC__
#define __CUDACC__ 1          // the library guards its device code with this
#endif
#define SE_SIMT_EMU 1

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types -------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct uchar3 { unsigned char x, y, z; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline double2 make_double2(double x, double y) { return {x, y}; }
inline int2 make_int2(int x, int y) { return {x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return {x, y, z, w}; }
inline uchar3 make_uchar3(unsigned char x, unsigned char y, unsigned char z) { return {x, y, z}; }

// the built-in variables: plain globals, set by the executor whenever a fiber is switched in
inline uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
inline dim3 blockDim, gridDim;

// ---- the SIMT executor ----------------------------------------------------------------------------------------
namespace simt {

enum Op { OP_NONE, OP_SYNCWARP, OP_ANY, OP_ALL, OP_BALLOT, OP_SHFL_IDX, OP_SHFL_DOWN, OP_SHFL_UP, OP_MATCH_ANY, OP_REDUCE_MAX };
enum State { RUNNABLE, WAIT_WARP, WAIT_CTA, DONE };

struct Warp {
  unsigned alive = 0, arrived = 0, mask = 0, gen = 0;
  int op = OP_NONE;
  unsigned long long val[32], res[32];
  int aux[32];
};
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  int lane = 0, warp = 0;
  State state = RUNNABLE;
  unsigned wait_gen = 0;
  bool polling = false;
};
struct Cta {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  int alive = 0, bar_arrived = 0;
  unsigned bar_gen = 0;
  unsigned char* dyn_smem = nullptr;
};

struct Globals {
  Fiber* cur = nullptr;
  Cta* cta = nullptr;
  ucontext_t sched;
  std::function<void()> body;
  char* stacks = nullptr;
  size_t stack_bytes = 256 << 10;
  int max_fibers = 1024;
  long long launches = 0, ctas = 0, switches = 0;
};
inline Globals& g() { static Globals G; return G; }

[[noreturn]] inline void die(const char* what) {
  std::fprintf(stderr, "simt_emu: %s (block %u,%u,%u)\n", what, blockIdx.x, blockIdx.y, blockIdx.z);
  std::abort();
}

inline void switch_to_scheduler() { ++g().switches; swapcontext(&g().cur->ctx, &g().sched); }

// a fiber that polls a condition some other lane will establish (the emulated mbarrier wait)
inline void yield() { g().cur->state = RUNNABLE; g().cur->polling = true; switch_to_scheduler(); }

inline void complete(Warp& w) {
  const unsigned part = w.arrived;
  switch (w.op) {
    case OP_ANY: case OP_ALL: case OP_BALLOT: {
      unsigned b = 0;
      for (int l = 0; l < 32; ++l) if ((part >> l & 1u) && w.val[l]) b |= 1u << l;
      const unsigned long long r = w.op == OP_BALLOT ? b : (w.op == OP_ANY ? (b != 0) : (b == part));
      for (int l = 0; l < 32; ++l) w.res[l] = r;
      break;
    }
    case OP_SHFL_IDX: case OP_SHFL_DOWN: case OP_SHFL_UP:
      for (int l = 0; l < 32; ++l) {
        if (!(part >> l & 1u)) continue;
        int src = w.op == OP_SHFL_IDX ? (w.aux[l] & 31) : (w.op == OP_SHFL_DOWN ? l + w.aux[l] : l - w.aux[l]);
        if (src < 0 || src > 31 || !(part >> src & 1u)) src = l;      // out of range / inactive source: own value
        w.res[l] = w.val[src];
      }
      break;
    case OP_MATCH_ANY:
      for (int l = 0; l < 32; ++l) {
        if (!(part >> l & 1u)) continue;
        unsigned m = 0;
        for (int j = 0; j < 32; ++j) if ((part >> j & 1u) && w.val[j] == w.val[l]) m |= 1u << j;
        w.res[l] = m;
      }
      break;
    case OP_REDUCE_MAX: {
      long long best = INT64_MIN;
      for (int l = 0; l < 32; ++l) if (part >> l & 1u) best = std::max(best, (long long)w.val[l]);
      for (int l = 0; l < 32; ++l) w.res[l] = (unsigned long long)best;
      break;
    }
    default: break;
  }
  w.arrived = 0; w.op = OP_NONE; ++w.gen;
}
inline void try_complete(Warp& w) { if (w.arrived && w.arrived == (w.mask & w.alive)) complete(w); }

inline unsigned long long collective(int op, unsigned mask, unsigned long long v, int aux = 0) {
  Fiber* f = g().cur;
  Warp& w = g().cta->warps[f->warp];
  if (!(mask >> f->lane & 1u)) die("a lane called a collective without naming itself in the mask");
  if (w.arrived == 0) { w.op = op; w.mask = mask; }
  else if (w.op != op || w.mask != mask) die("lanes of one warp reached different collectives (divergent __*_sync)");
  w.val[f->lane] = v; w.aux[f->lane] = aux;
  w.arrived |= 1u << f->lane;
  const unsigned my_gen = w.gen;
  try_complete(w);
  if (w.gen == my_gen) { f->state = WAIT_WARP; f->wait_gen = my_gen; switch_to_scheduler(); }
  return w.res[f->lane];
}

inline void syncthreads() {
  Fiber* f = g().cur;
  Cta& c = *g().cta;
  const unsigned my_gen = c.bar_gen;
  if (++c.bar_arrived == c.alive) { c.bar_arrived = 0; ++c.bar_gen; return; }
  f->state = WAIT_CTA; f->wait_gen = my_gen; switch_to_scheduler();
}

inline void fiber_main() {
  g().body();
  Fiber* f = g().cur;
  Cta& c = *g().cta;
  Warp& w = c.warps[f->warp];
  f->state = DONE;
  w.alive &= ~(1u << f->lane);
  try_complete(w);                                    // the others may have been waiting for this lane only
  --c.alive;
  if (c.alive > 0 && c.bar_arrived == c.alive) { c.bar_arrived = 0; ++c.bar_gen; }
  switch_to_scheduler();
  die("a finished fiber was resumed");
}

inline bool ready(const Cta& c, const Fiber& f) {
  switch (f.state) {
    case RUNNABLE: return true;
    case WAIT_WARP: return c.warps[f.warp].gen != f.wait_gen;
    case WAIT_CTA: return c.bar_gen != f.wait_gen;
    default: return false;
  }
}

inline void run_cta(int nthreads, size_t smem_bytes) {
  Globals& G = g();
  if (nthreads > G.max_fibers) die("block larger than 1024 threads");
  if (!G.stacks) {
    G.stacks = (char*)mmap(nullptr, G.stack_bytes * G.max_fibers, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (G.stacks == (char*)MAP_FAILED) die("mmap of the fiber stacks failed");
  }
  Cta c;
  c.fibers.resize(nthreads);
  c.warps.resize((nthreads + 31) / 32);
  c.alive = nthreads;
  std::vector<unsigned char> smem(smem_bytes + 128);
  c.dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
  G.cta = &c;
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = c.fibers[t];
    f.tid.x = t % blockDim.x; f.tid.y = (t / blockDim.x) % blockDim.y; f.tid.z = t / (blockDim.x * blockDim.y);
    f.lane = t & 31; f.warp = t >> 5;
    c.warps[f.warp].alive |= 1u << f.lane;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = G.stacks + (size_t)t * G.stack_bytes;
    f.ctx.uc_stack.ss_size = G.stack_bytes;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())fiber_main, 0);
  }
  // Warp after warp; inside a warp, round-robin over the lanes that can run, until every lane has finished, waits for
  // the CTA barrier, or only polls (then the other warps get their turn).
  int idle_passes = 0;
  while (c.alive > 0) {
    bool progress = false;
    for (size_t wi = 0; wi < c.warps.size(); ++wi) {
      for (;;) {
        bool worked = false;
        for (int l = 0; l < 32 && (int)wi * 32 + l < nthreads; ++l) {
          Fiber& f = c.fibers[wi * 32 + l];
          if (!ready(c, f)) continue;
          f.state = RUNNABLE; f.polling = false;
          G.cur = &f;
          threadIdx = f.tid;
          swapcontext(&G.sched, &f.ctx);
          if (!(f.state == RUNNABLE && f.polling)) worked = true;      // anything but "polled again"
        }
        if (!worked) break;
        progress = true;
      }
    }
    if (progress) idle_passes = 0;
    else if (++idle_passes > 1000) die("deadlock: no thread of the block can make progress");
  }
  G.cta = nullptr; G.cur = nullptr;
}

template <class... KArgs, class... Args>
void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* /*stream*/, Args&&... args) {
  Globals& G = g();
  if (G.cta) die("nested launch");
  std::tuple<std::decay_t<KArgs>...> bound(static_cast<KArgs>(args)...);
  G.body = [kernel, bound]() mutable { std::apply(kernel, bound); };
  blockDim = block; gridDim = grid;
  ++G.launches;
  const int nthreads = (int)(block.x * block.y * block.z);
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        blockIdx = uint3{x, y, z};
        ++G.ctas;
        run_cta(nthreads, smem);
      }
}

inline unsigned char* dynamic_smem() { return g().cta->dyn_smem; }

template <class T> inline unsigned long long to_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(unsigned long long b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace simt


// ---- warp / block primitives -----------------------------------------------------------------------------------
inline void __syncthreads() { simt::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::collective(simt::OP_SYNCWARP, mask, 0); }
inline void __threadfence() {}
inline int __any_sync(unsigned mask, int p) { return (int)simt::collective(simt::OP_ANY, mask, p != 0); }
inline int __all_sync(unsigned mask, int p) { return (int)simt::collective(simt::OP_ALL, mask, p != 0); }
inline unsigned __ballot_sync(unsigned mask, int p) { return (unsigned)simt::collective(simt::OP_BALLOT, mask, p != 0); }
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int = 32) { return simt::from_bits<T>(simt::collective(simt::OP_SHFL_IDX, mask, simt::to_bits(v), src)); }
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int = 32) { return simt::from_bits<T>(simt::collective(simt::OP_SHFL_DOWN, mask, simt::to_bits(v), (int)d)); }
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int = 32) { return simt::from_bits<T>(simt::collective(simt::OP_SHFL_UP, mask, simt::to_bits(v), (int)d)); }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) { return (unsigned)simt::collective(simt::OP_MATCH_ANY, mask, simt::to_bits(v)); }
inline int __reduce_max_sync(unsigned mask, int v) { return (int)(long long)simt::collective(simt::OP_REDUCE_MAX, mask, (unsigned long long)(long long)v); }

// ---- memory and bit intrinsics -----------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldca(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
// a monotonic counter stands in for the SM clock (k_raycast's cost records only steer the launch order)
inline long long clock64() { static long long t = 0; return t += 64; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
// float -> int, round down, saturating, NaN -> 0 (cvt.rmi.s32.f32)
inline int __float2int_rd(float f) {
  if (f != f) return 0;
  const float r = std::floor(f);
  if (r >= 2147483648.f) return INT32_MAX;
  if (r < -2147483648.f) return INT32_MIN;
  return (int)r;
}
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
// add.rm.f32 (round toward minus infinity)
inline float __fadd_rd(float a, float b) {
  const int old = std::fegetround();
  std::fesetround(FE_DOWNWARD);
  volatile float va = a, vb = b;
  volatile float r = va + vb;
  std::fesetround(old);
  return r;
}

// atomics: one OS thread, fibers switch only at collectives -> plain read-modify-write
template <class T> inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T* p, T v) { const T o = *p; *p = o - v; return o; }
template <class T> inline T atomicOr(T* p, T v) { const T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { const T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T* p, T v) { const T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { const T o = *p; if (o == c) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { const T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { const T o = *p; if (v > o) *p = v; return o; }
inline int atomicAdd(int* p, unsigned v) { return atomicAdd<int>(p, (int)v); }
inline unsigned atomicOr(unsigned* p, int v) { return atomicOr<unsigned>(p, (unsigned)v); }
inline int atomicOr(int* p, unsigned v) { return atomicOr<int>(p, (int)v); }

// CUDA's global min / max overloads
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
using std::isfinite;
using std::isnan;
using std::isinf;

// ---- runtime API -----------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorHostMemoryAlreadyRegistered = 712 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
typedef struct CUstream_st* cudaStream_t;
struct CUevent_st { std::chrono::steady_clock::time_point t; };
typedef CUevent_st* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 6 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes = 0; cudaStream_t stream = nullptr; cudaLaunchAttribute* attrs = nullptr; unsigned numAttrs = 0; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "simt_emu error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 4; std::strcpy(p->name, "simt_emu"); return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) {
  void* q = nullptr;
  if (posix_memalign(&q, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
  std::memset(q, 0xA5, n);                           // fresh device memory is not zero
  *p = (T*)q; return cudaSuccess;
}
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
constexpr unsigned cudaHostAllocMapped = 2, cudaHostRegisterDefault = 0, cudaHostRegisterMapped = 2;
template <class T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMallocHost(p, n); }
inline cudaError_t cudaHostGetDevicePointer(void** dev, void* host, unsigned) { *dev = host; return cudaSuccess; }
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* s, size_t n) { std::memcpy(&sym, s, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMemcpyToSymbolAsync(T& sym, const void* s, size_t n, size_t off, cudaMemcpyKind, cudaStream_t = nullptr) {
  std::memcpy((char*)&sym + off, s, n); return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new CUevent_st{std::chrono::steady_clock::now()}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
// every pointer is pageable host memory -- unless SIMT_HOST_IS_PINNED=1, which makes every pointer look page-locked and mapped
// (alias = the pointer itself), so that the zero-copy paths of the library can be exercised on the fiber executor
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  const char* e = std::getenv("SIMT_HOST_IS_PINNED");
  const bool pinned = e && e[0] == '1';
  a->type = pinned ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered; a->device = 0;
  a->devicePointer = pinned ? (void*)p : nullptr; a->hostPointer = (void*)p; return cudaSuccess;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
template <class... KArgs, class... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args&&... args) {
  simt::launch(kernel, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, (void*)cfg->stream, std::forward<Args>(args)...);
  return cudaSuccess;
}
