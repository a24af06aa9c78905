// se_b200.cu -- implementation of the C ABI in include/se_b200.h (one translation unit, like
// the reference's unity build of DenseSLAMSystem.cpp:44-50).
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        -Xcompiler -fPIC,-ffp-contract=off -shared -o libse_b200.so se_b200.cu
// -fmad=false / -ffp-contract=off are part of the arithmetic contract (se_math.cuh).
#include "../../include/se_b200.h"
#include "se_kernels.cuh"
#include "se_tracking.cuh"
#include "se_meshing.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

using namespace se_b200;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) { g_error = msg; return code; }

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      return fail(SE_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
    }                                                                                           \
  } while (0)

struct Pools {
  int* node_child = nullptr;
  unsigned long long* node_code = nullptr;
  unsigned int* node_side = nullptr;
  unsigned int* node_mask = nullptr;
  void* node_value = nullptr;
  unsigned long long* block_code = nullptr;
  int4* block_coord = nullptr;
  int* block_active = nullptr;
  void* block_data = nullptr;
  int* counters = nullptr;
  int* dir = nullptr;
  int* ndir = nullptr;
  unsigned char* cmask = nullptr;
};

}  // namespace

struct se_b200_map {
  int field = 0, size = 0, W = 0, H = 0, device = 0;
  float dim = 0.f;
  int max_level = 0, leaves_level = 0;
  int max_nodes = 0, max_blocks = 0, max_requests = 0;
  int num_sms = 148;
  size_t voxel_bytes = 8;
  Pools p;
  float* d_depth = nullptr;
  float* d_vertex = nullptr;
  float* d_normal = nullptr;
  uchar4* d_rgba = nullptr;
  unsigned short* d_depth_mm = nullptr;
  size_t depth_mm_capacity = 0;
  // The millimetre image of the last preprocess call that nobody has converted yet (mm2meters is deferred to the
  // allocation kernel, which reads every pixel exactly once anyway; resolve_depth() converts on demand for the other
  // consumers of the float image).  pending_slot: the async-upload staging buffer it lives in, or -1.
  const unsigned short* pending_mm = nullptr;
  int pending_inW = 0, pending_ratio = 1, pending_slot = -1;
  cudaEvent_t pending_upload = nullptr;   // the async upload that fills it (the converting stream waits for it), or nullptr
  int* d_active_list = nullptr;
  int* d_miss = nullptr;                  // SDF: directory cells of the blocks the allocation pass found missing (MissList)
  int miss_capacity = 0;
  unsigned long long* d_requests = nullptr;
  float* d_logodds = nullptr;             // OFusion: log-odds increment per (bspline slot(t), slot(t-3)) pair (k_fill_logodds)
  int* d_track = nullptr;
  size_t track_capacity = 0;
  // N1 (tracking front-end): depth pyramid, per-level vertex/normal maps, TrackData, reduction scratch
  int levels = 0;
  float* d_scaled_depth[8] = {};
  float* d_in_vertex[8] = {};
  float* d_in_normal[8] = {};
  TrackData* d_trackdata = nullptr;
  float* d_partial = nullptr;
  float* d_reduction = nullptr;
  float* h_reduction = nullptr;          // pinned, 32 floats + the pose (16)
  IcpState* d_icp = nullptr;
  // N4 (meshing)
  int8_t* d_mc_table = nullptr;
  float* d_mesh = nullptr;               // 9 floats per triangle
  long long mesh_triangles = 0, mesh_capacity = 0;
  int* h_counters = nullptr;              // pinned
  int* h_status = nullptr;                // pinned + mapped: [0] = the allocation pass's error bits, written by the integrate kernel
  int* d_status = nullptr;                // its device alias
  cudaStream_t stream = nullptr, own_stream = nullptr;
  // Cross-frame overlap (SDF): the allocation pass of frame f+1 only reads the block directory and writes things the ray
  // kernels never look at (float depth, active flags, the miss list) -- blocks are created later, inside the integrate
  // kernel -- so it runs on a stream of its own, behind frame f's integrate kernel but BESIDE frame f's raycast / render.
  cudaStream_t alloc_stream = nullptr;
  cudaEvent_t ev_integrated = nullptr, ev_alloc_done = nullptr, ev_depth_ready = nullptr, ev_depth_read = nullptr;
  bool integrated_valid = false, depth_ready_valid = false, depth_read_valid = false;   // integrated_valid: ev_integrated marks the LATEST integrate kernel
  bool overlap_alloc = true;
  // overlapped host I/O (se_b200_*_host_async): the copy engines run on their own streams, double-buffered in HBM, tied
  // to the kernel stream by events, so frame N+1's upload and frame N's download overlap the kernels
  struct AsyncIo {
    cudaStream_t up = nullptr, down = nullptr;
    unsigned short* d_mm[2] = {nullptr, nullptr};
    size_t mm_capacity[2] = {0, 0};
    uchar4* d_out[2] = {nullptr, nullptr};
    cudaEvent_t uploaded[2] = {}, consumed[2] = {}, rendered[2] = {}, downloaded[2] = {};
    bool consumed_valid[2] = {false, false}, downloaded_valid[2] = {false, false};
    int up_slot = 0, down_slot = 0;
    bool ready = false;
  } aio;
  cudaEvent_t ev_begin[SE_B200_NUM_STAGES] = {}, ev_end[SE_B200_NUM_STAGES] = {};
  bool ev_valid[SE_B200_NUM_STAGES] = {};
  // se_b200_set_render_target: the raycast also shades into rt_dev (device memory or the mapped alias of the caller's
  // pinned buffer rt_user); rt_valid = the last raycast filled it, with the light at rt_light
  uchar4* rt_dev = nullptr;
  const void* rt_user = nullptr;
  bool rt_valid = false;
  float rt_light[3] = {0.f, 0.f, 0.f};
  long long launches = 0;
  int grid_integrate = 0;
  int* d_ray_sched = nullptr;             // k_raycast's launch orders [2][groups] and costs [2][groups] (LaunchSchedule)
  unsigned ray_launches = 0;
  int* d_alloc_sched = nullptr;           // ... and the allocation kernels'
  unsigned alloc_launches = 0;
  int parity = 0;
  bool stage_timing = false;            // per-stage event pairs (se_b200_set_stage_timing): off by default -- an event between two launches breaks their programmatic-dependent-launch edge

  template <class V> MapView<V> view() const {
    MapView<V> v;
    v.size = size; v.dim = dim; v.max_level = max_level; v.leaves_level = leaves_level;
    v.inv_voxel = (float)size / dim; v.grad_scale = (0.5f * dim) / (float)size;
    v.max_nodes = max_nodes; v.max_blocks = max_blocks;
    v.node_child = p.node_child; v.node_code = p.node_code; v.node_side = p.node_side; v.node_mask = p.node_mask;
    v.node_value = (V*)p.node_value;
    v.block_code = p.block_code; v.block_coord = p.block_coord; v.block_active = p.block_active;
    v.block_data = (V*)p.block_data;
    v.counters = p.counters;
    v.dir = p.dir; v.dir_dim = size / kBlockSide;
    v.ndir = p.ndir;
    v.cmask = p.cmask;
    return v;
  }
};

namespace {

int ensure_tracking_buffers(se_b200_map* m, int levels);     // N1, defined with the tracking entry points below

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

M4 to_m4(const float* p) { M4 m; std::memcpy(m.m, p, sizeof(m.m)); return m; }

// Launch on the map's stream with programmatic stream serialisation (see pdl_prologue in se_map.cuh).
// SE_B200_NO_PDL=1 falls back to plain stream order (A/B measurements).
bool pdl_enabled() {
  static const bool on = [] { const char* e = std::getenv("SE_B200_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}
template <class... KArgs, class... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface in check_launch (cudaGetLastError)
}

// Device-visible alias of a page-locked host buffer (under unified addressing every cudaHostAlloc / cudaHostRegister'd
// range is mapped), or nullptr for pageable memory.  The `_host` entry points use it to let the last kernel write the
// caller's buffer directly over PCIe instead of staging in HBM and issuing a separate copy.  SE_B200_NO_ZEROCOPY=1 turns
// it off (A/B measurements).  (Reading a pinned INPUT in place was tried and dropped: no measurable gain over the DMA copy.)
void* mapped_alias(const void* host) {
  static const bool off = [] { const char* e = std::getenv("SE_B200_NO_ZEROCOPY"); return e && e[0] == '1'; }();
  if (off) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return a.devicePointer;
}
int check_launch(se_b200_map* m, int n = 1) {
  m->launches += n;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SE_B200_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  return SE_B200_OK;
}

void stage_begin(se_b200_map* m, int s) { if (m->stage_timing) cudaEventRecord(m->ev_begin[s], m->stream); }
void stage_end(se_b200_map* m, int s) { if (m->stage_timing) { cudaEventRecord(m->ev_end[s], m->stream); m->ev_valid[s] = true; } }

// B-spline table of bfusion/bspline_lookup.cc:36-37, regenerated from the closed form it samples
// (mapping_impl.hpp:94-106): evaluated in double at t = -3 + 6 i / 999 and rounded to float it
// reproduces the reference's printed table bit for bit (tests/test_bspline_lut.py pins the checksum).
void make_bspline_lut(float lut[1000]) {
  for (int i = 0; i < 1000; ++i) {
    const double t = -3.0 + 6.0 * (double)i / 999.0;
    double value = 0.0;
    if (t >= -3.0 && t <= -1.0) value = std::pow(3 + t, 3) / 48.0;
    else if (t > -1 && t <= 1) value = 0.5 + (t * (3 + t) * (3 - t)) / 24.0;
    else if (t > 1 && t <= 3) value = 1 - std::pow(3 - t, 3) / 48.0;
    else if (t > 3) value = 1.0;
    lut[i] = (float)value;
  }
}

// 0, or finite with magnitude in [2^-20, 2^20]
bool normal_range(float v) { const float a = std::fabs(v); return v == 0.f || (a >= 0x1p-20f && a <= 0x1p20f); }

// the schedule of this launch of a per-pixel kernel (LaunchSchedule in se_kernels.cuh): `sched` holds orders [2][groups] and costs [2][groups]
LaunchSchedule next_schedule(int* sched, int groups, unsigned& launches) {
  LaunchSchedule ls{nullptr, nullptr, nullptr, nullptr};
  if (!sched) return ls;
  const int cur = (int)(launches & 1u), nxt = cur ^ 1;
  ls.order = sched + (size_t)cur * groups; ls.order_next = sched + (size_t)nxt * groups;
  ls.cost = sched + (size_t)(2 + cur) * groups; ls.cost_prev = sched + (size_t)(2 + nxt) * groups;
  ++launches;
  return ls;
}
int* make_schedule(int groups) {
  std::vector<int> init((size_t)4 * groups, 0);
  for (int j = 0; j < 2; ++j) std::iota(init.begin() + (size_t)j * groups, init.begin() + (size_t)(j + 1) * groups, 0);
  int* d = nullptr;
  if (cudaMalloc(&d, init.size() * sizeof(int)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, init.data(), init.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
  return d;
}

// CTAs of the per-pixel ray kernels (tile_pixel in se_kernels.cuh: 8x4 pixel tiles per warp)
int pixel_tile_blocks(int W, int H, int threads) {
  const int tiles = ((W + 7) / 8) * ((H + 3) / 4);
  const int warps_per_block = threads / 32;
  return (tiles + warps_per_block - 1) / warps_per_block;
}

template <class V>
int create_pools(se_b200_map* m) {
  const size_t nn = (size_t)m->max_nodes, nb = (size_t)m->max_blocks;
  CUDA_TRY(cudaMalloc(&m->p.node_child, nn * 8 * sizeof(int)));
  CUDA_TRY(cudaMalloc(&m->p.node_code, nn * sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&m->p.node_side, nn * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc(&m->p.node_mask, nn * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc(&m->p.node_value, nn * 8 * sizeof(V)));
  CUDA_TRY(cudaMalloc(&m->p.block_code, nb * sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&m->p.block_coord, nb * sizeof(int4)));
  CUDA_TRY(cudaMalloc(&m->p.block_active, nb * sizeof(int)));
  CUDA_TRY(cudaMalloc(&m->p.block_data, (nb + 1) * kBlockVoxels * sizeof(V)));      // + the never-allocated initValue() payload (se_map.cuh)
  CUDA_TRY(cudaMalloc(&m->p.counters, kNumCounters * sizeof(int)));
  // block directory: (size/8)^3 ints; skipped above 8192^3 (4 GiB) or when SE_B200_DISABLE_DIRECTORY is set
  // (the tree descent is then used everywhere; tests run both ways)
  const size_t g = (size_t)m->size / kBlockSide;
  if (m->size <= 8192 && !getenv("SE_B200_DISABLE_DIRECTORY")) {
    CUDA_TRY(cudaMalloc(&m->p.dir, g * g * g * sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(m->p.dir, 0xFF, g * g * g * sizeof(int), m->stream));          // kEmpty
    // node directory for levels 1 .. leaves_level-1: (8^leaves - 8) / 7 cells
    const size_t ncells = (size_t)(((1ll << (3 * m->leaves_level)) - 8) / 7);
    if (ncells > 0) {
      CUDA_TRY(cudaMalloc(&m->p.ndir, ncells * sizeof(int)));
      CUDA_TRY(cudaMemsetAsync(m->p.ndir, 0xFF, ncells * sizeof(int), m->stream));
    }
    // children masks by heap index for the levels 0 .. leaves_level-1 (MapView::cmask): 37 KB at 512^3, 2.4 MB at 2048^3,
    // 153 MB at 8192^3 (heap indices stay below 2^31 up to there)
    if (m->leaves_level <= 10) {
      const size_t nmask = ((size_t)heap_level_offset(m->leaves_level) + 3) & ~(size_t)3;
      CUDA_TRY(cudaMalloc(&m->p.cmask, nmask));
      CUDA_TRY(cudaMemsetAsync(m->p.cmask, 0, nmask, m->stream));
    }
  }
  CUDA_TRY(cudaMemsetAsync(m->p.node_child, 0xFF, nn * 8 * sizeof(int), m->stream));      // kEmpty
  CUDA_TRY(cudaMemsetAsync(m->p.node_code, 0, nn * sizeof(unsigned long long), m->stream));
  CUDA_TRY(cudaMemsetAsync(m->p.node_side, 0, nn * sizeof(unsigned), m->stream));
  CUDA_TRY(cudaMemsetAsync(m->p.node_mask, 0, nn * sizeof(unsigned), m->stream));
  CUDA_TRY(cudaMemsetAsync(m->p.block_code, 0, nb * sizeof(unsigned long long), m->stream));
  CUDA_TRY(cudaMemsetAsync(m->p.block_coord, 0, nb * sizeof(int4), m->stream));
  CUDA_TRY(cudaMemsetAsync(m->p.block_active, 0, nb * sizeof(int), m->stream));
  CUDA_TRY(cudaMemsetAsync(m->p.counters, 0, kNumCounters * sizeof(int), m->stream));
  const int grid = m->num_sms * 8;
  k_fill_voxels<V><<<grid, 256, 0, m->stream>>>((V*)m->p.node_value, nn * 8);
  k_fill_voxels<V><<<grid, 256, 0, m->stream>>>((V*)m->p.block_data, (nb + 1) * kBlockVoxels);
  if (int r = check_launch(m, 2)) return r;
  // root: Octree::init (octree.hpp:411-421): node 0, code 0, side = size
  const int one = 1;
  const unsigned side = (unsigned)m->size;
  CUDA_TRY(cudaMemcpyAsync(m->p.counters + kCntNodes, &one, sizeof(int), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaMemcpyAsync(m->p.counters + kCntLastNodes, &one, sizeof(int), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaMemcpyAsync(m->p.node_side, &side, sizeof(unsigned), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int fetch_counters(se_b200_map* m) {
  CUDA_TRY(cudaMemcpyAsync(m->h_counters, m->p.counters, kCntTake * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int check_pool_error(se_b200_map* m);

// The allocation pass cannot fail synchronously (a pool runs out on the device, in the middle of a frame): find_or_create
// sets an error bit, the frame's integrate kernel copies the bits to a word of mapped page-locked memory, and the next
// stage call that sees them reports SE_B200_ERR_POOL -- without a device synchronisation on the per-frame path.  Reporting
// clears the bits (device and host side); the octants that did not fit are lost for that frame and requested again by
// the next one.
int report_pool_error(se_b200_map* m) {
  const int err = *(volatile int*)m->h_status;
  if (!err) return SE_B200_OK;
  m->h_status[0] = 0;
  CUDA_TRY(cudaMemsetAsync(m->p.counters + kCntError, 0, sizeof(int), m->stream));
  m->h_counters[kCntError] = err;
  const int r = check_pool_error(m);
  m->h_counters[kCntError] = 0;
  return r;
}

// the float depth image, for its consumers other than the allocation kernels (which convert the pending millimetre
// image themselves)
int resolve_depth(se_b200_map* m) {
  if (!m->pending_mm) return SE_B200_OK;
  if (m->pending_upload) cudaStreamWaitEvent(m->stream, m->pending_upload, 0);
  dim3 block(32, 8), grid((m->W + 31) / 32, (m->H + 7) / 8);
  launch_pdl(k_mm2meters, grid, block, 0, m->stream, m->d_depth, m->pending_mm, m->W, m->H, m->pending_inW, m->pending_ratio);
  if (m->pending_slot >= 0) { cudaEventRecord(m->aio.consumed[m->pending_slot], m->stream); m->aio.consumed_valid[m->pending_slot] = true; }
  m->pending_mm = nullptr; m->pending_slot = -1; m->pending_upload = nullptr;
  return check_launch(m);
}

template <class V>
int integrate_impl(se_b200_map* m, const float* pose_p, const float* k, float mu, unsigned frame) {
  if (int r = report_pool_error(m)) return r;
  const M4 pose = to_m4(pose_p);
  const float voxelsize = m->dim / (float)m->size;                       // DenseSLAMSystem.cpp:211
  const float band = FieldTraits<V>::is_sdf ? 2 * mu : 6 * mu;           // :223, :228
  MapView<V> view = m->view<V>();

  AllocParams ap;
  ap.kPose = mul44(pose, inverse_camera_matrix(k));
  ap.camera = v3(pose.m[3], pose.m[7], pose.m[11]);
  ap.inverseVoxelSize = 1 / voxelsize;
  ap.voxelSize = voxelsize;
  ap.band = band;
  ap.numSteps = (int)std::ceil(band * ap.inverseVoxelSize);
  ap.W = m->W; ap.H = m->H;
  bool alloc_fast = !getenv("SE_B200_IEEE_DIV") && normal_range(band) && band > 0.f && ap.numSteps >= 1;
  for (int i = 0; i < 12 && alloc_fast; ++i) alloc_fast = normal_range(ap.kPose.m[i]);
  ap.fast = alloc_fast ? 1 : 0;
  MissList miss;
  miss.cells = nullptr; miss.capacity = 0;
  if (FieldTraits<V>::is_sdf) {
    // The list of blocks the allocation pass finds missing (with duplicates: every warp reports the distinct ones of its
    // rays per round).  A ray enters at most numSteps * sqrt(3) / 8 + 4 blocks, so this capacity cannot overflow.
    const size_t need = (size_t)m->W * m->H * ((size_t)(ap.numSteps > 0 ? ap.numSteps : 0) * 2 / 8 + 4);
    if ((size_t)m->miss_capacity < need) {
      if (need > ((size_t)1 << 30)) return fail(SE_B200_ERR_ARG, "mu / voxel size too large for the allocation request list");
      CUDA_TRY(cudaStreamSynchronize(m->stream));
      cudaFree(m->d_miss); m->d_miss = nullptr; m->miss_capacity = 0;
      CUDA_TRY(cudaMalloc(&m->d_miss, need * sizeof(int)));
      m->miss_capacity = (int)need;
    }
    miss.cells = m->d_miss; miss.capacity = m->miss_capacity;
  }

  // counters: remember the pool sizes before the frame, clear the per-frame ones
  stage_begin(m, SE_B200_STAGE_ALLOC);
  m->parity ^= 1;
  const int parity = m->parity;
  const int threads = 256;
  const int grid_px = pixel_tile_blocks(m->W, m->H, threads);
  // a1 rides along: a pending millimetre image is converted by the allocation kernel's threads (one pixel each)
  DepthSource src;
  src.mm = m->pending_mm; src.inW = m->pending_inW; src.ratio = m->pending_ratio;
  // SDF: the allocation kernel goes to the map's second stream, ordered behind the previous frame's integrate kernel (float
  // depth, active flags, pools) and whatever still reads the float depth, but NOT behind that frame's raycast / render: it
  // reads the block directory and creates nothing (se_b200_map::alloc_stream).  Per-stage timing keeps one stream.
  // ... and not when the depth image itself arrives at the end of the main stream (se_b200_preprocess_depth_host's copy: the
  // synchronous per-frame loop, where there is nothing to overlap with): stream order then does it all, and the launches
  // keep their programmatic-dependent-launch edges (an event between two kernels breaks the edge).
  const bool depth_on_main = m->pending_mm && m->depth_ready_valid;
  const bool overlap = FieldTraits<V>::is_sdf && m->overlap_alloc && !m->stage_timing && !depth_on_main;
  cudaStream_t as = overlap ? m->alloc_stream : m->stream;
  if (overlap) {
    if (!m->integrated_valid) { CUDA_TRY(cudaEventRecord(m->ev_integrated, m->stream)); m->integrated_valid = true; }   // (first frame / after a frame without overlap: everything so far)
    CUDA_TRY(cudaStreamWaitEvent(as, m->ev_integrated, 0));
    if (m->depth_read_valid) CUDA_TRY(cudaStreamWaitEvent(as, m->ev_depth_read, 0));
  }
  if (m->pending_mm && m->pending_upload) CUDA_TRY(cudaStreamWaitEvent(as, m->pending_upload, 0));
  m->depth_read_valid = false; m->depth_ready_valid = false;
  if (FieldTraits<V>::is_sdf) {
    launch_pdl(k_alloc_sdf<V>, grid_px, threads, 0, as, view, m->d_depth, src, ap, miss, parity);
    if (int r = check_launch(m)) return r;
  } else {
    // (expensive tile groups first: these rays are of very different lengths -- 87 -> 63 us at 1024^3)
    launch_pdl(k_alloc_ofusion<V>, grid_px, threads, 0, m->stream, view, m->d_depth, src, ap, m->d_requests, m->max_requests, next_schedule(m->d_alloc_sched, grid_px, m->alloc_launches));
    launch_pdl(k_alloc_first_key_chain<V>, 1, 1024, 0, m->stream, view, m->d_requests, m->max_requests);
    if (int r = check_launch(m, 2)) return r;
  }
  if (m->pending_mm) {
    if (m->pending_slot >= 0) { CUDA_TRY(cudaEventRecord(m->aio.consumed[m->pending_slot], as)); m->aio.consumed_valid[m->pending_slot] = true; }
    m->pending_mm = nullptr; m->pending_slot = -1; m->pending_upload = nullptr;
  }
  const M4 Tcw = rigid_inverse(pose);
  const M4 K = camera_matrix(k);
  FrustumParams fp;
  fp.cam = mul44(K, Tcw);
  fp.voxelSize = voxelsize; fp.W = m->W; fp.H = m->H;
  // a8 for the blocks that existed before the frame: behind the allocation kernel on its stream (SDF; see k_filter_blocks)
  const int prefiltered = FieldTraits<V>::is_sdf ? 1 : 0;
  if (prefiltered) {
    launch_pdl(k_filter_blocks<V>, m->num_sms * 2, kListThreads, 0, as, view, fp, m->d_active_list, parity);
    if (int r = check_launch(m)) return r;
  }
  if (overlap) {
    CUDA_TRY(cudaEventRecord(m->ev_alloc_done, as));
    CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_alloc_done, 0));
  }
  stage_end(m, SE_B200_STAGE_ALLOC);

  stage_begin(m, SE_B200_STAGE_FUSE);
  IntegrateParams ip;
  ip.Tcw = Tcw; ip.K = K;
  ip.delta = rot3(Tcw, v3(voxelsize, 0.f, 0.f));
  ip.cameraDelta = rot3(K, ip.delta);
  ip.voxelSize = voxelsize; ip.mu = mu;
  ip.timestamp = (1.f / 30.f) * (float)frame;                            // DenseSLAMSystem.cpp:243
  ip.W = m->W; ip.H = m->H;
  ip.rmu = 1.f / mu; ip.wlim = (float)m->W - 1.5f; ip.hlim = (float)m->H - 1.5f; ip.wf = (float)m->W; ip.no_sample = kPixMagicBits + (unsigned)(m->W * m->H);
  ip.one2 = make_float2(1.f, 1.f); ip.mone2 = make_float2(-1.f, -1.f);
  ip.tz2 = make_float2(Tcw.m[2], Tcw.m[6]); ip.tt2 = make_float2(Tcw.m[3], Tcw.m[7]);
  ip.nkd2 = make_float2(-K.m[0], -K.m[5]); ip.nkz2 = make_float2(-K.m[2], -K.m[6]);
  // persistent grid-stride kernels: exactly as many CTAs as are resident at once (one wave), trip
  // counts read from device counters
  if (m->grid_integrate == 0) {
    int occ = 0;
    if (FieldTraits<V>::is_sdf) {
      CUDA_TRY(cudaFuncSetAttribute(k_integrate_sdf<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIntegrateSmem));
      CUDA_TRY(cudaFuncSetAttribute(k_integrate_sdf<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIntegrateSmem));
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_integrate_sdf<true>, kIntegrateWarps * 32, kIntegrateSmem) != cudaSuccess || occ < 1) occ = 2;
    } else {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_integrate_ofusion<true>, threads, 0) != cudaSuccess || occ < 1) occ = 2;
    }
    m->grid_integrate = m->num_sms * occ;
  }
  // the check-free division/sqrt sequences need every operand in the normal float range: guaranteed when
  // the matrices, the voxel size and mu are 0 or within [2^-20, 2^20]; anything else takes the plain IEEE kernel
  // (... and its pixel index is formed in fp32, exact for images of fewer than 2^22 pixels: sdf_voxel_pair)
  bool fast = !getenv("SE_B200_IEEE_DIV") && normal_range(voxelsize) && normal_range(mu) && mu > 0.f && (long long)m->W * m->H < (1ll << 22);
  for (int i = 0; i < 12 && fast; ++i) fast = normal_range(Tcw.m[i]) && normal_range(K.m[i]);
  if (FieldTraits<V>::is_sdf) {
    if (fast) launch_pdl(k_integrate_sdf<true>, m->grid_integrate, kIntegrateWarps * 32, kIntegrateSmem, m->stream, m->view<SdfVoxel>(), m->d_depth, ip, fp, m->d_active_list, miss, parity, m->d_status, prefiltered);
    else launch_pdl(k_integrate_sdf<false>, m->grid_integrate, kIntegrateWarps * 32, kIntegrateSmem, m->stream, m->view<SdfVoxel>(), m->d_depth, ip, fp, m->d_active_list, miss, parity, m->d_status, prefiltered);
  } else {
    // check-free sequences + tabulated log-odds increment: the default (fuse 130 -> 59 us on box_room_ofusion1024,
    // same bits); SE_B200_OFUSION_FAST=0 or SE_B200_IEEE_DIV select the instantiation with the plain operators
    const char* e = std::getenv("SE_B200_OFUSION_FAST");
    const bool ofusion_fast = !(e && e[0] == '0');
    if (fast && ofusion_fast && m->d_logodds) launch_pdl(k_integrate_ofusion<true>, m->grid_integrate, threads, 0, m->stream, m->view<OfuVoxel>(), m->d_depth, ip, fp, m->d_active_list, miss, parity, m->d_status, (const float*)m->d_logodds);
    else launch_pdl(k_integrate_ofusion<false>, m->grid_integrate, threads, 0, m->stream, m->view<OfuVoxel>(), m->d_depth, ip, fp, m->d_active_list, miss, parity, m->d_status, (const float*)m->d_logodds);
  }
  if (int r = check_launch(m, 1)) return r;
  // the next frame's allocation kernel may start from here on -- recorded only while frames do overlap
  if (overlap) CUDA_TRY(cudaEventRecord(m->ev_integrated, m->stream));
  m->integrated_valid = overlap;
  stage_end(m, SE_B200_STAGE_FUSE);
  return SE_B200_OK;
}

RaycastParams make_raycast_params(se_b200_map* m, const float* pose, const float* k, float mu, float farPlane, float largestep, int use_tcmin) {
  RaycastParams rp;
  rp.view = mul44(to_m4(pose), inverse_camera_matrix(k));
  rp.nearPlane = kNearPlane; rp.farPlane = farPlane; rp.mu = mu;
  rp.step = m->dim / (float)m->size;                                      // DenseSLAMSystem.cpp:197, :282
  rp.largestep = largestep;
  rp.W = m->W; rp.H = m->H; rp.use_tcmin = use_tcmin;
  // the ray-independent part of the ray set-up: the same single-rounding IEEE operations RayWalk::init performs per
  // thread (this file is compiled without FMA contraction)
  rp.so = v3(rp.view.m[3] / m->dim + 1.f, rp.view.m[7] / m->dim + 1.f, rp.view.m[11] / m->dim + 1.f);
  rp.eps = 1.0f / (float)m->size;
  rp.near_n = rp.nearPlane / m->dim;
  rp.far_n = rp.farPlane / m->dim;
  // the check-free division / square-root sequences of the ray kernels (normalized3_fast, RayWalk::init_pre) are used
  // when no operand can leave the normal float range; SE_B200_IEEE_DIV=1 forces the plain operators
  bool fast = !getenv("SE_B200_IEEE_DIV");
  for (int i = 0; i < 12 && fast; ++i) fast = normal_range(rp.view.m[i]);
  rp.fast = fast ? 1 : 0;
  return rp;
}

template <class V, bool DENSE>
void launch_raycast(se_b200_map* m, const RaycastParams& rp, unsigned long long* stats_dev, bool shade, V3 light) {
  const int grid = pixel_tile_blocks(m->W, m->H, kRayThreads);
  // the launch order of the tile groups: expensive first, from the costs of earlier launches (LaunchSchedule)
  const LaunchSchedule rs = stats_dev ? LaunchSchedule{nullptr, nullptr, nullptr, nullptr} : next_schedule(m->d_ray_sched, grid, m->ray_launches);
  if (stats_dev) launch_pdl(k_raycast<V, DENSE, true, false>, grid, kRayThreads, 0, m->stream, m->view<V>(), rp, m->d_vertex, m->d_normal, stats_dev, light, (uchar4*)nullptr, rs);
  else if (shade) launch_pdl(k_raycast<V, DENSE, false, true>, grid, kRayThreads, 0, m->stream, m->view<V>(), rp, m->d_vertex, m->d_normal, (unsigned long long*)nullptr, light, m->rt_dev, rs);
  else launch_pdl(k_raycast<V, DENSE, false, false>, grid, kRayThreads, 0, m->stream, m->view<V>(), rp, m->d_vertex, m->d_normal, (unsigned long long*)nullptr, light, (uchar4*)nullptr, rs);
}

template <class V>
int raycast_impl(se_b200_map* m, const float* pose, const float* k, float mu, unsigned long long* stats_dev) {
  if (int r = report_pool_error(m)) return r;
  const float step = m->dim / (float)m->size;
  const RaycastParams rp = make_raycast_params(m, pose, k, mu, kFarPlane, step * (float)kBlockSide, 1);
  stage_begin(m, SE_B200_STAGE_RAYCAST);
  const bool shade = m->rt_dev && !stats_dev;
  const V3 light = v3(pose[3], pose[7], pose[11]);             // the reuse path's light: view pose == raycast pose
  m->rt_valid = false;
  if (m->p.cmask) launch_raycast<V, true>(m, rp, stats_dev, shade, light);
  else launch_raycast<V, false>(m, rp, stats_dev, shade, light);
  if (shade) { m->rt_light[0] = pose[3]; m->rt_light[1] = pose[7]; m->rt_light[2] = pose[11]; m->rt_valid = true; }
  if (int r = check_launch(m)) return r;
  stage_end(m, SE_B200_STAGE_RAYCAST);
  return SE_B200_OK;
}

// true when the last raycast already rendered the reuse-path image of this view into the caller's render target
bool render_target_holds(const se_b200_map* m, const void* out, const float* view_pose, int reraycast) {
  return !reraycast && m->rt_valid && out && (out == m->rt_user || out == (const void*)m->rt_dev) &&
         view_pose[3] == m->rt_light[0] && view_pose[7] == m->rt_light[1] && view_pose[11] == m->rt_light[2];
}

template <class V>
int render_volume_impl(se_b200_map* m, uchar4* out_dev, const float* view_pose, const float* k, float mu, float largestep, int reraycast) {
  const RaycastParams rp = make_raycast_params(m, view_pose, k, mu, kFarPlane * 2.0f, largestep, 0);   // DenseSLAMSystem.cpp:283-288
  const V3 light = v3(view_pose[3], view_pose[7], view_pose[11]);
  stage_begin(m, SE_B200_STAGE_RENDER);
  if (reraycast && m->p.cmask) launch_pdl(k_render_volume<V, true>, pixel_tile_blocks(m->W, m->H, kRayThreads), kRayThreads, 0, m->stream, m->view<V>(), rp, light, 1, m->d_vertex, m->d_normal, out_dev);
  else if (reraycast) launch_pdl(k_render_volume<V, false>, pixel_tile_blocks(m->W, m->H, kRayThreads), kRayThreads, 0, m->stream, m->view<V>(), rp, light, 1, m->d_vertex, m->d_normal, out_dev);
  else launch_pdl(k_render_shade, (m->W * m->H + 255) / 256, 256, 0, m->stream, m->d_vertex, m->d_normal, light, m->W * m->H, out_dev);
  if (int r = check_launch(m)) return r;
  stage_end(m, SE_B200_STAGE_RENDER);
  return SE_B200_OK;
}

template <class V>
int download_blocks_sorted_impl(se_b200_map* m, uint64_t* keys, int32_t* coords, uint8_t* active, void* voxels) {
  if (int r = fetch_counters(m)) return r;
  const int n = std::min(m->h_counters[kCntBlocks], m->max_blocks);
  std::vector<unsigned long long> code(n);
  std::vector<int4> coord(n);
  std::vector<int> act(n);
  CUDA_TRY(cudaMemcpyAsync(code.data(), m->p.block_code, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaMemcpyAsync(coord.data(), m->p.block_coord, (size_t)n * sizeof(int4), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaMemcpyAsync(act.data(), m->p.block_active, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return code[a] < code[b]; });
  std::vector<V> tmp;
  if (voxels) {
    tmp.resize((size_t)n * kBlockVoxels);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), m->p.block_data, (size_t)n * kBlockVoxels * sizeof(V), cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
  }
  for (int i = 0; i < n; ++i) {
    const int s = order[i];
    if (keys) keys[i] = code[s];
    if (coords) { coords[3 * i] = coord[s].x; coords[3 * i + 1] = coord[s].y; coords[3 * i + 2] = coord[s].z; }
    if (active) active[i] = act[s] ? 1 : 0;
    if (voxels) std::memcpy((V*)voxels + (size_t)i * kBlockVoxels, tmp.data() + (size_t)s * kBlockVoxels, sizeof(V) * kBlockVoxels);
  }
  return SE_B200_OK;
}

template <class V>
int download_nodes_sorted_impl(se_b200_map* m, uint64_t* codes, uint32_t* side, uint8_t* mask, void* values) {
  if (int r = fetch_counters(m)) return r;
  const int n = std::min(m->h_counters[kCntNodes], m->max_nodes);
  std::vector<unsigned long long> code(n);
  std::vector<unsigned> sd(n), mk(n);
  std::vector<V> val((size_t)n * 8);
  CUDA_TRY(cudaMemcpyAsync(code.data(), m->p.node_code, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaMemcpyAsync(sd.data(), m->p.node_side, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaMemcpyAsync(mk.data(), m->p.node_mask, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaMemcpyAsync(val.data(), m->p.node_value, (size_t)n * 8 * sizeof(V), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return code[a] < code[b]; });
  for (int i = 0; i < n; ++i) {
    const int s = order[i];
    if (codes) codes[i] = code[s];
    if (side) side[i] = sd[s];
    if (mask) mask[i] = (uint8_t)mk[s];
    if (values) std::memcpy((V*)values + (size_t)i * 8, val.data() + (size_t)s * 8, sizeof(V) * 8);
  }
  return SE_B200_OK;
}

// scratch device buffer helper for the query entry points (not on the hot path)
struct Scratch {
  void* p = nullptr;
  ~Scratch() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
};

int check_pool_error(se_b200_map* m) {
  const int err = m->h_counters[kCntError];
  if (err & kErrBlockPoolFull) return fail(SE_B200_ERR_POOL, "VoxelBlock pool exhausted: raise max_blocks");
  if (err & kErrNodePoolFull) return fail(SE_B200_ERR_POOL, "node pool exhausted: raise max_nodes");
  if (err & kErrKeyListFull) return fail(SE_B200_ERR_POOL, "octant request list exhausted");
  if (err & kErrMissListFull) return fail(SE_B200_ERR_POOL, "allocation request list exhausted (requests of this frame were dropped)");
  return SE_B200_OK;
}

}  // namespace

// ---- N4: meshing ----------------------------------------------------------------------------
template <class V>
int extract_mesh_impl(se_b200_map* m, int64_t* n_triangles) {
  if (int r = fetch_counters(m)) return r;
  const int n = std::min(m->h_counters[kCntBlocks], m->max_blocks);
  m->mesh_triangles = 0;
  if (n_triangles) *n_triangles = 0;
  if (n == 0) return SE_B200_OK;
  if (!m->d_mc_table) {
    int8_t table[256 * kMcRow];
    mc_case_table(table);
    CUDA_TRY(cudaMalloc(&m->d_mc_table, sizeof(table)));
    CUDA_TRY(cudaMemcpyAsync(m->d_mc_table, table, sizeof(table), cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));          // `table` is on the stack
  }
  // block ids in ascending key order (the order a serial run over a sorted block list would visit them)
  Scratch ids_in, ids_out, keys_out, counts, offsets, tmp;
  CUDA_TRY(ids_in.alloc((size_t)n * sizeof(int)));
  CUDA_TRY(ids_out.alloc((size_t)n * sizeof(int)));
  CUDA_TRY(keys_out.alloc((size_t)n * sizeof(unsigned long long)));
  CUDA_TRY(counts.alloc((size_t)(n + 1) * sizeof(unsigned int)));
  CUDA_TRY(offsets.alloc((size_t)(n + 1) * sizeof(unsigned long long)));
  k_iota<<<(n + 255) / 256, 256, 0, m->stream>>>((int*)ids_in.p, n);
  size_t sort_bytes = 0, scan_bytes = 0;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)m->p.block_code, (unsigned long long*)keys_out.p,
                                           (const int*)ids_in.p, (int*)ids_out.p, n, 0, 64, m->stream));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const unsigned int*)counts.p, (unsigned long long*)offsets.p, n + 1, m->stream));
  CUDA_TRY(tmp.alloc(std::max(sort_bytes, scan_bytes)));
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, sort_bytes, (const unsigned long long*)m->p.block_code, (unsigned long long*)keys_out.p,
                                           (const int*)ids_in.p, (int*)ids_out.p, n, 0, 64, m->stream));
  CUDA_TRY(cudaMemsetAsync(counts.p, 0, (size_t)(n + 1) * sizeof(unsigned int), m->stream));
  k_mesh_blocks<V, false><<<n, kMeshThreads, 0, m->stream>>>(m->view<V>(), (const int*)ids_out.p, m->d_mc_table, nullptr, (unsigned int*)counts.p, nullptr);
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, (const unsigned int*)counts.p, (unsigned long long*)offsets.p, n + 1, m->stream));
  unsigned long long total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, (unsigned long long*)offsets.p + n, sizeof(total), cudaMemcpyDeviceToHost, m->stream));
  if (int r = check_launch(m, 2)) return r;
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  if ((long long)total > m->mesh_capacity) {
    cudaFree(m->d_mesh); m->d_mesh = nullptr; m->mesh_capacity = 0;
    CUDA_TRY(cudaMalloc(&m->d_mesh, (size_t)total * 9 * sizeof(float)));
    m->mesh_capacity = (long long)total;
  }
  if (total) {
    k_mesh_blocks<V, true><<<n, kMeshThreads, 0, m->stream>>>(m->view<V>(), (const int*)ids_out.p, m->d_mc_table, (const unsigned long long*)offsets.p, nullptr, m->d_mesh);
    if (int r = check_launch(m, 1)) return r;
    CUDA_TRY(cudaStreamSynchronize(m->stream));
  }
  m->mesh_triangles = (long long)total;
  if (n_triangles) *n_triangles = (int64_t)total;
  return SE_B200_OK;
}


#define FIELD_DISPATCH(m, call_sdf, call_ofu) ((m)->field == SE_B200_SDF ? (call_sdf) : (call_ofu))
#define REQUIRE_MAP(m) do { if (!(m)) return fail(SE_B200_ERR_ARG, "null map"); } while (0)

extern "C" {

const char* se_b200_last_error(void) { return g_error.c_str(); }

int se_b200_bspline_lut(float out[1000]) {
  if (!out) return fail(SE_B200_ERR_ARG, "out is null");
  make_bspline_lut(out);
  return SE_B200_OK;
}

int se_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int se_b200_create(se_b200_map** out, int field_type, int size, float dim, int W, int H,
                   int64_t max_blocks, int64_t max_nodes, int device) {
  if (!out) return fail(SE_B200_ERR_ARG, "out is null");
  *out = nullptr;
  if (field_type != SE_B200_SDF && field_type != SE_B200_OFUSION) return fail(SE_B200_ERR_ARG, "field_type must be SE_B200_SDF or SE_B200_OFUSION");
  if (size < 16 || (size & (size - 1)) != 0 || size > (1 << 15)) return fail(SE_B200_ERR_ARG, "size must be a power of two in [16, 32768]");
  if (!(dim > 0.f) || W <= 0 || H <= 0) return fail(SE_B200_ERR_ARG, "dim, W, H must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(SE_B200_ERR_CUDA, "no CUDA device: this library has no CPU fallback"); }
  if (device < 0 || device >= ndev) return fail(SE_B200_ERR_ARG, "bad device index");
  DeviceGuard guard(device);

  se_b200_map* m = new se_b200_map;
  m->field = field_type; m->size = size; m->dim = dim; m->W = W; m->H = H; m->device = device;
  m->max_level = 0; while ((1 << m->max_level) < size) ++m->max_level;
  m->leaves_level = m->max_level - 3;
  m->voxel_bytes = field_type == SE_B200_SDF ? sizeof(SdfVoxel) : sizeof(OfuVoxel);
  const int64_t grid_blocks = (int64_t)(size / 8) * (size / 8) * (size / 8);
  if (max_blocks <= 0) max_blocks = std::min<int64_t>(grid_blocks, 1 << 20);      // 4 GiB (SDF) / 8 GiB (OFusion) of payload at most: sized for 180 GB of HBM
  max_blocks = std::min<int64_t>(max_blocks, grid_blocks);
  if (max_nodes <= 0) max_nodes = max_blocks / 4 + 4096;
  if (max_blocks > (1ll << 30) || max_nodes > (1ll << 28)) { delete m; return fail(SE_B200_ERR_ARG, "pool sizes too large"); }
  m->max_blocks = (int)max_blocks; m->max_nodes = (int)max_nodes;
  m->max_requests = m->max_blocks + m->max_nodes;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) m->num_sms = prop.multiProcessorCount;

  auto cleanup = [&](int code) { se_b200_destroy(m); return code; };
#define CREATE_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { fail(SE_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); return cleanup(SE_B200_ERR_CUDA); } } while (0)
  CREATE_TRY(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
  m->stream = m->own_stream;
  CREATE_TRY(cudaStreamCreateWithFlags(&m->alloc_stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaEventCreateWithFlags(&m->ev_integrated, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreateWithFlags(&m->ev_alloc_done, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreateWithFlags(&m->ev_depth_ready, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreateWithFlags(&m->ev_depth_read, cudaEventDisableTiming));
  { const char* e = std::getenv("SE_B200_NO_ALLOC_OVERLAP"); m->overlap_alloc = !(e && e[0] == '1'); }
  for (int i = 0; i < SE_B200_NUM_STAGES; ++i) { CREATE_TRY(cudaEventCreate(&m->ev_begin[i])); CREATE_TRY(cudaEventCreate(&m->ev_end[i])); }
  CREATE_TRY(cudaMallocHost(&m->h_counters, kNumCounters * sizeof(int)));
  CREATE_TRY(cudaHostAlloc(&m->h_status, 4 * sizeof(int), cudaHostAllocMapped));
  std::memset(m->h_status, 0, 4 * sizeof(int));
  CREATE_TRY(cudaHostGetDevicePointer((void**)&m->d_status, m->h_status, 0));
  const size_t npx = (size_t)W * H;
  // (one float more than the image, always 0: "no depth sample" for the voxels the check-free SDF integrate finds outside the image -- sdf_voxel_pair)
  CREATE_TRY(cudaMalloc(&m->d_depth, (npx + 1) * sizeof(float)));
  CREATE_TRY(cudaMalloc(&m->d_vertex, npx * 3 * sizeof(float)));
  if (!getenv("SE_B200_LAUNCH_ORDER_OFF")) {      // (measurement switch: image order, nothing recorded)
    m->d_ray_sched = make_schedule(pixel_tile_blocks(m->W, m->H, kRayThreads));
    m->d_alloc_sched = make_schedule(pixel_tile_blocks(m->W, m->H, kAllocThreads));
    if (!m->d_ray_sched || !m->d_alloc_sched) { fail(SE_B200_ERR_CUDA, "cudaMalloc (launch schedules)"); return cleanup(SE_B200_ERR_CUDA); }
  }
  CREATE_TRY(cudaMalloc(&m->d_normal, npx * 3 * sizeof(float)));
  CREATE_TRY(cudaMalloc(&m->d_rgba, npx * sizeof(uchar4)));
  CREATE_TRY(cudaMalloc(&m->d_active_list, (size_t)m->max_blocks * sizeof(int)));
  CREATE_TRY(cudaMemsetAsync(m->d_active_list, 0xFF, (size_t)m->max_blocks * sizeof(int), m->stream));      // kEmpty: the list is streamed (ActiveList)
  CREATE_TRY(cudaMemsetAsync(m->d_depth, 0, (npx + 1) * sizeof(float), m->stream));
  CREATE_TRY(cudaMemsetAsync(m->d_vertex, 0, npx * 3 * sizeof(float), m->stream));
  CREATE_TRY(cudaMemsetAsync(m->d_normal, 0, npx * 3 * sizeof(float), m->stream));
  float lut[1000];      // function scope: it outlives the stream-ordered copy below (create_pools ends with a synchronisation)
  if (field_type == SE_B200_OFUSION) {
    CREATE_TRY(cudaMalloc(&m->d_requests, (size_t)m->max_requests * sizeof(unsigned long long)));
    make_bspline_lut(lut);
    // on the map's stream: ordered before k_fill_logodds below and before every integrate (the map's stream is
    // non-blocking, so a copy on the legacy default stream would not be)
    CREATE_TRY(cudaMemcpyToSymbolAsync(c_bspline_lut, lut, sizeof(lut), 0, cudaMemcpyHostToDevice, m->stream));
    const int ncell = kLogOddsDim * kLogOddsDim;
    CREATE_TRY(cudaMalloc(&m->d_logodds, (size_t)ncell * sizeof(float)));
    k_fill_logodds<<<(ncell + 255) / 256, 256, 0, m->stream>>>(m->d_logodds);
    CREATE_TRY(cudaGetLastError());
  }
#undef CREATE_TRY
  const int r = field_type == SE_B200_SDF ? create_pools<SdfVoxel>(m) : create_pools<OfuVoxel>(m);
  if (r) return cleanup(r);
  *out = m;
  return SE_B200_OK;
}

int se_b200_destroy(se_b200_map* m) {
  if (!m) return SE_B200_OK;
  DeviceGuard guard(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  cudaFree(m->p.node_child); cudaFree(m->p.node_code); cudaFree(m->p.node_side); cudaFree(m->p.node_mask); cudaFree(m->p.node_value);
  cudaFree(m->p.block_code); cudaFree(m->p.block_coord); cudaFree(m->p.block_active); cudaFree(m->p.block_data); cudaFree(m->p.counters); cudaFree(m->p.dir); cudaFree(m->p.ndir); cudaFree(m->p.cmask);
  cudaFree(m->d_ray_sched); cudaFree(m->d_alloc_sched);
  cudaFree(m->d_depth); cudaFree(m->d_vertex); cudaFree(m->d_normal); cudaFree(m->d_rgba); cudaFree(m->d_depth_mm);
  cudaFree(m->d_active_list); cudaFree(m->d_miss); cudaFree(m->d_requests); cudaFree(m->d_logodds); cudaFree(m->d_track);
  for (int i = 0; i < 8; ++i) { cudaFree(m->d_scaled_depth[i]); cudaFree(m->d_in_vertex[i]); cudaFree(m->d_in_normal[i]); }
  cudaFree(m->d_trackdata); cudaFree(m->d_partial); cudaFree(m->d_reduction); cudaFree(m->d_icp);
  cudaFree(m->d_mc_table); cudaFree(m->d_mesh);
  if (m->h_reduction) cudaFreeHost(m->h_reduction);
  if (m->h_counters) cudaFreeHost(m->h_counters);
  if (m->h_status) cudaFreeHost(m->h_status);
  for (int i = 0; i < SE_B200_NUM_STAGES; ++i) { if (m->ev_begin[i]) cudaEventDestroy(m->ev_begin[i]); if (m->ev_end[i]) cudaEventDestroy(m->ev_end[i]); }
  if (m->aio.ready) {
    cudaStreamSynchronize(m->aio.up); cudaStreamSynchronize(m->aio.down);
    for (int i = 0; i < 2; ++i) {
      cudaFree(m->aio.d_mm[i]); cudaFree(m->aio.d_out[i]);
      cudaEventDestroy(m->aio.uploaded[i]); cudaEventDestroy(m->aio.consumed[i]); cudaEventDestroy(m->aio.rendered[i]); cudaEventDestroy(m->aio.downloaded[i]);
    }
    cudaStreamDestroy(m->aio.up); cudaStreamDestroy(m->aio.down);
  }
  if (m->alloc_stream) { cudaStreamSynchronize(m->alloc_stream); cudaStreamDestroy(m->alloc_stream); }
  if (m->ev_integrated) cudaEventDestroy(m->ev_integrated);
  if (m->ev_alloc_done) cudaEventDestroy(m->ev_alloc_done);
  if (m->ev_depth_ready) cudaEventDestroy(m->ev_depth_ready);
  if (m->ev_depth_read) cudaEventDestroy(m->ev_depth_read);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  cudaGetLastError();
  delete m;
  return SE_B200_OK;
}

int se_b200_set_stream(se_b200_map* m, void* s) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->alloc_stream));
  m->integrated_valid = m->depth_ready_valid = m->depth_read_valid = false;
  m->stream = s ? (cudaStream_t)s : m->own_stream;
  for (bool& v : m->ev_valid) v = false;
  return SE_B200_OK;
}

int se_b200_sync(se_b200_map* m) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->alloc_stream));
  if (m->aio.ready) { CUDA_TRY(cudaStreamSynchronize(m->aio.up)); CUDA_TRY(cudaStreamSynchronize(m->aio.down)); }
  return SE_B200_OK;
}

// a1 is deferred: remember where the millimetre image is; the allocation kernel of the next integrate converts it
// (resolve_depth() does for any other consumer of the float image)
static int preprocess_common(se_b200_map* m, const uint16_t* src_dev, int inW, int inH, int slot, cudaEvent_t upload) {
  (void)inH;
  m->pending_mm = src_dev; m->pending_inW = inW; m->pending_ratio = inW / m->W; m->pending_slot = slot; m->pending_upload = upload;
  return SE_B200_OK;
}
static int check_ratio(se_b200_map* m, int inW, int inH) {
  // preprocessing.cpp:165-176 ("Invalid ratio." + exit(1) in the reference)
  if (inW < m->W || inH < m->H) return fail(SE_B200_ERR_ARG, "Invalid ratio.");
  if (inW % m->W != 0 || inH % m->H != 0) return fail(SE_B200_ERR_ARG, "Invalid ratio.");
  if (inW / m->W != inH / m->H) return fail(SE_B200_ERR_ARG, "Invalid ratio.");
  return SE_B200_OK;
}

int se_b200_register_host_buffer(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(SE_B200_ERR_ARG, "null buffer");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(SE_B200_ERR_CUDA, "no CUDA device: this library has no CPU fallback"); }
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterMapped);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return SE_B200_OK; }
  if (e != cudaSuccess) { cudaGetLastError(); return fail(SE_B200_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
  return SE_B200_OK;
}

int se_b200_unregister_host_buffer(void* ptr) {
  if (!ptr) return fail(SE_B200_ERR_ARG, "null buffer");
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(SE_B200_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); }
  return SE_B200_OK;
}

int se_b200_preprocess_depth_host(se_b200_map* m, const uint16_t* depth_mm, int inW, int inH) {
  REQUIRE_MAP(m);
  if (!depth_mm) return fail(SE_B200_ERR_ARG, "depth_mm is null");
  if (int r = check_ratio(m, inW, inH)) return r;
  DeviceGuard guard(m->device);
  const size_t bytes = (size_t)inW * inH * sizeof(uint16_t);
  if (m->depth_mm_capacity < bytes) {
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    cudaFree(m->d_depth_mm); m->d_depth_mm = nullptr; m->depth_mm_capacity = 0;
    CUDA_TRY(cudaMalloc(&m->d_depth_mm, bytes));
    m->depth_mm_capacity = bytes;
  }
  stage_begin(m, SE_B200_STAGE_PREPROCESS);
  CUDA_TRY(cudaMemcpyAsync(m->d_depth_mm, depth_mm, bytes, cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaEventRecord(m->ev_depth_ready, m->stream)); m->depth_ready_valid = true;      // (the allocation kernel may run on another stream)
  if (int r = preprocess_common(m, m->d_depth_mm, inW, inH, -1, nullptr)) return r;
  stage_end(m, SE_B200_STAGE_PREPROCESS);
  return SE_B200_OK;
}

int se_b200_preprocess_depth_device(se_b200_map* m, const uint16_t* depth_mm_dev, int inW, int inH) {
  REQUIRE_MAP(m);
  if (!depth_mm_dev) return fail(SE_B200_ERR_ARG, "depth_mm_dev is null");
  if (int r = check_ratio(m, inW, inH)) return r;
  DeviceGuard guard(m->device);
  stage_begin(m, SE_B200_STAGE_PREPROCESS);
  m->depth_ready_valid = false;
  if (int r = preprocess_common(m, depth_mm_dev, inW, inH, -1, nullptr)) return r;
  stage_end(m, SE_B200_STAGE_PREPROCESS);
  return SE_B200_OK;
}

static int ensure_async_io(se_b200_map* m) {
  if (m->aio.ready) return SE_B200_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&m->aio.up, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&m->aio.down, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(cudaMalloc(&m->aio.d_out[i], (size_t)m->W * m->H * sizeof(uchar4)));
    CUDA_TRY(cudaEventCreateWithFlags(&m->aio.uploaded[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->aio.consumed[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->aio.rendered[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->aio.downloaded[i], cudaEventDisableTiming));
  }
  m->aio.ready = true;
  return SE_B200_OK;
}

int se_b200_preprocess_depth_host_async(se_b200_map* m, const uint16_t* depth_mm, int inW, int inH) {
  REQUIRE_MAP(m);
  if (!depth_mm) return fail(SE_B200_ERR_ARG, "depth_mm is null");
  if (int r = check_ratio(m, inW, inH)) return r;
  DeviceGuard guard(m->device);
  if (int r = ensure_async_io(m)) return r;
  auto& a = m->aio;
  const int s = (a.up_slot ^= 1);
  const size_t bytes = (size_t)inW * inH * sizeof(uint16_t);
  if (a.mm_capacity[s] < bytes) {
    CUDA_TRY(cudaStreamSynchronize(m->stream)); CUDA_TRY(cudaStreamSynchronize(a.up));
    cudaFree(a.d_mm[s]); a.d_mm[s] = nullptr; a.mm_capacity[s] = 0; a.consumed_valid[s] = false;
    CUDA_TRY(cudaMalloc(&a.d_mm[s], bytes));
    a.mm_capacity[s] = bytes;
  }
  if (a.consumed_valid[s]) CUDA_TRY(cudaStreamWaitEvent(a.up, a.consumed[s], 0));      // the frame two calls ago has read this buffer
  CUDA_TRY(cudaMemcpyAsync(a.d_mm[s], depth_mm, bytes, cudaMemcpyHostToDevice, a.up));
  CUDA_TRY(cudaEventRecord(a.uploaded[s], a.up));
  m->depth_ready_valid = false;
  stage_begin(m, SE_B200_STAGE_PREPROCESS);
  if (int r = preprocess_common(m, a.d_mm[s], inW, inH, s, a.uploaded[s])) return r;      // whoever converts the image waits for the upload and records `consumed[s]`
  stage_end(m, SE_B200_STAGE_PREPROCESS);
  return SE_B200_OK;
}

int se_b200_render_volume_host_async(se_b200_map* m, uint8_t* out, const float view_pose[16], const float k[4],
                                     float mu, float largestep, int reraycast) {
  REQUIRE_MAP(m);
  if (!out) return fail(SE_B200_ERR_ARG, "out is null");
  DeviceGuard guard(m->device);
  if (int r = ensure_async_io(m)) return r;
  auto& a = m->aio;
  const int s = (a.down_slot ^= 1);
  if (a.downloaded_valid[s]) CUDA_TRY(cudaStreamWaitEvent(m->stream, a.downloaded[s], 0));   // the image two calls ago has left this buffer
  if (int r = se_b200_render_volume_device(m, (uint8_t*)a.d_out[s], view_pose, k, mu, largestep, reraycast)) return r;
  CUDA_TRY(cudaEventRecord(a.rendered[s], m->stream));
  CUDA_TRY(cudaStreamWaitEvent(a.down, a.rendered[s], 0));
  CUDA_TRY(cudaMemcpyAsync(out, a.d_out[s], (size_t)m->W * m->H * 4, cudaMemcpyDeviceToHost, a.down));
  CUDA_TRY(cudaEventRecord(a.downloaded[s], a.down));
  a.downloaded_valid[s] = true;
  return SE_B200_OK;
}

int se_b200_set_depth_m_host(se_b200_map* m, const float* depth_m) {
  REQUIRE_MAP(m);
  if (!depth_m) return fail(SE_B200_ERR_ARG, "depth_m is null");
  DeviceGuard guard(m->device);
  m->pending_mm = nullptr; m->pending_slot = -1;           // this image replaces whatever preprocess left pending
  CUDA_TRY(cudaMemcpyAsync(m->d_depth, depth_m, (size_t)m->W * m->H * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_integrate(se_b200_map* m, const float pose[16], const float k[4], float mu, unsigned frame) {
  REQUIRE_MAP(m);
  if (!pose || !k) return fail(SE_B200_ERR_ARG, "pose/k is null");
  DeviceGuard guard(m->device);
  return FIELD_DISPATCH(m, integrate_impl<SdfVoxel>(m, pose, k, mu, frame), integrate_impl<OfuVoxel>(m, pose, k, mu, frame));
}

int se_b200_raycast(se_b200_map* m, const float pose[16], const float k[4], float mu) {
  REQUIRE_MAP(m);
  if (!pose || !k) return fail(SE_B200_ERR_ARG, "pose/k is null");
  DeviceGuard guard(m->device);
  return FIELD_DISPATCH(m, raycast_impl<SdfVoxel>(m, pose, k, mu, nullptr), raycast_impl<OfuVoxel>(m, pose, k, mu, nullptr));
}

int se_b200_raycast_count_samples(se_b200_map* m, const float pose[16], const float k[4], float mu, uint64_t samples[4]) {
  REQUIRE_MAP(m);
  if (!pose || !k || !samples) return fail(SE_B200_ERR_ARG, "null argument");
  DeviceGuard guard(m->device);
  Scratch d;
  CUDA_TRY(d.alloc(4 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(d.p, 0, 4 * sizeof(unsigned long long), m->stream));
  if (int r = FIELD_DISPATCH(m, raycast_impl<SdfVoxel>(m, pose, k, mu, (unsigned long long*)d.p), raycast_impl<OfuVoxel>(m, pose, k, mu, (unsigned long long*)d.p))) return r;
  CUDA_TRY(cudaMemcpyAsync(samples, d.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_download_vertex_normal(se_b200_map* m, float* vertex, float* normal) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  const size_t bytes = (size_t)m->W * m->H * 3 * sizeof(float);
  if (vertex) CUDA_TRY(cudaMemcpyAsync(vertex, m->d_vertex, bytes, cudaMemcpyDeviceToHost, m->stream));
  if (normal) CUDA_TRY(cudaMemcpyAsync(normal, m->d_normal, bytes, cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_upload_vertex_normal(se_b200_map* m, const float* vertex, const float* normal) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  const size_t bytes = (size_t)m->W * m->H * 3 * sizeof(float);
  m->rt_valid = false;
  if (vertex) CUDA_TRY(cudaMemcpyAsync(m->d_vertex, vertex, bytes, cudaMemcpyHostToDevice, m->stream));
  if (normal) CUDA_TRY(cudaMemcpyAsync(m->d_normal, normal, bytes, cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_render_volume_device(se_b200_map* m, uint8_t* out_dev, const float view_pose[16], const float k[4],
                                 float mu, float largestep, int reraycast) {
  REQUIRE_MAP(m);
  if (!out_dev || !view_pose || !k) return fail(SE_B200_ERR_ARG, "null argument");
  if (render_target_holds(m, out_dev, view_pose, reraycast)) return SE_B200_OK;      // se_b200_set_render_target: already there, in stream order
  DeviceGuard guard(m->device);
  return FIELD_DISPATCH(m, render_volume_impl<SdfVoxel>(m, (uchar4*)out_dev, view_pose, k, mu, largestep, reraycast),
                        render_volume_impl<OfuVoxel>(m, (uchar4*)out_dev, view_pose, k, mu, largestep, reraycast));
}

int se_b200_render_volume_host(se_b200_map* m, uint8_t* out, const float view_pose[16], const float k[4],
                               float mu, float largestep, int reraycast) {
  REQUIRE_MAP(m);
  if (!out) return fail(SE_B200_ERR_ARG, "out is null");
  if (!view_pose || !k) return fail(SE_B200_ERR_ARG, "null argument");
  DeviceGuard guard(m->device);
  if (render_target_holds(m, out, view_pose, reraycast)) {      // se_b200_set_render_target: the raycast wrote it; wait for it
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return SE_B200_OK;
  }
  if (void* alias = mapped_alias(out)) {            // pinned destination: the shading kernel writes it in place
    if (int r = se_b200_render_volume_device(m, (uint8_t*)alias, view_pose, k, mu, largestep, reraycast)) return r;
  } else {
    if (int r = se_b200_render_volume_device(m, (uint8_t*)m->d_rgba, view_pose, k, mu, largestep, reraycast)) return r;
    CUDA_TRY(cudaMemcpyAsync(out, m->d_rgba, (size_t)m->W * m->H * 4, cudaMemcpyDeviceToHost, m->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_set_render_target(se_b200_map* m, uint8_t* out) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  CUDA_TRY(cudaStreamSynchronize(m->stream));          // a raycast in flight may still be writing the previous target
  m->rt_dev = nullptr; m->rt_user = nullptr; m->rt_valid = false;
  if (!out) return SE_B200_OK;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, out) != cudaSuccess) { cudaGetLastError(); return fail(SE_B200_ERR_ARG, "render target: not a CUDA-visible pointer"); }
  void* dev = nullptr;
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) dev = out;
  else if (a.type == cudaMemoryTypeHost && a.devicePointer) dev = a.devicePointer;
  if (!dev) return fail(SE_B200_ERR_ARG, "render target must be device memory or page-locked (pinned) host memory");
  m->rt_dev = (uchar4*)dev; m->rt_user = out;
  return SE_B200_OK;
}

int se_b200_render_depth_host(se_b200_map* m, uint8_t* out) {
  REQUIRE_MAP(m);
  if (!out) return fail(SE_B200_ERR_ARG, "out is null");
  DeviceGuard guard(m->device);
  const int n = m->W * m->H;
  if (int r = resolve_depth(m)) return r;
  k_render_depth<<<(n + 255) / 256, 256, 0, m->stream>>>(m->d_rgba, m->d_depth, n, kNearPlane, kFarPlane);
  if (int r = check_launch(m)) return r;
  CUDA_TRY(cudaEventRecord(m->ev_depth_read, m->stream)); m->depth_read_valid = true;      // (the next allocation kernel overwrites the float depth from another stream)
  CUDA_TRY(cudaMemcpyAsync(out, m->d_rgba, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_render_track_host(se_b200_map* m, uint8_t* out, const int* track_result, int stride_ints) {
  REQUIRE_MAP(m);
  if (!out || (track_result && stride_ints < 1)) return fail(SE_B200_ERR_ARG, "bad argument");
  DeviceGuard guard(m->device);
  const int n = m->W * m->H;
  if (!track_result) {              // the result of the last se_b200_track, already on the device
    if (!m->d_trackdata) { if (int r = ensure_tracking_buffers(m, 1)) return r; }
    k_render_track<<<(n + 255) / 256, 256, 0, m->stream>>>(m->d_rgba, (const int*)m->d_trackdata, (int)(sizeof(TrackData) / sizeof(int)), n);
    if (int r = check_launch(m)) return r;
    CUDA_TRY(cudaMemcpyAsync(out, m->d_rgba, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return SE_B200_OK;
  }
  const size_t bytes = (size_t)n * stride_ints * sizeof(int);
  if (m->track_capacity < bytes) {
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    cudaFree(m->d_track); m->d_track = nullptr; m->track_capacity = 0;
    CUDA_TRY(cudaMalloc(&m->d_track, bytes));
    m->track_capacity = bytes;
  }
  CUDA_TRY(cudaMemcpyAsync(m->d_track, track_result, bytes, cudaMemcpyHostToDevice, m->stream));
  k_render_track<<<(n + 255) / 256, 256, 0, m->stream>>>(m->d_rgba, m->d_track, stride_ints, n);
  if (int r = check_launch(m)) return r;
  CUDA_TRY(cudaMemcpyAsync(out, m->d_rgba, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_block_count(se_b200_map* m, int* out) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  if (int r = fetch_counters(m)) return r;
  if (out) *out = std::min(m->h_counters[kCntBlocks], m->max_blocks);
  return check_pool_error(m);
}
int se_b200_node_count(se_b200_map* m, int* out) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  if (int r = fetch_counters(m)) return r;
  if (out) *out = std::min(m->h_counters[kCntNodes], m->max_nodes);
  return check_pool_error(m);
}

int se_b200_download_blocks_sorted(se_b200_map* m, uint64_t* keys, int32_t* coords, uint8_t* active, void* voxels) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  return FIELD_DISPATCH(m, download_blocks_sorted_impl<SdfVoxel>(m, keys, coords, active, voxels),
                        download_blocks_sorted_impl<OfuVoxel>(m, keys, coords, active, voxels));
}
int se_b200_download_nodes_sorted(se_b200_map* m, uint64_t* codes, uint32_t* side, uint8_t* mask, void* values) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  return FIELD_DISPATCH(m, download_nodes_sorted_impl<SdfVoxel>(m, codes, side, mask, values),
                        download_nodes_sorted_impl<OfuVoxel>(m, codes, side, mask, values));
}

int se_b200_allocate_keys(se_b200_map* m, const uint64_t* keys, int n) {
  REQUIRE_MAP(m);
  if (n < 0 || (n > 0 && !keys)) return fail(SE_B200_ERR_ARG, "bad key list");
  if (n == 0) return SE_B200_OK;              // the reference would process one stale key here (unique.hpp:51-60)
  DeviceGuard guard(m->device);
  // host side: sort + filter_ancestors to find keys[0] of the reference's list (its extra chain)
  std::vector<unsigned long long> k(keys, keys + n);
  std::sort(k.begin(), k.end());
  int e = 0;
  for (int i = 0; i < n; ++i) { if (key_descendant(k[i], k[e], m->max_level)) k[e] = k[i]; else k[++e] = k[i]; }
  k.resize(e + 1);
  if (key_level(k[0]) < m->leaves_level) k.push_back(key_code(k[0]) | (unsigned long long)m->leaves_level);
  Scratch d;
  CUDA_TRY(d.alloc(k.size() * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemcpyAsync(d.p, k.data(), k.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, m->stream));
  const int cnt = (int)k.size();
  if (m->field == SE_B200_SDF) k_allocate_keys<SdfVoxel><<<(cnt + 127) / 128, 128, 0, m->stream>>>(m->view<SdfVoxel>(), (unsigned long long*)d.p, cnt);
  else k_allocate_keys<OfuVoxel><<<(cnt + 127) / 128, 128, 0, m->stream>>>(m->view<OfuVoxel>(), (unsigned long long*)d.p, cnt);
  if (int r = check_launch(m)) return r;
  if (int r = fetch_counters(m)) return r;
  return check_pool_error(m);
}

int se_b200_upload_blocks(se_b200_map* m, const uint64_t* keys, const void* voxels, int n) {
  REQUIRE_MAP(m);
  if (n < 0 || (n > 0 && (!keys || !voxels))) return fail(SE_B200_ERR_ARG, "bad block list");
  if (n == 0) return SE_B200_OK;
  DeviceGuard guard(m->device);
  Scratch dk, dv;
  const size_t vbytes = (size_t)n * kBlockVoxels * m->voxel_bytes;
  CUDA_TRY(dk.alloc((size_t)n * sizeof(unsigned long long)));
  CUDA_TRY(dv.alloc(vbytes));
  CUDA_TRY(cudaMemcpyAsync(dk.p, keys, (size_t)n * sizeof(unsigned long long), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaMemcpyAsync(dv.p, voxels, vbytes, cudaMemcpyHostToDevice, m->stream));
  const int blocks = (n * 32 + 255) / 256;
  if (m->field == SE_B200_SDF) k_upload_blocks<SdfVoxel><<<blocks, 256, 0, m->stream>>>(m->view<SdfVoxel>(), (unsigned long long*)dk.p, (SdfVoxel*)dv.p, n);
  else k_upload_blocks<OfuVoxel><<<blocks, 256, 0, m->stream>>>(m->view<OfuVoxel>(), (unsigned long long*)dk.p, (OfuVoxel*)dv.p, n);
  if (int r = check_launch(m)) return r;
  if (int r = fetch_counters(m)) return r;
  return check_pool_error(m);
}

int se_b200_upload_nodes(se_b200_map* m, const uint64_t* codes, const void* values, int n) {
  REQUIRE_MAP(m);
  if (n < 0 || (n > 0 && (!codes || !values))) return fail(SE_B200_ERR_ARG, "bad node list");
  if (n == 0) return SE_B200_OK;
  DeviceGuard guard(m->device);
  Scratch dk, dv;
  const size_t vbytes = (size_t)n * 8 * m->voxel_bytes;
  CUDA_TRY(dk.alloc((size_t)n * sizeof(unsigned long long)));
  CUDA_TRY(dv.alloc(vbytes));
  CUDA_TRY(cudaMemcpyAsync(dk.p, codes, (size_t)n * sizeof(unsigned long long), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaMemcpyAsync(dv.p, values, vbytes, cudaMemcpyHostToDevice, m->stream));
  if (m->field == SE_B200_SDF) k_upload_nodes<SdfVoxel><<<(n + 127) / 128, 128, 0, m->stream>>>(m->view<SdfVoxel>(), (unsigned long long*)dk.p, (SdfVoxel*)dv.p, n);
  else k_upload_nodes<OfuVoxel><<<(n + 127) / 128, 128, 0, m->stream>>>(m->view<OfuVoxel>(), (unsigned long long*)dk.p, (OfuVoxel*)dv.p, n);
  if (int r = check_launch(m)) return r;
  if (int r = fetch_counters(m)) return r;
  return check_pool_error(m);
}

int se_b200_query_voxels(se_b200_map* m, const int32_t* xyz, int n, void* out) {
  REQUIRE_MAP(m);
  if (n <= 0) return SE_B200_OK;
  DeviceGuard guard(m->device);
  Scratch dx, dout;
  CUDA_TRY(dx.alloc((size_t)n * 3 * sizeof(int)));
  CUDA_TRY(dout.alloc((size_t)n * m->voxel_bytes));
  CUDA_TRY(cudaMemcpyAsync(dx.p, xyz, (size_t)n * 3 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  if (m->field == SE_B200_SDF) k_query_voxels<SdfVoxel><<<(n + 127) / 128, 128, 0, m->stream>>>(m->view<SdfVoxel>(), (int*)dx.p, n, (SdfVoxel*)dout.p);
  else k_query_voxels<OfuVoxel><<<(n + 127) / 128, 128, 0, m->stream>>>(m->view<OfuVoxel>(), (int*)dx.p, n, (OfuVoxel*)dout.p);
  if (int r = check_launch(m)) return r;
  CUDA_TRY(cudaMemcpyAsync(out, dout.p, (size_t)n * m->voxel_bytes, cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_query_interp(se_b200_map* m, const float* pos, int n, float* out) {
  REQUIRE_MAP(m);
  if (n <= 0) return SE_B200_OK;
  DeviceGuard guard(m->device);
  Scratch dx, dout;
  CUDA_TRY(dx.alloc((size_t)n * 3 * sizeof(float)));
  CUDA_TRY(dout.alloc((size_t)n * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(dx.p, pos, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  if (m->field == SE_B200_SDF) k_query_interp<SdfVoxel><<<(n + kRayThreads - 1) / kRayThreads, kRayThreads, 0, m->stream>>>(m->view<SdfVoxel>(), (float*)dx.p, n, (float*)dout.p);
  else k_query_interp<OfuVoxel><<<(n + kRayThreads - 1) / kRayThreads, kRayThreads, 0, m->stream>>>(m->view<OfuVoxel>(), (float*)dx.p, n, (float*)dout.p);
  if (int r = check_launch(m)) return r;
  CUDA_TRY(cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_query_grad(se_b200_map* m, const float* pos, int n, float* out) {
  REQUIRE_MAP(m);
  if (n <= 0) return SE_B200_OK;
  DeviceGuard guard(m->device);
  Scratch dx, dout;
  CUDA_TRY(dx.alloc((size_t)n * 3 * sizeof(float)));
  CUDA_TRY(dout.alloc((size_t)n * 3 * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(dx.p, pos, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  if (m->field == SE_B200_SDF) k_query_grad<SdfVoxel><<<(n + kRayThreads - 1) / kRayThreads, kRayThreads, 0, m->stream>>>(m->view<SdfVoxel>(), (float*)dx.p, n, (float*)dout.p);
  else k_query_grad<OfuVoxel><<<(n + kRayThreads - 1) / kRayThreads, kRayThreads, 0, m->stream>>>(m->view<OfuVoxel>(), (float*)dx.p, n, (float*)dout.p);
  if (int r = check_launch(m)) return r;
  CUDA_TRY(cudaMemcpyAsync(out, dout.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}


// ---- N4: meshing ----------------------------------------------------------------------------
int se_b200_extract_mesh(se_b200_map* m, int64_t* n_triangles) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  return FIELD_DISPATCH(m, extract_mesh_impl<SdfVoxel>(m, n_triangles), extract_mesh_impl<OfuVoxel>(m, n_triangles));
}

int se_b200_download_mesh(se_b200_map* m, float* triangles, int64_t capacity) {
  REQUIRE_MAP(m);
  if (capacity < m->mesh_triangles) return fail(SE_B200_ERR_ARG, "mesh buffer smaller than the triangle count se_b200_extract_mesh returned");
  if (m->mesh_triangles == 0) return SE_B200_OK;
  if (!triangles) return fail(SE_B200_ERR_ARG, "null mesh buffer");
  DeviceGuard guard(m->device);
  CUDA_TRY(cudaMemcpyAsync(triangles, m->d_mesh, (size_t)m->mesh_triangles * 9 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

void se_b200_mc_table(int8_t table[4096]) { mc_case_table(table); }

int se_b200_set_voxels(se_b200_map* m, const int32_t* xyz, const void* voxels, int n) {
  REQUIRE_MAP(m);
  if (n <= 0) return SE_B200_OK;
  DeviceGuard guard(m->device);
  Scratch dx, dv;
  CUDA_TRY(dx.alloc((size_t)n * 3 * sizeof(int)));
  CUDA_TRY(dv.alloc((size_t)n * m->voxel_bytes));
  CUDA_TRY(cudaMemcpyAsync(dx.p, xyz, (size_t)n * 3 * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  CUDA_TRY(cudaMemcpyAsync(dv.p, voxels, (size_t)n * m->voxel_bytes, cudaMemcpyHostToDevice, m->stream));
  if (m->field == SE_B200_SDF) k_set_voxels<SdfVoxel><<<(n + 127) / 128, 128, 0, m->stream>>>(m->view<SdfVoxel>(), (int*)dx.p, (SdfVoxel*)dv.p, n);
  else k_set_voxels<OfuVoxel><<<(n + 127) / 128, 128, 0, m->stream>>>(m->view<OfuVoxel>(), (int*)dx.p, (OfuVoxel*)dv.p, n);
  if (int r = check_launch(m)) return r;
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_query_rays(se_b200_map* m, const float* origin_dir, int n, float near_plane, float far_plane,
                       uint64_t* first_block_key, float* tinfo) {
  REQUIRE_MAP(m);
  if (n <= 0) return SE_B200_OK;
  DeviceGuard guard(m->device);
  Scratch dx, dk, dt;
  CUDA_TRY(dx.alloc((size_t)n * 6 * sizeof(float)));
  CUDA_TRY(dk.alloc((size_t)n * sizeof(unsigned long long)));
  CUDA_TRY(dt.alloc((size_t)n * 3 * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(dx.p, origin_dir, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  const int qgrid = (n + kRayThreads - 1) / kRayThreads;
  if (m->field == SE_B200_SDF) {
    if (m->p.cmask) k_query_ray<SdfVoxel, true><<<qgrid, kRayThreads, 0, m->stream>>>(m->view<SdfVoxel>(), (float*)dx.p, n, near_plane, far_plane, (unsigned long long*)dk.p, (float*)dt.p);
    else k_query_ray<SdfVoxel, false><<<qgrid, kRayThreads, 0, m->stream>>>(m->view<SdfVoxel>(), (float*)dx.p, n, near_plane, far_plane, (unsigned long long*)dk.p, (float*)dt.p);
  } else {
    if (m->p.cmask) k_query_ray<OfuVoxel, true><<<qgrid, kRayThreads, 0, m->stream>>>(m->view<OfuVoxel>(), (float*)dx.p, n, near_plane, far_plane, (unsigned long long*)dk.p, (float*)dt.p);
    else k_query_ray<OfuVoxel, false><<<qgrid, kRayThreads, 0, m->stream>>>(m->view<OfuVoxel>(), (float*)dx.p, n, near_plane, far_plane, (unsigned long long*)dk.p, (float*)dt.p);
  }
  if (int r = check_launch(m)) return r;
  if (first_block_key) CUDA_TRY(cudaMemcpyAsync(first_block_key, dk.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
  if (tinfo) CUDA_TRY(cudaMemcpyAsync(tinfo, dt.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_elapsed_ms(se_b200_map* m, int stage, float* ms) {
  REQUIRE_MAP(m);
  if (stage < 0 || stage >= SE_B200_NUM_STAGES || !ms) return fail(SE_B200_ERR_ARG, "bad stage");
  if (!m->ev_valid[stage]) { *ms = 0.f; return SE_B200_OK; }
  DeviceGuard guard(m->device);
  CUDA_TRY(cudaEventSynchronize(m->ev_end[stage]));
  CUDA_TRY(cudaEventElapsedTime(ms, m->ev_begin[stage], m->ev_end[stage]));
  return SE_B200_OK;
}

int se_b200_set_stage_timing(se_b200_map* m, int enable) {
  REQUIRE_MAP(m);
  m->stage_timing = enable != 0;
  if (!m->stage_timing) for (bool& v : m->ev_valid) v = false;
  return SE_B200_OK;
}

int se_b200_counters(se_b200_map* m, int32_t out[8]) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  if (int r = fetch_counters(m)) return r;
  for (int i = 0; i < 8; ++i) out[i] = m->h_counters[i];
  out[2] = m->h_counters[counter_slot(kCntActive, m->parity)];
  out[6] = m->h_counters[kCntKeysReport];
  out[7] = 0;
  return check_pool_error(m);
}

#ifdef SE_TIMELINE
int se_b200_debug_timeline(se_b200_map* m, unsigned long long* out, int n) {
  cudaStreamSynchronize(m->stream);
  return cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * (size_t)n) == cudaSuccess ? 0 : -1;
}
#endif
int se_b200_launch_count(se_b200_map* m, int64_t* out) {
  REQUIRE_MAP(m);
  if (out) *out = m->launches;
  return SE_B200_OK;
}

int se_b200_device_image(se_b200_map* m, int which, void** ptr) {
  REQUIRE_MAP(m);
  if (!ptr) return fail(SE_B200_ERR_ARG, "ptr is null");
  switch (which) {
    case 0: { DeviceGuard guard(m->device); if (int r = resolve_depth(m)) return r; } *ptr = m->d_depth; break;
    case 1: *ptr = m->d_vertex; break;
    case 2: *ptr = m->d_normal; break;
    default: return fail(SE_B200_ERR_ARG, "which must be 0..2");
  }
  return SE_B200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// N1: tracking front-end
// ---------------------------------------------------------------------------------------------
namespace {

int ensure_tracking_buffers(se_b200_map* m, int levels) {
  if (levels < 1 || levels > 8) return fail(SE_B200_ERR_ARG, "pyramid levels must be in [1, 8]");
  if ((m->W >> (levels - 1)) < 1 || (m->H >> (levels - 1)) < 1) return fail(SE_B200_ERR_ARG, "too many pyramid levels for this image size");
  for (int i = 0; i < levels; ++i) {
    if (m->d_scaled_depth[i]) continue;
    const size_t n = (size_t)(m->W >> i) * (m->H >> i);
    CUDA_TRY(cudaMalloc(&m->d_scaled_depth[i], n * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->d_in_vertex[i], n * 3 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->d_in_normal[i], n * 3 * sizeof(float)));
    CUDA_TRY(cudaMemsetAsync(m->d_scaled_depth[i], 0, n * sizeof(float), m->stream));
    CUDA_TRY(cudaMemsetAsync(m->d_in_vertex[i], 0, n * 3 * sizeof(float), m->stream));
    CUDA_TRY(cudaMemsetAsync(m->d_in_normal[i], 0, n * 3 * sizeof(float), m->stream));
  }
  if (!m->d_trackdata) {
    const size_t n = (size_t)m->W * m->H;
    CUDA_TRY(cudaMalloc(&m->d_trackdata, n * sizeof(TrackData)));
    CUDA_TRY(cudaMemsetAsync(m->d_trackdata, 0, n * sizeof(TrackData), m->stream));
    CUDA_TRY(cudaMalloc(&m->d_partial, ((n + kTrackThreads - 1) / kTrackThreads) * 32 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->d_reduction, 32 * sizeof(float)));
    CUDA_TRY(cudaMemsetAsync(m->d_reduction, 0, 32 * sizeof(float), m->stream));
    CUDA_TRY(cudaMalloc(&m->d_icp, sizeof(IcpState)));
    CUDA_TRY(cudaMallocHost(&m->h_reduction, 48 * sizeof(float)));
    std::memset(m->h_reduction, 0, 48 * sizeof(float));
  }
  m->levels = std::max(m->levels, levels);
  return SE_B200_OK;
}

dim3 grid2d(int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8); }

}  // namespace

extern "C" {

int se_b200_filter_depth(se_b200_map* m, int filter, int levels) {
  REQUIRE_MAP(m);
  DeviceGuard guard(m->device);
  if (int r = ensure_tracking_buffers(m, levels)) return r;
  if (int r = resolve_depth(m)) return r;
  if (filter) {
    Gauss5 gs;
    for (int i = 0; i < 5; ++i) { const int x = i - 2; gs.g[i] = expf(-(float)(x * x) / (2 * kGaussDelta * kGaussDelta)); }   // DenseSLAMSystem.cpp:111-118
    launch_pdl(k_bilateral, grid2d(m->W, m->H), dim3(32, 8), 0, m->stream, m->d_scaled_depth[0], m->d_depth, m->W, m->H, gs);
    if (int r = check_launch(m)) return r;
  } else {
    CUDA_TRY(cudaMemcpyAsync(m->d_scaled_depth[0], m->d_depth, (size_t)m->W * m->H * sizeof(float), cudaMemcpyDeviceToDevice, m->stream));
  }
  CUDA_TRY(cudaEventRecord(m->ev_depth_read, m->stream)); m->depth_read_valid = true;
  return SE_B200_OK;
}

int se_b200_track(se_b200_map* m, float pose_io[16], const float raycast_pose[16], const float k[4], float icp_threshold,
                  const int* iterations, int levels, int* tracked) {
  REQUIRE_MAP(m);
  if (!pose_io || !raycast_pose || !k || !iterations) return fail(SE_B200_ERR_ARG, "null argument");
  DeviceGuard guard(m->device);
  if (int r = ensure_tracking_buffers(m, levels)) return r;
  const dim3 tb(32, 8);
  // pyramid + per-level vertex / normal maps (DenseSLAMSystem.cpp:149-164)
  for (int i = 1; i < levels; ++i)
    launch_pdl(k_half_sample, grid2d(m->W >> i, m->H >> i), tb, 0, m->stream, m->d_scaled_depth[i], m->d_scaled_depth[i - 1], m->W >> i, m->H >> i, m->W >> (i - 1), kEDelta * 3, 1);
  for (int i = 0; i < levels; ++i) {
    const float s = (float)(1 << i);
    const float ks[4] = { k[0] / s, k[1] / s, k[2] / s, k[3] / s };
    launch_pdl(k_depth2vertex, grid2d(m->W >> i, m->H >> i), tb, 0, m->stream, m->d_in_vertex[i], m->d_scaled_depth[i], m->W >> i, m->H >> i, inverse_camera_matrix(ks));
    launch_pdl(k_vertex2normal, grid2d(m->W >> i, m->H >> i), tb, 0, m->stream, m->d_in_normal[i], m->d_in_vertex[i], m->W >> i, m->H >> i, k[1] < 0 ? 1 : 0);
  }
  if (int r = check_launch(m, 3 * levels - 1)) return r;

  // The coarse-to-fine ICP loop (DenseSLAMSystem.cpp:169-186) is enqueued as a whole: the pose lives on the device,
  // k_icp_update solves and applies each step there and raises `converged` where the reference breaks out of a level
  // (later iterations of that level then return at once).  One copy back at the end.
  M4 pose = to_m4(pose_io);
  const M4 old_pose = pose;
  IcpState init;
  std::memcpy(init.pose, pose.m, sizeof(init.pose));
  init.converged = 0; init.pad_[0] = init.pad_[1] = init.pad_[2] = 0;
  CUDA_TRY(cudaMemcpyAsync(m->d_icp, &init, sizeof(init), cudaMemcpyHostToDevice, m->stream));   // pageable source: staged before the call returns
  TrackParams tp;
  tp.view = mul44(camera_matrix(k), rigid_inverse(to_m4(raycast_pose)));          // projectReference, :167
  tp.refW = m->W; tp.refH = m->H; tp.dist_threshold = kDistThreshold; tp.normal_threshold = kNormalThreshold;
  int launched = 0;
  for (int level = levels - 1; level >= 0; --level) {
    tp.inW = m->W / (1 << level); tp.inH = m->H / (1 << level);
    const int n = tp.inW * tp.inH, ctas = (n + kTrackThreads - 1) / kTrackThreads;
    if (iterations[level] > 0) CUDA_TRY(cudaMemsetAsync(&m->d_icp->converged, 0, sizeof(int), m->stream));     // a new level starts unconverged
    for (int i = 0; i < iterations[level]; ++i) {
      launch_pdl(k_track, ctas, kTrackThreads, 0, m->stream, m->d_trackdata, m->d_in_vertex[level], m->d_in_normal[level], m->d_vertex, m->d_normal, tp, m->d_icp, m->d_partial);
      launch_pdl(k_icp_update, 1, 256, 0, m->stream, m->d_partial, ctas, m->d_reduction, m->d_icp, icp_threshold);
      launched += 2;
    }
  }
  if (int r = check_launch(m, launched)) return r;
  CUDA_TRY(cudaMemcpyAsync(m->h_reduction, m->d_reduction, 32 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaMemcpyAsync(m->h_reduction + 32, m->d_icp, 16 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  std::memcpy(pose.m, m->h_reduction + 32, sizeof(pose.m));
  // checkPoseKernel (tracking.cpp:320-336)
  const float* v = m->h_reduction;
  bool ok = true;
  if ((std::sqrt(v[0] / v[28]) > 2e-2) || (v[28] / (float)(m->W * m->H) < kTrackThreshold)) { pose = old_pose; ok = false; }
  std::memcpy(pose_io, pose.m, sizeof(pose.m));
  if (tracked) *tracked = ok ? 1 : 0;
  return SE_B200_OK;
}

int se_b200_download_pyramid(se_b200_map* m, int level, float* depth, float* vertex, float* normal) {
  REQUIRE_MAP(m);
  if (level < 0 || level >= m->levels || !m->d_scaled_depth[level]) return fail(SE_B200_ERR_ARG, "pyramid level not built");
  DeviceGuard guard(m->device);
  const size_t n = (size_t)(m->W >> level) * (m->H >> level);
  if (depth) CUDA_TRY(cudaMemcpyAsync(depth, m->d_scaled_depth[level], n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  if (vertex) CUDA_TRY(cudaMemcpyAsync(vertex, m->d_in_vertex[level], n * 3 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  if (normal) CUDA_TRY(cudaMemcpyAsync(normal, m->d_in_normal[level], n * 3 * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return SE_B200_OK;
}

int se_b200_download_tracking(se_b200_map* m, void* track_data, float reduction[32]) {
  REQUIRE_MAP(m);
  if (!m->d_trackdata) return fail(SE_B200_ERR_ARG, "tracking has not run");
  DeviceGuard guard(m->device);
  if (track_data) CUDA_TRY(cudaMemcpyAsync(track_data, m->d_trackdata, (size_t)m->W * m->H * sizeof(TrackData), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (reduction) std::memcpy(reduction, m->h_reduction, 32 * sizeof(float));
  return SE_B200_OK;
}

}  // extern "C"
