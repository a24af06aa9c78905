// se_map.cuh -- the octree as flat, index-addressed pools in HBM (no host pointers) and the
// device-side accessors every kernel shares.
//
// Layout (DESIGN.md "Data layout in HBM"):
//   nodes  : SoA, index 0 is the root.   node_child[8n+i] = index of child i (a node index
//            above the leaves level, a block index at it), kEmpty when absent, kBusy while a
//            thread is creating it.  node_code / node_side / node_mask / node_value[8n+i]
//            mirror Node<T>::code_/side_/children_mask_/value_ (se_core/include/se/node.hpp:45-90).
//   blocks : block_code, block_coord (low corner, voxel units), block_active and
//            block_data[512 b + x + 8y + 64z] mirror VoxelBlock<T> (node.hpp:92-145).  One
//            block's payload is one contiguous, 4 KiB (SDF) / 8 KiB (OFusion) aligned run, so a
//            warp streams it as float4/int4 rows.
//   dir    : optional dense (size/8)^3 directory block coordinate -> block index (O(1) fetch);
//            64^3 x 4 B = 1 MiB at 512^3, 64 MiB at 2048^3 -- affordable with 180 GB of HBM.
//   counters[kCntNodes/kCntBlocks] are bump allocators (MemoryPool::acquire_block,
//            se_core/include/se/utils/memory_pool.hpp:69-76); pools are pre-initialised to
//            initValue() at creation, so allocation never touches the payload.
//
// Accessor semantics follow se_core/include/se/octree.hpp (fetch :440-458, get_fine :356-377,
// interp :541-563 + interpolation/interp_gather.hpp:105-237, grad :652-737).
#pragma once
#include "se_math.cuh"
#include "se_ptx.cuh"

namespace se_b200 {

constexpr int kEmpty = -1;
constexpr int kBusy = -2;

// Line 0 (ints 0..31): pool bump allocators and per-frame bookkeeping.  The counters the integrate kernels' in-kernel
// active list hammers (build_active_list) each sit in a 128-byte line of their own, double-buffered by frame parity
// q = frame & 1 at  base + 32 q  -- frame f uses slot q and clears slot q ^ 1 for the next frame, so no reset launch is needed:
//   kCntTicket  chunk tickets drawn          kCntDone  chunks finished          kCntActive  active-list length
//   kCntMiss    length of the allocation pass's list of missing blocks (k_alloc_sdf appends, the integrate kernel creates)
//   kCntTake    list entries handed to the integrate kernel's warps beyond each warp's first two (ActiveList::draw): kTakeClasses
//               counters, 32 ints apart, then the same again for the other parity
enum Counter { kCntNodes = 0, kCntBlocks = 1, kCntError = 3, kCntNewBlocksBase = 4, kCntNewNodesBase = 5,
               kCntKeys = 6, kCntKeysReport = 7, kCntLastBlocks = 10, kCntLastNodes = 11, kCntBlocksBefore = 12, kCntNodesBefore = 13,
               kCntTicket = 32, kCntDone = 96, kCntActive = 160, kCntMiss = 224, kCntTake = 288, kNumCounters = 288 + 2 * 16 * 32 };
constexpr int kTakeClasses = 16;      // ticket counters of ActiveList::draw, each in a 128-byte line of its own, per frame parity
SE_HD int counter_slot(int base, int parity) { return base + 32 * parity; }
enum ErrorBits { kErrBlockPoolFull = 1, kErrNodePoolFull = 2, kErrKeyListFull = 4, kErrMissListFull = 8 };

// ---- field types (se_denseslam/include/se/volume_traits.hpp:41-72) --------------------
struct SdfVoxel { float x; float y; };                                    // tsdf, weight
struct __align__(16) OfuVoxel { float x; float pad_; double y; };         // log-odds, timestamp (double, as in the reference)

template <class V> struct FieldTraits;
template <> struct FieldTraits<SdfVoxel> {
  static constexpr bool is_sdf = true;
  SE_HD static SdfVoxel init() { SdfVoxel v; v.x = 1.f; v.y = 0.f; return v; }
  SE_HD static float empty_x() { return 1.f; }
};
template <> struct FieldTraits<OfuVoxel> {
  static constexpr bool is_sdf = false;
  SE_HD static OfuVoxel init() { OfuVoxel v; v.x = 0.f; v.pad_ = 0.f; v.y = 0.0; return v; }
  SE_HD static float empty_x() { return 0.f; }
};
// NB: for both field types empty().x == initValue().x (1 for the TSDF, 0 for the log-odds); gather_points relies on it.

template <class V> struct MapView {
  int size;            // voxels per side
  float dim;           // metres per side
  // (float)size / dim and (0.5f * dim) / (float)size: the metres -> voxels factor of VolumeTemplate::{get,interp,grad}
  // (volume_template.hpp:77-102) and the gradient's scale (octree.hpp:735), each ONE correctly rounded division that is the
  // same for every sample -- computed once on the host (se_b200_map::view) instead of per call on the device
  float inv_voxel, grad_scale;
  int max_level;       // log2(size)
  int leaves_level;    // max_level - 3
  int max_nodes, max_blocks;
  int* node_child;
  unsigned long long* node_code;
  unsigned int* node_side;
  unsigned int* node_mask;
  V* node_value;
  unsigned long long* block_code;
  int4* block_coord;   // x, y, z, unused
  int* block_active;
  V* block_data;       // max_blocks + 1 payloads: the last one is never allocated and always holds initValue() (reads of unallocated blocks go there)
  int* counters;
  // Block directory: dense (size/8)^3 grid of block indices (kEmpty where nothing is allocated),
  // cell = bx + G (by + G bz).  It turns Octree::fetch into one load.  The octree nodes stay the
  // authoritative structure (ray walk, node values, export); the directory is a pure index on the
  // leaves, written once when a block is created.  nullptr => fall back to the tree descent.
  int* dir;
  int dir_dim;         // G = size / 8
  // Node directory: the same idea for the internal levels 1 .. leaves_level-1 (level l is a dense (2^l)^3 grid
  // at offset (8^l - 8) / 7).  It costs 1/7 of the block directory and lets the multi-level (OFusion)
  // allocation pass test "does the octant at level l exist" with one load.  nullptr => tree descent.
  int* ndir;
  // Children masks by implicit position: the octant with heap index g (root 0, child s of g = 8 g + 1 + s, s = x | y<<1 | z<<2)
  // has one byte, bit s set <=> its child s exists -- Node::children_mask_ addressed by where the octant IS instead of
  // by a pool index, for the levels 0 .. leaves_level-1: (8^leaves_level - 1) / 7 bytes (37 KB at 512^3, 2.4 MB at 2048^3).
  // The ray walk (RayWalk, se_kernels.cuh) reads one byte per descent instead of one child pointer per step, and the
  // byte's address follows from the ray's position, not from the previous load.  nullptr => the walk chases node_child.
  unsigned char* cmask;
};

// heap index of the first octant of `level` (= number of octants above it)
SE_HD unsigned heap_level_offset(int level) { return (unsigned)(((1ull << (3 * level)) - 1ull) / 7ull); }

// index of the level-`level` octant containing voxel (x, y, z) in MapView::ndir
template <class V>
SE_HD int node_dir_index(const MapView<V>& m, int x, int y, int z, int level) {
  const int sh = m.max_level - level;
  const int off = (int)(((1ll << (3 * level)) - 8) / 7);
  return off + ((((z >> sh) << level) | (y >> sh)) << level | (x >> sh));
}

// ---- device accessors -------------------------------------------------------------------
#ifdef __CUDACC__

template <class V>
__device__ __forceinline__ bool in_volume(const MapView<V>& m, int x, int y, int z) {
  return ((unsigned)x < (unsigned)m.size) & ((unsigned)y < (unsigned)m.size) & ((unsigned)z < (unsigned)m.size);
}

// Octree::fetch (octree.hpp:440-458); kEmpty when the block is not allocated.  Coordinates
// outside the volume read as "not allocated" (the reference indexes out of bounds there).
// Read-only accessors (this one, get_fine, interp, grad) go through the non-coherent cache
// (__ldg): they are only used by kernels that do not modify the tree.
// the tree descent itself, kept out of line: with the directory it is only the fallback.  (The out-of-line helpers of this
// file take what they need BY VALUE: a reference to the kernel's MapView parameter would force every thread to copy the
// whole 140-byte struct to local memory at kernel entry -- 17 STL.64 per thread in the round-1 kernels.)
__device__ __noinline__ int fetch_block_tree_impl(const int* __restrict__ node_child, int size, int x, int y, int z) {
  int n = 0;
  for (int edge = size >> 1; edge >= kBlockSide; edge >>= 1) {
    const int slot = ((x & edge) != 0) | (((y & edge) != 0) << 1) | (((z & edge) != 0) << 2);
    n = __ldg(node_child + 8 * n + slot);
    if (n < 0) return kEmpty;
  }
  return n;
}
template <class V>
__device__ __forceinline__ int fetch_block_tree(const MapView<V>& m, int x, int y, int z) { return fetch_block_tree_impl(m.node_child, m.size, x, y, z); }
template <class V>
__device__ __forceinline__ int fetch_block(const MapView<V>& m, int x, int y, int z) {
  if (!in_volume(m, x, y, z)) return kEmpty;
  if (m.dir) return __ldg(m.dir + ((z >> 3) * m.dir_dim + (y >> 3)) * m.dir_dim + (x >> 3));
  return fetch_block_tree(m, x, y, z);
}
// the same for a BLOCK coordinate (gx, gy, gz) = voxel >> 3, or kEmpty outside the grid: one compare chain, one
// multiply-add chain, one load (the gather / gradient neighbourhoods are enumerated in block coordinates)
template <class V>
__device__ __forceinline__ int fetch_block_cell(const MapView<V>& m, int gx, int gy, int gz) {
  const unsigned G = (unsigned)m.dir_dim;
  if (!(((unsigned)gx < G) & ((unsigned)gy < G) & ((unsigned)gz < G))) return kEmpty;
  if (m.dir) return __ldg(m.dir + (gz * (int)G + gy) * (int)G + gx);
  return fetch_block_tree(m, gx << 3, gy << 3, gz << 3);
}

// Octree::fetch_octant (octree.hpp:460-478): node/block at `depth`, is_block tells which pool.
template <class V>
__device__ __forceinline__ int fetch_octant(const MapView<V>& m, int x, int y, int z, int depth, bool& is_block) {
  is_block = false;
  if (!in_volume(m, x, y, z)) return kEmpty;
  int n = 0;
  int d = 1;
  for (int edge = m.size >> 1; edge >= kBlockSide && d <= depth; edge >>= 1, ++d) {
    const int slot = ((x & edge) != 0) | (((y & edge) != 0) << 1) | (((z & edge) != 0) << 2);
    n = __ldcg(m.node_child + 8 * n + slot);
    if (n < 0) return kEmpty;
    is_block = (edge == kBlockSide);
  }
  return n;
}

// A one-entry per-thread cache of the last block looked up: successive samples of a ray, the
// 8 corners of an interpolation and the 32 voxels of a gradient mostly fall in one block, so
// this removes most root-to-leaf descents without changing any result.
struct BlockCache {
  int bx, by, bz, idx;
  int n_get, n_interp, n_grad, n_walk;     // sample counters (SURVEY 8(d) algorithmic bytes); dead code unless a kernel reads them
  __device__ __forceinline__ BlockCache() : bx(-1), by(-1), bz(-1), idx(kEmpty), n_get(0), n_interp(0), n_grad(0), n_walk(0) {}
};
template <class V>
__device__ __forceinline__ int fetch_block_cached(const MapView<V>& m, BlockCache& c, int x, int y, int z) {
  const int bx = x >> 3, by = y >> 3, bz = z >> 3;
  if (bx == c.bx && by == c.by && bz == c.bz) return c.idx;
  const int idx = fetch_block(m, x, y, z);
  if (in_volume(m, x, y, z)) { c.bx = bx; c.by = by; c.bz = bz; c.idx = idx; }
  return idx;
}

__device__ __forceinline__ float load_x(const SdfVoxel* p) { return __ldg(&p->x); }
__device__ __forceinline__ float load_x(const OfuVoxel* p) { return __ldg(&p->x); }
__device__ __forceinline__ SdfVoxel load_voxel(const SdfVoxel* p) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(p)); SdfVoxel v; v.x = t.x; v.y = t.y; return v;
}
__device__ __forceinline__ OfuVoxel load_voxel(const OfuVoxel* p) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  OfuVoxel v; v.x = __int_as_float((int)(__double_as_longlong(t.x) & 0xffffffffll)); v.pad_ = 0.f; v.y = t.y; return v;
}

template <class V>
__device__ __forceinline__ int voxel_offset(int x, int y, int z) { return (x & 7) | ((y & 7) << 3) | ((z & 7) << 6); }

// Octree::get_fine (octree.hpp:356-377): initValue() where nothing is allocated
template <class V>
__device__ __forceinline__ V get_fine(const MapView<V>& m, BlockCache& c, int x, int y, int z) {
  const int b = fetch_block_cached(m, c, x, y, z);
  if (b < 0) return FieldTraits<V>::init();
  return load_voxel(m.block_data + (size_t)b * kBlockVoxels + voxel_offset<V>(x, y, z));
}
template <class V>
__device__ __forceinline__ float get_fine_x(const MapView<V>& m, BlockCache& c, int x, int y, int z) {
  const int b = fetch_block_cached(m, c, x, y, z);
  if (b < 0) return FieldTraits<V>::init().x;
  return load_x(m.block_data + (size_t)b * kBlockVoxels + voxel_offset<V>(x, y, z));
}

// gather_points (interp_gather.hpp:105-237): the 8 corners are grouped by the block they fall
// in; one fetch per group; a missing block reads empty() in cases 0..6 and initValue() in the
// all-axes-crossing case 7 -- the same number for both field types (volume_traits.hpp:41-72), and what the pool's
// never-allocated payload holds.  The corners span at most two blocks per axis: corner s = (ox, oy, oz) lies in the
// block at directory cell  cell0 + (ox & cx) + G (oy & cy) + G^2 (oz & cz),  c* = "the base voxel is the last of its block
// along *".
//
// gather_points_general: any position, with or without the directory (volume faces, SE_B200_DISABLE_DIRECTORY).
template <class V>
__device__ __forceinline__ void gather_points_general(const MapView<V>& m, int bx, int by, int bz, float p[8]) {
  const bool cx = (bx & 7) == 7, cy = (by & 7) == 7, cz = (bz & 7) == 7;
  const int Bx = bx >> 3, By = by >> 3, Bz = bz >> 3;
  auto fetch = [&](int ox, int oy, int oz) -> int { const int b = fetch_block_cell(m, Bx + ox, By + oy, Bz + oz); return b < 0 ? m.max_blocks : b; };
  int id[8];
  id[0] = fetch(0, 0, 0);
  id[1] = cx ? fetch(1, 0, 0) : id[0];
  id[2] = cy ? fetch(0, 1, 0) : id[0];
  id[3] = cx ? (cy ? fetch(1, 1, 0) : id[1]) : id[2];
  id[4] = cz ? fetch(0, 0, 1) : id[0];
  id[5] = cx ? (cz ? fetch(1, 0, 1) : id[1]) : id[4];
  id[6] = cy ? (cz ? fetch(0, 1, 1) : id[2]) : id[4];
  id[7] = cx ? (cy ? (cz ? fetch(1, 1, 1) : id[3]) : id[5]) : id[6];
  const int xo[2] = { bx & 7, (bx + 1) & 7 };
  const int yo[2] = { (by & 7) << 3, ((by + 1) & 7) << 3 };
  const int zo[2] = { (bz & 7) << 6, ((bz + 1) & 7) << 6 };
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int off = xo[i & 1] + yo[(i >> 1) & 1] + zo[(i >> 2) & 1];
    p[i] = load_x(m.block_data + (size_t)id[i] * kBlockVoxels + off);
  }
}

// The usual case -- the base voxel and its +1 neighbours inside the volume, directory present -- straight-line: the seven
// neighbour-block ids are PREDICATED loads from the directory (a lane whose corner stays in the base block issues no
// memory operation and keeps the base id).  A warp almost always holds lanes of every crossing case, so branches on the
// case would make every warp run every branch; predication costs one issue slot per possible neighbour instead.
template <class V>
__device__ __forceinline__ bool gather_is_interior(const MapView<V>& m, int bx, int by, int bz) {
  const unsigned lim = (unsigned)(m.size - 1);
  return m.dir && (((unsigned)bx < lim) & ((unsigned)by < lim) & ((unsigned)bz < lim));
}
template <class V>
__device__ __forceinline__ void gather_points(const MapView<V>& m, BlockCache& c, int bx, int by, int bz, float p[8]) {
  const int G = m.dir_dim;
  const bool cx = (bx & 7) == 7, cy = (by & 7) == 7, cz = (bz & 7) == 7;
  const int b0 = fetch_block_cached(m, c, bx, by, bz);
  const int id0 = b0 < 0 ? m.max_blocks : b0;                 // unallocated -> the initValue() payload
  const int* cell0 = m.dir + (((bz >> 3) * G + (by >> 3)) * G + (bx >> 3));
  const int ox = cx ? 1 : 0, oy = cy ? G : 0, oz = cz ? G * G : 0;
  int id[8];
  id[0] = id0;
  id[1] = ldg_if(cx, cell0 + ox, b0);
  id[2] = ldg_if(cy, cell0 + oy, b0);
  id[3] = ldg_if(cx | cy, cell0 + (ox + oy), b0);
  id[4] = ldg_if(cz, cell0 + oz, b0);
  id[5] = ldg_if(cx | cz, cell0 + (ox + oz), b0);
  id[6] = ldg_if(cy | cz, cell0 + (oy + oz), b0);
  id[7] = ldg_if(cx | cy | cz, cell0 + (ox + oy + oz), b0);
  const int xo[2] = { bx & 7, (bx + 1) & 7 };
  const int yo[2] = { (by & 7) << 3, ((by + 1) & 7) << 3 };
  const int zo[2] = { (bz & 7) << 6, ((bz + 1) & 7) << 6 };
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int off = xo[i & 1] + yo[(i >> 1) & 1] + zo[(i >> 2) & 1];
    const int b = i == 0 ? id0 : (int)min((unsigned)id[i], (unsigned)m.max_blocks);      // kEmpty (-1) -> the initValue() payload
    p[i] = load_x(m.block_data + (size_t)b * kBlockVoxels + off);
  }
}

// Octree::interp (octree.hpp:541-563), pos in voxel units
__device__ __forceinline__ float trilinear(const float (&p)[8], float fx, float fy, float fz) {
  return (((p[0] * (1 - fx) + p[1] * fx) * (1 - fy)
         + (p[2] * (1 - fx) + p[3] * fx) * fy) * (1 - fz)
        + ((p[4] * (1 - fx) + p[5] * fx) * (1 - fy)
         + (p[6] * (1 - fx) + p[7] * fx) * fy) * fz);
}
// positions on the volume's faces / maps without the directory: out of line (rays almost never sample there)
template <class V>
__device__ __noinline__ float interp_field_general(const MapView<V> m, int bx, int by, int bz, float fx, float fy, float fz) {
  float p[8];
  gather_points_general(m, bx, by, bz, p);
  return trilinear(p, fx, fy, fz);
}
template <class V>
__device__ __forceinline__ float interp_field(const MapView<V>& m, BlockCache& c, V3 pos) {
  const float flx = floorf(pos.x), fly = floorf(pos.y), flz = floorf(pos.z);
  const float fx = pos.x - flx, fy = pos.y - fly, fz = pos.z - flz;
  const int bx = max((int)flx, 0), by = max((int)fly, 0), bz = max((int)flz, 0);
  if (!gather_is_interior(m, bx, by, bz)) return interp_field_general(m, bx, by, bz, fx, fy, fz);
  float p[8];
  gather_points(m, c, bx, by, bz, p);
  return (((p[0] * (1 - fx) + p[1] * fx) * (1 - fy)
         + (p[2] * (1 - fx) + p[3] * fx) * fy) * (1 - fz)
        + ((p[4] * (1 - fx) + p[5] * fx) * (1 - fy)
         + (p[6] * (1 - fx) + p[7] * fx) * fy) * fz);
}

// Octree::grad(pos, select) (octree.hpp:652-737): central differences blended trilinearly.
// The 48 reads of the reference expression touch 32 distinct voxels: per axis the clamped
// coordinates {ll, lu, ul, uu} = {max(b-1,0), max(b,0), min(b+1,hi), min(b+2,hi)}, which span at most
// two blocks per axis.  Same values and the same float expression as the reference -- only the addressing
// differs.  The samples, by their per-axis coordinate indices (jx, jy, jz) in 0..3 (lower = 1, upper = 2):
//   GX(jx, y, z)  jx = 0..3, y, z in {1,2}   the four x rows of the inner 2x2 columns (16, includes the inner 2x2x2)
//   GY(x, e, z)   jy = 0 | 3 (e = 0 | 1), x, z in {1,2}                                  (8)
//   GZ(x, y, e)   jz = 0 | 3 (e = 0 | 1), x, y in {1,2}                                  (8)
enum { kGradSamples = 32 };
#define SE_GX(JX, Y, Z) ((JX) + 4 * ((Y) - 1) + 8 * ((Z) - 1))
#define SE_GY(X, E, Z) (16 + ((X) - 1) + 2 * (E) + 4 * ((Z) - 1))
#define SE_GZ(X, Y, E) (24 + ((X) - 1) + 2 * ((Y) - 1) + 4 * (E))

// the blend of octree.hpp:669-733 over the 32 samples
__device__ __forceinline__ V3 grad_blend(const float (&g)[kGradSamples], float wx1, float wy1, float wz1, float scale) {
  const float wx0 = 1 - wx1, wy0 = 1 - wy1, wz0 = 1 - wz1;
  V3 r;
  {
    // gradient(0): octree.hpp:669-689
    const float t00 = (g[SE_GX(2, 1, 1)] - g[SE_GX(0, 1, 1)]) * wx0 + (g[SE_GX(3, 1, 1)] - g[SE_GX(1, 1, 1)]) * wx1;
    const float t10 = (g[SE_GX(2, 2, 1)] - g[SE_GX(0, 2, 1)]) * wx0 + (g[SE_GX(3, 2, 1)] - g[SE_GX(1, 2, 1)]) * wx1;
    const float t01 = (g[SE_GX(2, 1, 2)] - g[SE_GX(0, 1, 2)]) * wx0 + (g[SE_GX(3, 1, 2)] - g[SE_GX(1, 1, 2)]) * wx1;
    const float t11 = (g[SE_GX(2, 2, 2)] - g[SE_GX(0, 2, 2)]) * wx0 + (g[SE_GX(3, 2, 2)] - g[SE_GX(1, 2, 2)]) * wx1;
    r.x = (t00 * wy0 + t10 * wy1) * wz0 + (t01 * wy0 + t11 * wy1) * wz1;
  }
  {
    // gradient(1): octree.hpp:691-711: (v(x,ul,z) - v(x,ll,z)), (v(x,uu,z) - v(x,lu,z)) at x, z in {lower, upper}
    const float t00 = (g[SE_GX(1, 2, 1)] - g[SE_GY(1, 0, 1)]) * wx0 + (g[SE_GX(2, 2, 1)] - g[SE_GY(2, 0, 1)]) * wx1;
    const float t10 = (g[SE_GY(1, 1, 1)] - g[SE_GX(1, 1, 1)]) * wx0 + (g[SE_GY(2, 1, 1)] - g[SE_GX(2, 1, 1)]) * wx1;
    const float t01 = (g[SE_GX(1, 2, 2)] - g[SE_GY(1, 0, 2)]) * wx0 + (g[SE_GX(2, 2, 2)] - g[SE_GY(2, 0, 2)]) * wx1;
    const float t11 = (g[SE_GY(1, 1, 2)] - g[SE_GX(1, 1, 2)]) * wx0 + (g[SE_GY(2, 1, 2)] - g[SE_GX(2, 1, 2)]) * wx1;
    r.y = (t00 * wy0 + t10 * wy1) * wz0 + (t01 * wy0 + t11 * wy1) * wz1;
  }
  {
    // gradient(2): octree.hpp:713-733: (v(x,y,ul) - v(x,y,ll)), (v(x,y,uu) - v(x,y,lu)) at x, y in {lower, upper}
    const float t00 = (g[SE_GX(1, 1, 2)] - g[SE_GZ(1, 1, 0)]) * wx0 + (g[SE_GX(2, 1, 2)] - g[SE_GZ(2, 1, 0)]) * wx1;
    const float t10 = (g[SE_GX(1, 2, 2)] - g[SE_GZ(1, 2, 0)]) * wx0 + (g[SE_GX(2, 2, 2)] - g[SE_GZ(2, 2, 0)]) * wx1;
    const float t01 = (g[SE_GZ(1, 1, 1)] - g[SE_GX(1, 1, 1)]) * wx0 + (g[SE_GZ(2, 1, 1)] - g[SE_GX(2, 1, 1)]) * wx1;
    const float t11 = (g[SE_GZ(1, 2, 1)] - g[SE_GX(1, 2, 1)]) * wx0 + (g[SE_GZ(2, 2, 1)] - g[SE_GX(2, 2, 1)]) * wx1;
    r.z = (t00 * wy0 + t10 * wy1) * wz0 + (t01 * wy0 + t11 * wy1) * wz1;
  }
  return v3(scale * r.x, scale * r.y, scale * r.z);
}

// The general gather, for positions on or beyond the volume's faces (some clamped coordinate lies outside
// [0, hi]: the reference clamps only one side of each -- max(b,0) can exceed hi, min(b+1,hi) can be negative --
// and then reads out of bounds; here such a sample reads initValue(), like get_fine on an unallocated block)
// and for pools too large for 32-bit voxel indices.  Out of line: rays almost never end there.
template <class V>
__device__ __noinline__ V3 grad_field_general(const MapView<V> m, int b0, int b1, int b2, float wx1, float wy1, float wz1, float scale) {
  float g[kGradSamples];
  const int hi = m.size - 1;
  const int x4[4] = { max(b0 - 1, 0), max(b0, 0), min(b0 + 1, hi), min(b0 + 2, hi) };
  const int y4[4] = { max(b1 - 1, 0), max(b1, 0), min(b1 + 1, hi), min(b1 + 2, hi) };
  const int z4[4] = { max(b2 - 1, 0), max(b2, 0), min(b2 + 1, hi), min(b2 + 2, hi) };
  const float initx = FieldTraits<V>::init().x;
  auto sample = [&](int jx, int jy, int jz) -> float {
    const int x = x4[jx], y = y4[jy], z = z4[jz];
    if (!in_volume(m, x, y, z)) return initx;
    const int id = fetch_block(m, x, y, z);
    return id < 0 ? initx : load_x(m.block_data + (size_t)id * kBlockVoxels + voxel_offset<V>(x, y, z));
  };
  for (int z = 1; z <= 2; ++z)
    for (int y = 1; y <= 2; ++y)
      for (int jx = 0; jx < 4; ++jx) g[SE_GX(jx, y, z)] = sample(jx, y, z);
  for (int z = 1; z <= 2; ++z)
    for (int e = 0; e < 2; ++e)
      for (int x = 1; x <= 2; ++x) g[SE_GY(x, e, z)] = sample(x, 3 * e, z);
  for (int e = 0; e < 2; ++e)
    for (int y = 1; y <= 2; ++y)
      for (int x = 1; x <= 2; ++x) g[SE_GZ(x, y, e)] = sample(x, y, 3 * e);
  return grad_blend(g, wx1, wy1, wz1, scale);
}

// The usual case -- the whole 4x4x4 neighbourhood inside the volume -- in 12 rows: a row is the run of samples
// that differ only in x; its (at most two) blocks come as one int2 from a per-thread shared-memory column
// (`pairs[sy + 2 sz][thread]` = block ids at x-block 0 and 1, unallocated ones replaced by the pool's
// never-allocated initValue() block, so no sample needs a validity test), and every sample then is one select,
// one add, one load.
template <class V>
__device__ __forceinline__ V3 grad_field(const MapView<V>& m, int2 (*pairs)[/*threads*/ 128], V3 pos) {
  const float flx = floorf(pos.x), fly = floorf(pos.y), flz = floorf(pos.z);
  const int b0 = (int)flx, b1 = (int)fly, b2 = (int)flz;
  const float scale = m.grad_scale;
  const int hi = m.size - 1;
  // every clamped coordinate inside [0, hi]  <=>  -1 <= b <= hi on each axis
  const bool inside = ((unsigned)(b0 + 1) <= (unsigned)(hi + 1)) & ((unsigned)(b1 + 1) <= (unsigned)(hi + 1)) & ((unsigned)(b2 + 1) <= (unsigned)(hi + 1));
  typedef unsigned VoxelIndex;      // block * 512 + offset, unsigned: pools up to 2^23 - 1 blocks (32 GB of SDF payload) stay on this path
  const int kIndexablePool = (1 << 23) - 1;
  if (!inside || m.max_blocks > kIndexablePool) return grad_field_general(m, b0, b1, b2, pos.x - flx, pos.y - fly, pos.z - flz, scale);
  float g[kGradSamples];
  const int x4[4] = { max(b0 - 1, 0), max(b0, 0), min(b0 + 1, hi), min(b0 + 2, hi) };
  const int y4[4] = { max(b1 - 1, 0), max(b1, 0), min(b1 + 1, hi), min(b1 + 2, hi) };
  const int z4[4] = { max(b2 - 1, 0), max(b2, 0), min(b2 + 1, hi), min(b2 + 2, hi) };
  const int Bx = x4[0] >> 3, By = y4[0] >> 3, Bz = z4[0] >> 3;      // the coordinates are non-decreasing: lowest block per axis
  int sx[4], ox[4], ry[4], oy[4], rz[4], oz[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sx[j] = (x4[j] >> 3) - Bx;        ox[j] = x4[j] & 7;
    ry[j] = (y4[j] >> 3) - By;        oy[j] = (y4[j] & 7) << 3;
    rz[j] = ((z4[j] >> 3) - Bz) << 1; oz[j] = (z4[j] & 7) << 6;
  }
  const int t = threadIdx.x;
  // The 2x2x2 directory cells from one base index.  (Bx, By, Bz) is inside the grid here, only the +1 cells can fall
  // outside it: such a cell re-reads an inside one and the value is discarded.  (Eight generic lookups cost ~31
  // instructions each -- three bounds tests, the index arithmetic, the directory test; this is ~8 per cell.
  // Measured on the device, round 2: raycast 42.0 -> 41.4 us at 512^3, 130 -> 90 us at 2048^3 with its 4 M-block pool.)
  if (m.dir) {
    const int G = m.dir_dim;
    const bool ux = Bx + 1 < G, uy = By + 1 < G, uz = Bz + 1 < G;
    const int base = (Bz * G + By) * G + Bx;
    const int dxo = ux ? 1 : 0, dyo = uy ? G : 0, dzo = uz ? G * G : 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int cell = base + ((r & 1) ? dyo : 0) + ((r >> 1) ? dzo : 0);
      const bool row_ok = ((r & 1) ? uy : true) & ((r >> 1) ? uz : true);
      const int lo = __ldg(m.dir + cell), up = __ldg(m.dir + cell + dxo);
      pairs[r][t] = make_int2((row_ok && lo >= 0) ? lo : m.max_blocks, (row_ok && ux && up >= 0) ? up : m.max_blocks);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int lo = fetch_block_cell(m, Bx, By + (r & 1), Bz + (r >> 1)), up = fetch_block_cell(m, Bx + 1, By + (r & 1), Bz + (r >> 1));
      pairs[r][t] = make_int2(lo < 0 ? m.max_blocks : lo, up < 0 ? m.max_blocks : up);
    }
  }
  struct Row { VoxelIndex lo, up; };
  auto row = [&](int jy, int jz) {
    const int2 pr = pairs[ry[jy] + rz[jz]][t];
    const int o = oy[jy] + oz[jz];
    Row r; r.lo = (VoxelIndex)pr.x * kBlockVoxels + o; r.up = (VoxelIndex)pr.y * kBlockVoxels + o;
    return r;
  };
  auto S = [&](const Row& r, int jx) { return load_x(m.block_data + ((sx[jx] ? r.up : r.lo) + ox[jx])); };
#pragma unroll
  for (int z = 1; z <= 2; ++z)
#pragma unroll
    for (int y = 1; y <= 2; ++y) {
      const Row r = row(y, z);
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) g[SE_GX(jx, y, z)] = S(r, jx);
    }
#pragma unroll
  for (int z = 1; z <= 2; ++z)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const Row r = row(3 * e, z);
      g[SE_GY(1, e, z)] = S(r, 1); g[SE_GY(2, e, z)] = S(r, 2);
    }
#pragma unroll
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int y = 1; y <= 2; ++y) {
      const Row r = row(y, 3 * e);
      g[SE_GZ(1, y, e)] = S(r, 1); g[SE_GZ(2, y, e)] = S(r, 2);
    }
  return grad_blend(g, pos.x - flx, pos.y - fly, pos.z - flz, scale);
}

// VolumeTemplate::{get,interp,grad} (se_denseslam/include/se/continuous/volume_template.hpp:77-102):
// metres -> voxels by size/dim; get truncates toward zero, interp/grad floor.
template <class V>
__device__ __forceinline__ V vol_get(const MapView<V>& m, BlockCache& c, V3 p) {
  c.n_get++;
  const float inv = m.inv_voxel;
  return get_fine(m, c, (int)(inv * p.x), (int)(inv * p.y), (int)(inv * p.z));
}
template <class V>
__device__ __forceinline__ float vol_interp(const MapView<V>& m, BlockCache& c, V3 p) {
  c.n_interp++;
  const float inv = m.inv_voxel;
  return interp_field(m, c, v3(inv * p.x, inv * p.y, inv * p.z));
}
// (Measured on the device, round 2: issuing get(p)'s load and the eight corner loads interp(p) may need right after it in ONE
// round -- the march of raycast() calls both at the same point -- LOSES: raycast 32.6 -> 35.8 us at 512^3, 172 -> 198 us
// for OFusion 1024^3.  The extra address arithmetic and loads of the samples that need no interpolation cost more than
// the saved round trip.)
template <class V>
__device__ __forceinline__ V3 vol_grad(const MapView<V>& m, BlockCache& c, int2 (*ids)[128], V3 p) {
  c.n_grad++;
  const float inv = m.inv_voxel;
  return grad_field(m, ids, v3(inv * p.x, inv * p.y, inv * p.z));
}

// ---- insertion -----------------------------------------------------------------------------
// Find-or-create the octant `code` at `target_level`, creating the path from the root:
// Octree::allocate_level's walk (octree.hpp:819-856) done by whoever gets there first.
// A missing child slot is claimed with atomicCAS(kEmpty -> kBusy); the winner takes an index
// from the bump allocator, fills the metadata, and publishes the index with a fence; losers
// re-read the slot until it is published.  The winner never waits on anyone, so the scheme is
// starvation-free under independent thread scheduling.  A new block is also entered in the
// directory (after it is published in the tree; a reader that still sees kEmpty there falls back
// to this walk, which finds the block).
// Returns the node/block index, or kEmpty when a pool is exhausted (error bit set);
// created_target tells whether this call created the octant at target_level itself.
template <class V>
__device__ __forceinline__ int find_or_create(const MapView<V>& m, unsigned long long code, int target_level, bool& created_target) {
  created_target = false;
  int n = 0;
  unsigned g = 0;                       // heap index of node n (MapView::cmask)
  int edge = m.size >> 1;
  for (int level = 1; level <= target_level; ++level, edge >>= 1) {
    const int slot = key_child_id(code, level, m.max_level);
    int* p = m.node_child + 8 * n + slot;
    // First read through L1: published indices never change, and a stale kEmpty/kBusy is re-validated
    // by the atomicCAS / volatile re-read below.  (Reading the hot upper levels through L2 only made
    // every walk of every warp queue on the same few L2 lines.)
    int c = __ldca(p);
    while (c < 0) {
      if (c == kEmpty) {
        const int old = atomicCAS(p, kEmpty, kBusy);
        if (old == kEmpty) {
          const unsigned long long prefix = code & level_mask(kMaxBits - m.max_level + level - 1);
          int idx;
          if (level == m.leaves_level) {
            idx = atomicAdd(m.counters + kCntBlocks, 1);
            if (idx >= m.max_blocks) {
              atomicSub(m.counters + kCntBlocks, 1);
              atomicOr(m.counters + kCntError, kErrBlockPoolFull);
              atomicExch(p, kEmpty);
              return kEmpty;
            }
            int x, y, z;
            morton_decode(prefix, x, y, z);
            m.block_code[idx] = prefix | (unsigned long long)level;
            m.block_coord[idx] = make_int4(x, y, z, 0);
            m.block_active[idx] = 1;
          } else {
            idx = atomicAdd(m.counters + kCntNodes, 1);
            if (idx >= m.max_nodes) {
              atomicSub(m.counters + kCntNodes, 1);
              atomicOr(m.counters + kCntError, kErrNodePoolFull);
              atomicExch(p, kEmpty);
              return kEmpty;
            }
            m.node_code[idx] = prefix | (unsigned long long)level;
            m.node_side[idx] = (unsigned)edge;
          }
          if (level == target_level) created_target = true;
          atomicOr(m.node_mask + n, 1u << slot);
          if (m.cmask) atomicOr(reinterpret_cast<unsigned*>(m.cmask) + (g >> 2), (1u << slot) << ((g & 3u) * 8u));
          __threadfence();
          atomicExch(p, idx);
          if (level == m.leaves_level && m.dir) {
            int x, y, z;
            morton_decode(prefix, x, y, z);
            atomicExch(m.dir + ((z >> 3) * m.dir_dim + (y >> 3)) * m.dir_dim + (x >> 3), idx);
          } else if (level < m.leaves_level && m.ndir) {
            int x, y, z;
            morton_decode(prefix, x, y, z);
            atomicExch(m.ndir + node_dir_index(m, x, y, z, level), idx);
          }
          c = idx;
        } else {
          c = old;
        }
      } else {
        c = *((volatile int*)p);     // kBusy: someone is publishing this slot
      }
    }
    n = c;
    g = 8u * g + 1u + (unsigned)slot;
  }
  return n;
}

#endif  // __CUDACC__
}  // namespace se_b200
