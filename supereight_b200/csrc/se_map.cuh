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

// kCntActive0/1: the active-list length, double-buffered by frame parity (the list kernel of frame f counts into
// slot f&1 and clears slot (f+1)&1 for the next frame -- no separate reset launch).
enum Counter { kCntNodes = 0, kCntBlocks = 1, kCntError = 3, kCntNewBlocksBase = 4, kCntNewNodesBase = 5,
               kCntKeys = 6, kCntKeysReport = 7, kCntActive0 = 8, kCntActive1 = 9, kCntLastBlocks = 10, kCntLastNodes = 11,
               kNumCounters = 16 };
enum ErrorBits { kErrBlockPoolFull = 1, kErrNodePoolFull = 2, kErrKeyListFull = 4 };

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

template <class V> struct MapView {
  int size;            // voxels per side
  float dim;           // metres per side
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
  V* block_data;
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
};

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
// the tree descent itself, kept out of line: with the directory it is only the fallback
template <class V>
__device__ __noinline__ int fetch_block_tree(const MapView<V>& m, int x, int y, int z) {
  int n = 0;
  for (int edge = m.size >> 1; edge >= kBlockSide; edge >>= 1) {
    const int slot = ((x & edge) != 0) | (((y & edge) != 0) << 1) | (((z & edge) != 0) << 2);
    n = __ldg(m.node_child + 8 * n + slot);
    if (n < 0) return kEmpty;
  }
  return n;
}
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
// all-axes-crossing case 7.  The corners span at most two blocks per axis.  One straight-line path for
// all 8 crossing cases (a warp almost always holds both crossing and non-crossing lanes, so a separate
// fast path would only add its instructions to the slow one): corner s = (ox, oy, oz) lies in block
// id[s & crossmask]; the eight ids are built with selects, fetching only the blocks whose stepped axes all
// cross (usually none), and stay in registers.
template <class V>
__device__ __forceinline__ void gather_points(const MapView<V>& m, BlockCache& c, int bx, int by, int bz, float p[8]) {
  const bool cx = (bx & 7) == 7, cy = (by & 7) == 7, cz = (bz & 7) == 7;
  const int Bx = bx >> 3, By = by >> 3, Bz = bz >> 3;
  auto fetch = [&](int ox, int oy, int oz) -> int { return fetch_block_cell(m, Bx + ox, By + oy, Bz + oz); };
  int id[8];
  id[0] = fetch_block_cached(m, c, bx, by, bz);
  id[1] = cx ? fetch(1, 0, 0) : id[0];
  id[2] = cy ? fetch(0, 1, 0) : id[0];
  id[3] = cx ? (cy ? fetch(1, 1, 0) : id[1]) : id[2];
  id[4] = cz ? fetch(0, 0, 1) : id[0];
  id[5] = cx ? (cz ? fetch(1, 0, 1) : id[1]) : id[4];
  id[6] = cy ? (cz ? fetch(0, 1, 1) : id[2]) : id[4];
  id[7] = cx ? (cy ? (cz ? fetch(1, 1, 1) : id[3]) : id[5]) : id[6];
  const float missing = (cx & cy & cz) ? FieldTraits<V>::init().x : FieldTraits<V>::empty_x();
  const int xo[2] = { bx & 7, (bx + 1) & 7 };
  const int yo[2] = { (by & 7) << 3, ((by + 1) & 7) << 3 };
  const int zo[2] = { (bz & 7) << 6, ((bz + 1) & 7) << 6 };
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int off = xo[i & 1] + yo[(i >> 1) & 1] + zo[(i >> 2) & 1];
    p[i] = (id[i] < 0) ? missing : load_x(m.block_data + (size_t)id[i] * kBlockVoxels + off);
  }
}

// Octree::interp (octree.hpp:541-563), pos in voxel units
template <class V>
__device__ __forceinline__ float interp_field(const MapView<V>& m, BlockCache& c, V3 pos) {
  const float flx = floorf(pos.x), fly = floorf(pos.y), flz = floorf(pos.z);
  const float fx = pos.x - flx, fy = pos.y - fly, fz = pos.z - flz;
  const int bx = max((int)flx, 0), by = max((int)fly, 0), bz = max((int)flz, 0);
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
// two blocks per axis.  The (at most) 8 block indices are looked up once into `ids` (a per-thread
// column of shared memory); every sample then is: pick the block by three 0/1 selectors, one load.
// Same values and the same float expression as the reference -- only the addressing differs.
template <class V>
__device__ __forceinline__ V3 grad_field(const MapView<V>& m, int (*ids)[/*threads*/ 128], V3 pos) {
  const float flx = floorf(pos.x), fly = floorf(pos.y), flz = floorf(pos.z);
  const int b0 = (int)flx, b1 = (int)fly, b2 = (int)flz;
  const float wx1 = pos.x - flx, wy1 = pos.y - fly, wz1 = pos.z - flz;
  const float wx0 = 1 - wx1, wy0 = 1 - wy1, wz0 = 1 - wz1;
  const int hi = m.size - 1;
  // Packed per axis-coordinate code: (block selector 0/1) << 16|17|18, offset inside the block
  // pre-scaled in the low 9 bits, bit 28+ set when the coordinate lies outside [0, hi].  (The
  // reference clamps only one side of each coordinate -- max(b,0) can exceed hi, min(b+1,hi) can be
  // negative when the position is outside the volume -- and then reads out of bounds; here such a
  // sample reads initValue(), like get_fine on an unallocated block.)
  int cx[4], cy[4], cz[4];
  const int x4[4] = { max(b0 - 1, 0), max(b0, 0), min(b0 + 1, hi), min(b0 + 2, hi) };
  const int y4[4] = { max(b1 - 1, 0), max(b1, 0), min(b1 + 1, hi), min(b1 + 2, hi) };
  const int z4[4] = { max(b2 - 1, 0), max(b2, 0), min(b2 + 1, hi), min(b2 + 2, hi) };
  int Bx = 0x7fffffff, By = 0x7fffffff, Bz = 0x7fffffff;      // lowest block touched per axis
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if ((unsigned)x4[j] <= (unsigned)hi) Bx = min(Bx, x4[j] >> 3);
    if ((unsigned)y4[j] <= (unsigned)hi) By = min(By, y4[j] >> 3);
    if ((unsigned)z4[j] <= (unsigned)hi) Bz = min(Bz, z4[j] >> 3);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    cx[j] = (unsigned)x4[j] <= (unsigned)hi ? ((((x4[j] >> 3) - Bx) << 16) | (x4[j] & 7)) : (1 << 28);
    cy[j] = (unsigned)y4[j] <= (unsigned)hi ? ((((y4[j] >> 3) - By) << 17) | ((y4[j] & 7) << 3)) : (1 << 28);
    cz[j] = (unsigned)z4[j] <= (unsigned)hi ? ((((z4[j] >> 3) - Bz) << 18) | ((z4[j] & 7) << 6)) : (1 << 28);
  }
  const int t = threadIdx.x;
#pragma unroll
  for (int s = 0; s < 8; ++s) ids[s][t] = fetch_block_cell(m, Bx + (s & 1), By + ((s >> 1) & 1), Bz + (s >> 2));
  const float initx = FieldTraits<V>::init().x;
#define S(JX, JY, JZ) ([&]() { const int code = cx[JX] + cy[JY] + cz[JZ]; const int id = ids[(code >> 16) & 7][t]; \
                               return ((code >> 28) != 0 || id < 0) ? initx : load_x(m.block_data + (size_t)id * kBlockVoxels + (code & 0x1ff)); }())
  // the 32 distinct samples (indices into {ll, lu, ul, uu} per axis: 0..3; lower = 1, upper = 2)
  const float x_a00 = S(0, 1, 1), x_b00 = S(1, 1, 1), x_c00 = S(2, 1, 1), x_d00 = S(3, 1, 1);
  const float x_a10 = S(0, 2, 1), x_b10 = S(1, 2, 1), x_c10 = S(2, 2, 1), x_d10 = S(3, 2, 1);
  const float x_a01 = S(0, 1, 2), x_b01 = S(1, 1, 2), x_c01 = S(2, 1, 2), x_d01 = S(3, 1, 2);
  const float x_a11 = S(0, 2, 2), x_b11 = S(1, 2, 2), x_c11 = S(2, 2, 2), x_d11 = S(3, 2, 2);
  const float y_a00 = S(1, 0, 1), y_d00 = S(1, 3, 1), y_a10 = S(2, 0, 1), y_d10 = S(2, 3, 1);
  const float y_a01 = S(1, 0, 2), y_d01 = S(1, 3, 2), y_a11 = S(2, 0, 2), y_d11 = S(2, 3, 2);
  const float z_a00 = S(1, 1, 0), z_d00 = S(1, 1, 3), z_a10 = S(2, 1, 0), z_d10 = S(2, 1, 3);
  const float z_a01 = S(1, 2, 0), z_d01 = S(1, 2, 3), z_a11 = S(2, 2, 0), z_d11 = S(2, 2, 3);
#undef S
  // inner 2x2x2 (x index 1|2, y 1|2, z 1|2) by name: x_b/x_c rows above
  // v(x,y,z) with x,y,z in {1,2}:  v(1,1,1)=x_b00 v(2,1,1)=x_c00 v(1,2,1)=x_b10 v(2,2,1)=x_c10
  //                                v(1,1,2)=x_b01 v(2,1,2)=x_c01 v(1,2,2)=x_b11 v(2,2,2)=x_c11
  V3 r;
  {
    // gradient(0): octree.hpp:669-689
    const float t00 = (x_c00 - x_a00) * wx0 + (x_d00 - x_b00) * wx1;
    const float t10 = (x_c10 - x_a10) * wx0 + (x_d10 - x_b10) * wx1;
    const float t01 = (x_c01 - x_a01) * wx0 + (x_d01 - x_b01) * wx1;
    const float t11 = (x_c11 - x_a11) * wx0 + (x_d11 - x_b11) * wx1;
    r.x = (t00 * wy0 + t10 * wy1) * wz0 + (t01 * wy0 + t11 * wy1) * wz1;
  }
  {
    // gradient(1): octree.hpp:691-711 ; y index 0..3 at x in {1,2}
    const float t00 = (x_b10 - y_a00) * wx0 + (x_c10 - y_a10) * wx1;     // (v(lo,ul,lo) - v(lo,ll,lo)), (v(up,ul,lo) - v(up,ll,lo))
    const float t10 = (y_d00 - x_b00) * wx0 + (y_d10 - x_c00) * wx1;     // (v(lo,uu,lo) - v(lo,lu,lo)), (v(up,uu,lo) - v(up,lu,lo))
    const float t01 = (x_b11 - y_a01) * wx0 + (x_c11 - y_a11) * wx1;
    const float t11 = (y_d01 - x_b01) * wx0 + (y_d11 - x_c01) * wx1;
    r.y = (t00 * wy0 + t10 * wy1) * wz0 + (t01 * wy0 + t11 * wy1) * wz1;
  }
  {
    // gradient(2): octree.hpp:713-733 ; z index 0..3 at (x,y) in {1,2}^2
    const float t00 = (x_b01 - z_a00) * wx0 + (x_c01 - z_a10) * wx1;     // y = lo: (v(lo,lo,ul) - v(lo,lo,ll)), (v(up,lo,ul) - v(up,lo,ll))
    const float t10 = (x_b11 - z_a01) * wx0 + (x_c11 - z_a11) * wx1;     // y = up
    const float t01 = (z_d00 - x_b00) * wx0 + (z_d10 - x_c00) * wx1;     // y = lo: (v(lo,lo,uu) - v(lo,lo,lu)), ...
    const float t11 = (z_d01 - x_b10) * wx0 + (z_d11 - x_c10) * wx1;     // y = up
    r.z = (t00 * wy0 + t10 * wy1) * wz0 + (t01 * wy0 + t11 * wy1) * wz1;
  }
  const float s = (0.5f * m.dim) / (float)m.size;
  return v3(s * r.x, s * r.y, s * r.z);
}

// VolumeTemplate::{get,interp,grad} (se_denseslam/include/se/continuous/volume_template.hpp:77-102):
// metres -> voxels by size/dim; get truncates toward zero, interp/grad floor.
template <class V>
__device__ __forceinline__ V vol_get(const MapView<V>& m, BlockCache& c, V3 p) {
  c.n_get++;
  const float inv = (float)m.size / m.dim;
  return get_fine(m, c, (int)(inv * p.x), (int)(inv * p.y), (int)(inv * p.z));
}
template <class V>
__device__ __forceinline__ float vol_interp(const MapView<V>& m, BlockCache& c, V3 p) {
  c.n_interp++;
  const float inv = (float)m.size / m.dim;
  return interp_field(m, c, v3(inv * p.x, inv * p.y, inv * p.z));
}
template <class V>
__device__ __forceinline__ V3 vol_grad(const MapView<V>& m, BlockCache& c, int (*ids)[128], V3 p) {
  c.n_grad++;
  const float inv = (float)m.size / m.dim;
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
  }
  return n;
}

#endif  // __CUDACC__
}  // namespace se_b200
