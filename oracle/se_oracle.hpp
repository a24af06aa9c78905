// se_oracle.hpp -- CPU ORACLE for the supereight per-frame dense-SLAM hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may build,
// link or call anything under oracle/.  The product (supereight_b200/) never
// includes this file and never falls back to it.
//
// It is a dependency-free (no Eigen, no Sophus) restatement of the reference
// algorithm, SURVEY.md section 8(a) rows a1..a19.  Every function names the
// reference file:line (relative to /root/reference) whose behaviour it follows.
//
// PARITY STATUS: pinned to the reference's own code.
//  * oracle/_ref (oracle/Makefile, ref_capi.cpp) is the reference's own DenseSLAMSystem.cpp and everything it includes,
//    compiled where it lies under /root/reference, unmodified, against stand-in Eigen / Sophus headers
//    (oracle/ref_standin; the image has neither library).  tests/test_reference_build.py requires this oracle to be
//    BIT-IDENTICAL to that build for every array of every stage (allocation, integration, raycast, render, tracking,
//    SDF and OFusion), and tests/golden/seq_*.npz are written from it.
//  * The integer/structural layer is additionally pinned by the reference's own se_core GTest known-answer tests,
//    re-expressed in tests/test_oracle_kats.py.
//  * Not pinned: the few-ulp difference between the stand-in linear algebra (sums left to right, closed-form rigid /
//    camera inverses, SE3 as R p + t) and real Eigen / Sophus builds.  Where the reference leaves float evaluation order
//    to those libraries (K.inverse(), 4x4 products, the point transform, normalized()), this file and the stand-in use
//    the order of the "arithmetic contract" below, and the CUDA path follows the same contract, so that the GPU result
//    can be compared bit-for-bit.  Against a stock build these choices differ by a few ulp, inside the 1e-4 relative
//    tolerance north_star states.
//
// Arithmetic contract (shared by oracle and GPU, implemented independently):
//  * all float ops are IEEE-754 binary32, round-to-nearest, NO fused multiply-add
//    (build parity objects with -ffp-contract=off);
//  * dot/mat-vec rows are summed left to right: ((a0*b0 + a1*b1) + a2*b2) + a3*b3;
//  * K^-1 is the closed form of commons.h:264-271 (the reference's alloc path calls
//    the general Eigen K.inverse(), kfusion/alloc_impl.hpp:63);
//  * Tcw = inverse of the rigid pose: R^T and -(R^T t) (the reference goes through
//    Sophus::SE3f(pose).inverse(), i.e. a quaternion, DenseSLAMSystem.cpp:237);
//  * normalized(v) = v / sqrt((x*x + y*y) + z*z), component-wise division, identity
//    when the squared norm is 0 (Eigen semantics);
//  * float->int conversions truncate toward zero; where the reference's result is
//    undefined (inf/nan/out-of-range in filter.hpp:44-45) the value is "outside".
//
// Deliberate deviations from reference *undefined behaviour* (documented, tested):
//  * allocate(keys, 0) is a no-op (reference processes one stale key, unique.hpp:51-60);
//  * voxel coordinates outside [0,size) read as "missing block" instead of indexing
//    past a VoxelBlock array (octree.hpp:356-377 with x>=size);
//  * ray_iterator pop to scale 23 does not read stack[23] (ray_iterator.hpp:147-151).
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

namespace seo {

// ----------------------------------------------------------------------------
// small linear algebra with an explicit evaluation order
// ----------------------------------------------------------------------------
struct V3 { float x, y, z; };
struct V3i { int x, y, z; };
struct V4 { float x, y, z, w; };
struct M4 { float m[16]; float& at(int r, int c) { return m[4*r+c]; } float at(int r, int c) const { return m[4*r+c]; } };

static inline V3 operator+(V3 a, V3 b) { return {a.x+b.x, a.y+b.y, a.z+b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x-b.x, a.y-b.y, a.z-b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x*s, a.y*s, a.z*s}; }
static inline V3 operator*(float s, V3 a) { return {s*a.x, s*a.y, s*a.z}; }
static inline V3 operator/(V3 a, float s) { return {a.x/s, a.y/s, a.z/s}; }
static inline float dot3(V3 a, V3 b) { return (a.x*b.x + a.y*b.y) + a.z*b.z; }
static inline float sqnorm3(V3 a) { return dot3(a, a); }
static inline float norm3(V3 a) { return std::sqrt(sqnorm3(a)); }
static inline V3 normalized3(V3 a) { float n2 = sqnorm3(a); if (n2 > 0.f) return a / std::sqrt(n2); return a; }

static inline M4 mul44(const M4& A, const M4& B) {
  M4 C;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      C.at(i,j) = ((A.at(i,0)*B.at(0,j) + A.at(i,1)*B.at(1,j)) + A.at(i,2)*B.at(2,j)) + A.at(i,3)*B.at(3,j);
  return C;
}
// top-left 3x3 times vector
static inline V3 rot3(const M4& A, V3 v) {
  return { (A.at(0,0)*v.x + A.at(0,1)*v.y) + A.at(0,2)*v.z,
           (A.at(1,0)*v.x + A.at(1,1)*v.y) + A.at(1,2)*v.z,
           (A.at(2,0)*v.x + A.at(2,1)*v.y) + A.at(2,2)*v.z };
}
// top 3x4 times homogeneous point (w = 1)
static inline V3 xform3(const M4& A, V3 v) {
  return { ((A.at(0,0)*v.x + A.at(0,1)*v.y) + A.at(0,2)*v.z) + A.at(0,3),
           ((A.at(1,0)*v.x + A.at(1,1)*v.y) + A.at(1,2)*v.z) + A.at(1,3),
           ((A.at(2,0)*v.x + A.at(2,1)*v.y) + A.at(2,2)*v.z) + A.at(2,3) };
}
// inverse of a rigid transform; stands in for Sophus::SE3f(pose).inverse()
// (DenseSLAMSystem.cpp:237,249)
static inline M4 rigid_inverse(const M4& T) {
  M4 R;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.at(i,j) = T.at(j,i);
  for (int i = 0; i < 3; ++i)
    R.at(i,3) = -((T.at(0,i)*T.at(0,3) + T.at(1,i)*T.at(1,3)) + T.at(2,i)*T.at(2,3));
  R.at(3,0) = 0.f; R.at(3,1) = 0.f; R.at(3,2) = 0.f; R.at(3,3) = 1.f;
  return R;
}
// commons.h:255-262
static inline M4 camera_matrix(const float k[4]) {
  M4 K{}; K.at(0,0) = k[0]; K.at(0,2) = k[2]; K.at(1,1) = k[1]; K.at(1,2) = k[3]; K.at(2,2) = 1.f; K.at(3,3) = 1.f; return K;
}
// commons.h:264-271
static inline M4 inverse_camera_matrix(const float k[4]) {
  M4 K{}; K.at(0,0) = 1.0f / k[0]; K.at(0,2) = -k[2] / k[0]; K.at(1,1) = 1.0f / k[1]; K.at(1,2) = -k[3] / k[1]; K.at(2,2) = 1.f; K.at(3,3) = 1.f; return K;
}

// ----------------------------------------------------------------------------
// constants: constant_parameters.h:17-37, commons.h:71, octree_defines.h:38-41
// ----------------------------------------------------------------------------
constexpr float kNearPlane = 0.4f;
constexpr float kFarPlane = 4.0f;
constexpr int   kMaxWeight = 100;           // DenseSLAMSystem.cpp:235 passes 100
constexpr float kInvalid = -2.f;
constexpr float kAmbient = 0.1f;
constexpr int   kBlockSide = 8;
constexpr int   kMaxBits = 21;
constexpr int   kCastStackDepth = 23;
constexpr uint64_t kScaleMask = 0x1FFull;

// ----------------------------------------------------------------------------
// a5: key codec.  morton_utils.hpp:37-72, octant_ops.hpp:41-113, octree_defines.h:58-80
// ----------------------------------------------------------------------------
static inline uint64_t level_mask(int i) {
  // MASK[i] = MASK[i-1] | (MASK[0] >> 3i), MASK[0] = 0x7000000000000000 (octree_defines.h:49-57)
  uint64_t m = 0;
  for (int j = 0; j <= i; ++j) m |= (0x7000000000000000ull >> (3 * j));
  return m;
}
static inline uint64_t spread3(uint64_t v) {       // morton_utils.hpp:37-45
  uint64_t x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8)  & 0x100f00f00f00f00full;
  x = (x | x << 4)  & 0x10c30c30c30c30c3ull;
  x = (x | x << 2)  & 0x1249249249249249ull;
  return x;
}
static inline uint64_t squeeze3(uint64_t v) {      // morton_utils.hpp:47-55
  uint64_t x = v & 0x1249249249249249ull;
  x = (x | x >> 2)  & 0x10c30c30c30c30c3ull;
  x = (x | x >> 4)  & 0x100f00f00f00f00full;
  x = (x | x >> 8)  & 0x1f0000ff0000ffull;
  x = (x | x >> 16) & 0x1f00000000ffffull;
  x = (x | x >> 32) & 0x1fffffull;
  return x;
}
static inline uint64_t morton_encode(uint64_t x, uint64_t y, uint64_t z) { return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2); }
static inline V3i morton_decode(uint64_t c) { return { (int)squeeze3(c), (int)squeeze3(c >> 1), (int)squeeze3(c >> 2) }; }

static inline uint64_t key_code(uint64_t k) { return k & ~kScaleMask; }
static inline int key_level(uint64_t k) { return (int)(k & kScaleMask); }
static inline uint64_t key_encode(int x, int y, int z, int level, int max_depth) {    // octant_ops.hpp:49-53
  return (morton_encode((uint64_t)(int64_t)x, (uint64_t)(int64_t)y, (uint64_t)(int64_t)z) & level_mask(kMaxBits - max_depth + level - 1)) | (uint64_t)level;
}
static inline V3i key_decode(uint64_t k) { return morton_decode(k & ~kScaleMask); }
static inline bool key_descendant(uint64_t octant, uint64_t ancestor, int max_depth) { // octant_ops.hpp:81-88
  const int level = key_level(ancestor);
  const int idx = kMaxBits - max_depth + level - 1;
  ancestor = key_code(ancestor);
  octant = key_code(octant) & level_mask(idx);
  return (ancestor ^ octant) == 0;
}
static inline uint64_t key_parent(uint64_t octant, int max_depth) {                    // octant_ops.hpp:95-99
  const int level = key_level(octant) - 1;
  const int idx = kMaxBits - max_depth + level - 1;
  return (octant & level_mask(idx)) | (uint64_t)level;
}
static inline int key_child_id(uint64_t octant, int level, int max_depth) {            // octant_ops.hpp:107-113
  int shift = max_depth - level;
  octant = key_code(octant) >> (shift * 3);
  return (int)(octant & 7ull);
}
static inline V3i key_far_corner(uint64_t octant, int level, int max_depth) {          // octant_ops.hpp:121-129
  const int side = 1 << (max_depth - level);
  const int idx = key_child_id(octant, level, max_depth);
  const V3i c = key_decode(octant);
  return { c.x + (idx & 1) * side, c.y + ((idx & 2) >> 1) * side, c.z + ((idx & 4) >> 2) * side };
}
static inline V3i key_face_neighbour(uint64_t o, unsigned face, unsigned l, unsigned max_depth) { // octant_ops.hpp:63-72
  V3i c = key_decode(o);
  const int side = 1 << (max_depth - l);
  c.x += (face == 0) ? -side : (face == 1) ? side : 0;
  c.y += (face == 2) ? -side : (face == 3) ? side : 0;
  c.z += (face == 4) ? -side : (face == 5) ? side : 0;
  return c;
}
static inline void key_exterior_neighbours(uint64_t out[7], uint64_t octant, int level, int max_depth) { // octant_ops.hpp:141-167
  const int idx = key_child_id(octant, level, max_depth);
  int dx = (idx & 1) ? 1 : -1, dy = (idx & 2) ? 1 : -1, dz = (idx & 4) ? 1 : -1;
  const V3i b = key_far_corner(octant, level, max_depth);
  const int hi = (1 << max_depth) - 1;
  auto in = [&](int v) { return v >= 0 && v <= hi; };
  if (!in(b.x + dx)) dx = 0;
  if (!in(b.y + dy)) dy = 0;
  if (!in(b.z + dz)) dz = 0;
  out[0] = key_encode(b.x + dx, b.y,      b.z,      level, max_depth);
  out[1] = key_encode(b.x,      b.y + dy, b.z,      level, max_depth);
  out[2] = key_encode(b.x + dx, b.y + dy, b.z,      level, max_depth);
  out[3] = key_encode(b.x,      b.y,      b.z + dz, level, max_depth);
  out[4] = key_encode(b.x + dx, b.y,      b.z + dz, level, max_depth);
  out[5] = key_encode(b.x,      b.y + dy, b.z + dz, level, max_depth);
  out[6] = key_encode(b.x + dx, b.y + dy, b.z + dz, level, max_depth);
}
static inline void key_siblings(uint64_t out[8], uint64_t octant, int max_depth) {     // octant_ops.hpp:176-184
  const int level = key_level(octant);
  const int shift = 3 * (max_depth - level);
  const uint64_t p = key_parent(octant, max_depth) + 1;
  for (int i = 0; i < 8; ++i) out[i] = p | ((uint64_t)i << shift);
}

// a6 helpers: algorithms/unique.hpp:36-79
static inline int keys_unique(uint64_t* keys, int n) {
  int end = 1;
  if (n < 2) return end;
  for (int i = 1; i < n; ++i) if (keys[i] != keys[i-1]) keys[end++] = keys[i];
  return end;
}
static inline int keys_filter_ancestors(uint64_t* keys, int n, int max_depth) {
  int e = 0;
  for (int i = 0; i < n; ++i) {
    if (key_descendant(keys[i], keys[e], max_depth)) keys[e] = keys[i];
    else keys[++e] = keys[i];
  }
  return e + 1;
}
static inline int keys_unique_multiscale(uint64_t* keys, int n, unsigned current_level) {
  int e = 0;
  for (int i = 1; i < n; ++i) {
    const unsigned level = (unsigned)key_level(keys[i]);
    if (level >= current_level) {
      if (key_code(keys[i]) != key_code(keys[e])) keys[++e] = keys[i];
      else if (key_level(keys[i]) > key_level(keys[e])) keys[e] = keys[i];
    }
  }
  return e + 1;
}

// ----------------------------------------------------------------------------
// a7: field types.  volume_traits.hpp:41-81
// ----------------------------------------------------------------------------
struct SDF { float x; float y; };
struct OFusion { float x; double y; };
template <class F> struct traits;
template <> struct traits<SDF> {
  static SDF empty() { return {1.f, -1.f}; }
  static SDF init()  { return {1.f, 0.f}; }
  static constexpr bool is_sdf = true;
};
template <> struct traits<OFusion> {
  static OFusion empty() { return {0.f, 0.0}; }
  static OFusion init()  { return {0.f, 0.0}; }
  static constexpr bool is_sdf = false;
};

// node.hpp:45-137
template <class F> struct Node {
  F value[8];
  uint64_t code = 0;
  uint32_t side = 0;
  uint8_t children_mask = 0;
  int child[8];
  Node() { for (int i = 0; i < 8; ++i) { value[i] = traits<F>::init(); child[i] = -1; } }
};
template <class F> struct Block {
  uint64_t code = 0;
  uint32_t side = 0;
  V3i coords{0,0,0};
  bool active = false;        // reference leaves it uninitialised (node.hpp:132); every creation path sets true
  F data[512];
  Block() { for (int i = 0; i < 512; ++i) data[i] = traits<F>::init(); }
};

// utils/memory_pool.hpp:42-99 : paged pool, 1024 entries per page, stable addresses
template <class T> struct Pool {
  static constexpr int kPage = 1024;
  std::vector<std::unique_ptr<T[]>> pages;
  size_t count = 0;
  size_t size() const { return count; }
  T& operator[](size_t i) { return pages[i / kPage][i % kPage]; }
  const T& operator[](size_t i) const { return pages[i / kPage][i % kPage]; }
  size_t acquire() {
    if (count == pages.size() * (size_t)kPage) pages.emplace_back(new T[kPage]);
    return count++;
  }
};

struct Counters {   // for SURVEY 8(d) algorithmic bytes
  uint64_t n_get = 0, n_interp = 0, n_grad = 0;
  uint64_t n_active = 0, n_nodes = 0, n_new_blocks = 0, n_new_nodes = 0, n_unique_keys = 0, n_keys_raw = 0;
};

// ----------------------------------------------------------------------------
// a6/a17: the octree.  octree.hpp:88-273
// ----------------------------------------------------------------------------
template <class F> struct Octree {
  int size = 0;
  float dim = 0;
  int max_level = 0;
  int leaves_level = 0;
  Pool<Node<F>> nodes;      // nodes[0] is the root
  Pool<Block<F>> blocks;

  void init(int s, float d) {                     // octree.hpp:411-421
    size = s; dim = d;
    max_level = (int)std::log2((double)s);
    leaves_level = max_level - 3;
    size_t r = nodes.acquire();
    nodes[r].side = (uint32_t)s;
  }
  bool in_range(int x, int y, int z) const { return x >= 0 && y >= 0 && z >= 0 && x < size && y < size && z < size; }

  uint64_t hash(int x, int y, int z) const { return key_encode(x, y, z, leaves_level, max_level); }   // octree.hpp:205-208
  uint64_t hash(int x, int y, int z, int level) const { return key_encode(x, y, z, level, max_level); } // :210-212

  // octree.hpp:440-458 ; -1 == NULL
  int fetch(int x, int y, int z) const {
    if (!in_range(x, y, z)) return -1;
    int n = 0;
    unsigned edge = (unsigned)size / 2;
    for (; edge >= (unsigned)kBlockSide; edge /= 2) {
      n = nodes[n].child[((x & edge) > 0u) + 2 * ((y & edge) > 0u) + 4 * ((z & edge) > 0u)];
      if (n < 0) return -1;
    }
    return n;   // a block index
  }
  // octree.hpp:460-478 ; returns (index, is_block) ; index -1 == NULL
  int fetch_octant(int x, int y, int z, int depth, bool* is_block = nullptr) const {
    if (is_block) *is_block = false;
    if (!in_range(x, y, z)) return -1;
    int n = 0;
    unsigned edge = (unsigned)size / 2;
    for (int d = 1; edge >= (unsigned)kBlockSide && d <= depth; edge /= 2, ++d) {
      n = nodes[n].child[((x & edge) > 0u) + 2 * ((y & edge) > 0u) + 4 * ((z & edge) > 0u)];
      if (n < 0) return -1;
      if (is_block) *is_block = (edge == (unsigned)kBlockSide);
    }
    return n;
  }
  // octree.hpp:356-377
  F get_fine(int x, int y, int z) const {
    int b = fetch(x, y, z);
    if (b < 0) return traits<F>::init();
    const Block<F>& blk = blocks[b];
    return blk.data[(x - blk.coords.x) + (y - blk.coords.y) * 8 + (z - blk.coords.z) * 64];
  }
  // octree.hpp:332-354 : coarse get, returns the parent's value_ slot where the tree stops
  F get(int x, int y, int z) const {
    if (!in_range(x, y, z)) return traits<F>::init();
    int n = 0;
    unsigned edge = (unsigned)size >> 1;
    for (; edge >= (unsigned)kBlockSide; edge >>= 1) {
      const int id = ((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0);
      int c = nodes[n].child[id];
      if (c < 0) return nodes[n].value[id];
      n = c;
    }
    const Block<F>& blk = blocks[n];
    return blk.data[(x - blk.coords.x) + (y - blk.coords.y) * 8 + (z - blk.coords.z) * 64];
  }
  // octree.hpp:310-329
  void set(int x, int y, int z, F v) {
    int b = fetch(x, y, z);
    if (b < 0) return;
    Block<F>& blk = blocks[b];
    blk.data[(x - blk.coords.x) + (y - blk.coords.y) * 8 + (z - blk.coords.z) * 64] = v;
  }

  // octree.hpp:819-856 (serial here: keys at a level are unique, order only changes pool indices)
  void allocate_level(const uint64_t* keys, int num, int target_level, Counters* ctr) {
    for (int i = 0; i < num; ++i) {
      int n = 0;
      const uint64_t myKey = key_code(keys[i]);
      int edge = size / 2;
      for (int level = 1; level <= target_level; ++level) {
        const int index = key_child_id(myKey, level, max_level);
        const int parent = n;
        int c = nodes[parent].child[index];
        if (c < 0) {
          if (level == leaves_level) {
            c = (int)blocks.acquire();
            Block<F>& b = blocks[c];
            b.side = (uint32_t)edge;
            b.coords = morton_decode(myKey);
            b.active = true;
            b.code = myKey | (uint64_t)level;
            if (ctr) ctr->n_new_blocks++;
          } else {
            c = (int)nodes.acquire();
            Node<F>& nn = nodes[c];
            nn.code = myKey | (uint64_t)level;
            nn.side = (uint32_t)edge;
            if (ctr) ctr->n_new_nodes++;
          }
          nodes[parent].child[index] = c;
          nodes[parent].children_mask |= (uint8_t)(1 << index);
        }
        n = c;
        edge /= 2;
      }
    }
  }
  // octree.hpp:792-817
  bool allocate(uint64_t* keys, int num_elem, Counters* ctr = nullptr) {
    if (num_elem <= 0) return false;             // deviation: reference processes one stale key
#ifdef _OPENMP
    __gnu_parallel::sort(keys, keys + num_elem);
#else
    std::sort(keys, keys + num_elem);
#endif
    num_elem = keys_filter_ancestors(keys, num_elem, max_level);
    if (ctr) ctr->n_unique_keys += (uint64_t)num_elem;
    std::vector<uint64_t> at_level((size_t)num_elem);
    const unsigned shift = kMaxBits - max_level - 1;
    for (int level = 1; level <= leaves_level; ++level) {
      const uint64_t mask = level_mask(level + shift) | kScaleMask;
#pragma omp parallel for
      for (int i = 0; i < num_elem; ++i) at_level[i] = keys[i] & mask;     // morton_utils.hpp:74-81
      const int last = keys_unique_multiscale(at_level.data(), num_elem, (unsigned)level);
      allocate_level(at_level.data(), last, level, ctr);
    }
    return true;
  }

  // ---- a17: interpolation/interp_gather.hpp:105-237 + octree.hpp:541-563 -------------
  // select == .x for both field types on this path (rendering.cpp:75, *rendering_impl.hpp)
  float block_x(int b, int x, int y, int z, bool empty_if_missing) const {
    if (b < 0) return empty_if_missing ? traits<F>::empty().x : traits<F>::init().x;
    const Block<F>& blk = blocks[b];
    return blk.data[(x - blk.coords.x) + (y - blk.coords.y) * 8 + (z - blk.coords.z) * 64].x;
  }
  void gather_points(int bx, int by, int bz, float p[8]) const {
    static const int off[8][3] = {{0,0,0},{1,0,0},{0,1,0},{1,1,0},{0,0,1},{1,0,1},{0,1,1},{1,1,1}};
    const unsigned cross = ((unsigned)(bx % 8 == 7) << 2) | ((unsigned)(by % 8 == 7) << 1) | (unsigned)(bz % 8 == 7);
    if (cross == 7) {      // interp_gather.hpp:214-235 : eight get_fine() -> init value when missing
      for (int i = 0; i < 8; ++i) p[i] = get_fine(bx + off[i][0], by + off[i][1], bz + off[i][2]).x;
      return;
    }
    // cases 0..6 (:122-212): the corners are grouped by the block they fall in -- one fetch per
    // group, 1/2/4 groups for 0/1/2 crossing axes; a missing block reads empty().
    // A group is a subset g of the crossing axes; corner i belongs to the group made of the
    // crossing axes on which it steps over.
    for (unsigned g = cross;; g = (g - 1) & cross) {
      const int b = fetch(bx + (int)((g >> 2) & 1), by + (int)((g >> 1) & 1), bz + (int)(g & 1));
      for (int i = 0; i < 8; ++i) {
        const unsigned cb = ((unsigned)off[i][0] << 2) | ((unsigned)off[i][1] << 1) | (unsigned)off[i][2];
        if ((cb & cross) == g) p[i] = block_x(b, bx + off[i][0], by + off[i][1], bz + off[i][2], true);
      }
      if (g == 0) break;
    }
  }
  float interp(V3 pos, Counters* ctr = nullptr) const {        // octree.hpp:541-563
    if (ctr) ctr->n_interp++;
    const float fx = std::floor(pos.x), fy = std::floor(pos.y), fz = std::floor(pos.z);
    const V3 f = { pos.x - fx, pos.y - fy, pos.z - fz };
    const int bx = std::max((int)fx, 0), by = std::max((int)fy, 0), bz = std::max((int)fz, 0);
    float p[8];
    gather_points(bx, by, bz, p);
    return (((p[0] * (1 - f.x) + p[1] * f.x) * (1 - f.y)
           + (p[2] * (1 - f.x) + p[3] * f.x) * f.y) * (1 - f.z)
          + ((p[4] * (1 - f.x) + p[5] * f.x) * (1 - f.y)
           + (p[6] * (1 - f.x) + p[7] * f.x) * f.y) * f.z);
  }
  // octree.hpp:652-737 ; the `cached` block of the reference only short-cuts the descent,
  // every sample equals get_fine() at clamped coordinates
  V3 grad(V3 pos, Counters* ctr = nullptr) const {
    if (ctr) ctr->n_grad++;
    const float flx = std::floor(pos.x), fly = std::floor(pos.y), flz = std::floor(pos.z);
    const int b[3] = { (int)flx, (int)fly, (int)flz };
    const float f[3] = { pos.x - flx, pos.y - fly, pos.z - flz };
    int ll[3], lu[3], ul[3], uu[3];
    for (int a = 0; a < 3; ++a) {
      ll[a] = std::max(b[a] - 1, 0);
      lu[a] = std::max(b[a], 0);
      ul[a] = std::min(b[a] + 1, size - 1);
      uu[a] = std::min(b[a] + 2, size - 1);
    }
    auto g = [&](int x, int y, int z) { return get_fine(x, y, z).x; };
    const int* lo = lu; const int* up = ul;
    const float wx0 = 1 - f[0], wx1 = f[0], wy0 = 1 - f[1], wy1 = f[1], wz0 = 1 - f[2], wz1 = f[2];
    V3 r;
    {
      auto t = [&](int y, int z) { return (g(ul[0], y, z) - g(ll[0], y, z)) * wx0 + (g(uu[0], y, z) - g(lu[0], y, z)) * wx1; };
      r.x = (t(lo[1], lo[2]) * wy0 + t(up[1], lo[2]) * wy1) * wz0 + (t(lo[1], up[2]) * wy0 + t(up[1], up[2]) * wy1) * wz1;
    }
    {
      auto t = [&](int yh, int yl, int z) { return (g(lo[0], yh, z) - g(lo[0], yl, z)) * wx0 + (g(up[0], yh, z) - g(up[0], yl, z)) * wx1; };
      r.y = (t(ul[1], ll[1], lo[2]) * wy0 + t(uu[1], lu[1], lo[2]) * wy1) * wz0 + (t(ul[1], ll[1], up[2]) * wy0 + t(uu[1], lu[1], up[2]) * wy1) * wz1;
    }
    {
      auto t = [&](int y, int zh, int zl) { return (g(lo[0], y, zh) - g(lo[0], y, zl)) * wx0 + (g(up[0], y, zh) - g(up[0], y, zl)) * wx1; };
      r.z = (t(lo[1], ul[2], ll[2]) * wy0 + t(up[1], ul[2], ll[2]) * wy1) * wz0 + (t(lo[1], uu[2], lu[2]) * wy0 + t(up[1], uu[2], lu[2]) * wy1) * wz1;
    }
    const float s = (0.5f * dim) / (float)size;
    return { s * r.x, s * r.y, s * r.z };
  }
};

// continuous/volume_template.hpp:77-102 : metric -> voxel wrappers
template <class F> static inline F vol_get(const Octree<F>& o, V3 p, Counters* ctr) {
  if (ctr) ctr->n_get++;
  const float inv = (float)o.size / o.dim;
  return o.get_fine((int)(inv * p.x), (int)(inv * p.y), (int)(inv * p.z));
}
template <class F> static inline float vol_interp(const Octree<F>& o, V3 p, Counters* ctr) {
  const float inv = (float)o.size / o.dim;
  return o.interp({inv * p.x, inv * p.y, inv * p.z}, ctr);
}
template <class F> static inline V3 vol_grad(const Octree<F>& o, V3 p, Counters* ctr) {
  const float inv = (float)o.size / o.dim;
  return o.grad({inv * p.x, inv * p.y, inv * p.z}, ctr);
}

// ----------------------------------------------------------------------------
// a14: ray/octree traversal.  ray_iterator.hpp:49-289
// ----------------------------------------------------------------------------
static inline int f2i(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float i2f(int i) { float f; std::memcpy(&f, &i, 4); return f; }

template <class F> struct RayIterator {
  const Octree<F>& map;
  V3 dir, t_coef, t_bias, pos, t_corner;
  struct Entry { int parent; float t_max; } stack[kCastStackDepth];
  int parent, child, idx, scale, min_scale, octant_mask;
  float scale_exp2, t_min, t_min_init, t_max, t_max_init, tc_max, h;
  enum { INIT, ADVANCE, FINISHED } state;
  bool child_is_block = false;

  RayIterator(const Octree<F>& m, V3 origin, V3 direction, float nearP, float farP) : map(m) {
    pos = {1.f, 1.f, 1.f};
    idx = 0; parent = 0; child = -1;
    scale_exp2 = 0.5f;
    scale = kCastStackDepth - 1;
    min_scale = kCastStackDepth - (int)std::log2((double)(m.size / kBlockSide));
    const float eps = 1.0f / (float)m.size;          // exp2f(-log2(size)), size is a power of two
    state = INIT;
    for (auto& e : stack) e = {0, 0.f};
    auto fix = [&](float d) { return std::fabs(d) < eps ? std::copysign(eps, d) : d; };
    dir = { fix(direction.x), fix(direction.y), fix(direction.z) };
    const V3 so = { origin.x / m.dim + 1.f, origin.y / m.dim + 1.f, origin.z / m.dim + 1.f };
    t_coef = { -1.f * (1.f / std::fabs(dir.x)), -1.f * (1.f / std::fabs(dir.y)), -1.f * (1.f / std::fabs(dir.z)) };
    t_bias = { t_coef.x * so.x, t_coef.y * so.y, t_coef.z * so.z };
    octant_mask = 7;
    if (dir.x > 0.0f) { octant_mask ^= 1; t_bias.x = 3.0f * t_coef.x - t_bias.x; }
    if (dir.y > 0.0f) { octant_mask ^= 2; t_bias.y = 3.0f * t_coef.y - t_bias.y; }
    if (dir.z > 0.0f) { octant_mask ^= 4; t_bias.z = 3.0f * t_coef.z - t_bias.z; }
    t_min = std::fmax(std::fmax(2.0f * t_coef.x - t_bias.x, 2.0f * t_coef.y - t_bias.y), 2.0f * t_coef.z - t_bias.z);
    t_max = std::fmin(std::fmin(t_coef.x - t_bias.x, t_coef.y - t_bias.y), t_coef.z - t_bias.z);
    h = t_max;
    t_min = t_min_init = std::fmax(t_min, nearP / m.dim);
    t_max = t_max_init = std::fmin(t_max, farP / m.dim);
    if (1.5f * t_coef.x - t_bias.x > t_min) { idx ^= 1; pos.x = 1.5f; }
    if (1.5f * t_coef.y - t_bias.y > t_min) { idx ^= 2; pos.y = 1.5f; }
    if (1.5f * t_coef.z - t_bias.z > t_min) { idx ^= 4; pos.z = 1.5f; }
    tc_max = 0.f; t_corner = {0, 0, 0};
  }
  void advance_ray() {                                  // :116-167
    const int step_mask = (int)(t_corner.x <= tc_max) | ((int)(t_corner.y <= tc_max) << 1) | ((int)(t_corner.z <= tc_max) << 2);
    pos.x -= scale_exp2 * (float)((step_mask & 1) != 0);
    pos.y -= scale_exp2 * (float)((step_mask & 2) != 0);
    pos.z -= scale_exp2 * (float)((step_mask & 4) != 0);
    t_min = tc_max;
    idx ^= step_mask;
    if ((idx & step_mask) != 0) {                       // pop
      unsigned differing = 0;
      if (step_mask & 1) differing |= (unsigned)(f2i(pos.x) ^ f2i(pos.x + scale_exp2));
      if (step_mask & 2) differing |= (unsigned)(f2i(pos.y) ^ f2i(pos.y + scale_exp2));
      if (step_mask & 4) differing |= (unsigned)(f2i(pos.z) ^ f2i(pos.z + scale_exp2));
      scale = (f2i((float)differing) >> 23) - 127;
      scale_exp2 = i2f((scale - kCastStackDepth + 127) << 23);
      if (scale < kCastStackDepth) { parent = stack[scale].parent; t_max = stack[scale].t_max; }
      const int shx = f2i(pos.x) >> scale, shy = f2i(pos.y) >> scale, shz = f2i(pos.z) >> scale;
      pos.x = i2f(shx << scale); pos.y = i2f(shy << scale); pos.z = i2f(shz << scale);
      idx = (shx & 1) | ((shy & 1) << 1) | ((shz & 1) << 2);
      h = 0.0f;
      child = -1;
    }
  }
  void descend() {                                      // :172-199
    const float tv_max = std::fmin(t_max, tc_max);
    const float half = scale_exp2 * 0.5f;
    const V3 t_center = { half * t_coef.x + t_corner.x, half * t_coef.y + t_corner.y, half * t_coef.z + t_corner.z };
    if (tc_max < h) stack[scale] = { parent, t_max };
    h = tc_max;
    parent = child;
    idx = 0;
    scale--;
    scale_exp2 = half;
    idx ^= (t_center.x > t_min) ? 1 : 0;
    idx ^= (t_center.y > t_min) ? 2 : 0;
    idx ^= (t_center.z > t_min) ? 4 : 0;
    pos.x += scale_exp2 * (float)((idx & 1) != 0);
    pos.y += scale_exp2 * (float)((idx & 2) != 0);
    pos.z += scale_exp2 * (float)((idx & 4) != 0);
    t_max = tv_max;
    child = -1;
  }
  // :205-226 ; returns a block index or -1
  int next() {
    if (state == ADVANCE) advance_ray();
    else if (state == FINISHED) return -1;
    while (scale < kCastStackDepth) {
      t_corner = { pos.x * t_coef.x - t_bias.x, pos.y * t_coef.y - t_bias.y, pos.z * t_coef.z - t_bias.z };
      tc_max = std::fmin(std::fmin(t_corner.x, t_corner.y), t_corner.z);
      child = map.nodes[parent].child[idx ^ octant_mask ^ 7];
      if (scale == min_scale && child >= 0) { state = ADVANCE; return child; }
      else if (child >= 0 && t_min <= t_max) { descend(); continue; }
      advance_ray();
    }
    return -1;
  }
  float tmin() const { return t_min_init * map.dim; }
  float tmax() const { return t_max_init * map.dim; }
  float tcmin() const { return t_min * map.dim; }
  float tcmax() const { return tc_max * map.dim; }
};

// ----------------------------------------------------------------------------
// a15/a16: surface search along a ray
// ----------------------------------------------------------------------------
// kfusion/rendering_impl.hpp:34-74
static inline V4 raycast_field(const Octree<SDF>& vol, V3 origin, V3 direction, float tnear, float tfar,
                               float mu, float step, float largestep, Counters* ctr) {
  if (tnear < tfar) {
    float t = tnear;
    float stepsize = largestep;
    V3 position = origin + direction * t;
    float f_t = vol_interp(vol, position, ctr);
    float f_tt = 0;
    if (f_t > 0) {
      for (; t < tfar; t += stepsize) {
        SDF data = vol_get(vol, position, ctr);
        if (data.y == 0) {
          stepsize = largestep;
          position = position + stepsize * direction;
          continue;
        }
        f_tt = data.x;
        if ((double)f_tt <= 0.1 && f_tt >= -0.5f) f_tt = vol_interp(vol, position, ctr);
        if (f_tt < 0) break;
        stepsize = std::fmax(f_tt * mu, step);
        position = position + stepsize * direction;
        f_t = f_tt;
      }
      if (f_tt < 0) {
        t = t + stepsize * f_tt / (f_t - f_tt);
        const V3 p = origin + direction * t;
        return { p.x, p.y, p.z, t };
      }
    }
  }
  return {0, 0, 0, 0};
}
// bfusion/rendering_impl.hpp:35-68
static inline V4 raycast_field(const Octree<OFusion>& vol, V3 origin, V3 direction, float tnear, float tfar,
                               float /*mu*/, float step, float /*largestep*/, Counters* ctr) {
  if (tnear < tfar) {
    float t = tnear;
    const float stepsize = step;
    float f_t = vol_interp(vol, origin + direction * t, ctr);
    float f_tt = 0;
    if (f_t <= 0.f) {
      for (; t < tfar; t += stepsize) {
        const V3 pos = origin + direction * t;
        OFusion data = vol_get(vol, pos, ctr);
        if (data.x > -100.f && data.y > 0.0) f_tt = vol_interp(vol, origin + direction * t, ctr);
        if (f_tt > 0.f) break;
        f_t = f_tt;
      }
      if (f_tt > 0.f) {
        t = t - stepsize * (f_tt - 0.f) / (f_tt - f_t);
        const V3 p = origin + direction * t;
        return { p.x, p.y, p.z, t };
      }
    }
  }
  return {0, 0, 0, 0};
}

// ----------------------------------------------------------------------------
// a11: OFusion B-spline machinery.  bfusion/mapping_impl.hpp:94-155, bspline_lookup.cc:36-37
// The 1000-entry table is regenerated from the closed form it samples: bspline() at
// t = -3 + 6 i / 999 evaluated in double and rounded to float reproduces the reference's
// printed table bit for bit (tests/golden/bspline_lut.sha256 pins it).
// ----------------------------------------------------------------------------
static inline double bspline_closed(double t) {          // mapping_impl.hpp:94-106, in double
  double value = 0.0;
  if (t >= -3.0 && t <= -1.0) value = std::pow(3 + t, 3) / 48.0;
  else if (t > -1 && t <= 1) value = 0.5 + (t * (3 + t) * (3 - t)) / 24.0;
  else if (t > 1 && t <= 3) value = 1 - std::pow(3 - t, 3) / 48.0;
  else if (t > 3) value = 1.0;
  return value;
}
struct BsplineLut {
  float v[1000];
  BsplineLut();
  static const BsplineLut& get() { static BsplineLut l; return l; }
};
static inline float bspline_memoized(float t) {          // mapping_impl.hpp:126-137
  float value = 0.f;
  constexpr float inverseRange = 1 / 6.f;
  if (t >= -3.0f && t <= 3.0f) {
    unsigned idx = (unsigned)(((t + 3.f) * inverseRange) * (1000 - 1) + 0.5f);
    return BsplineLut::get().v[idx];
  } else if (t > 3) value = 1.f;
  return value;
}
static inline float h_new(float val) { return bspline_memoized(val) - bspline_memoized(val - 3) * 0.5f; } // :139-143

// ----------------------------------------------------------------------------
// N1: tracking front-end.  preprocessing.cpp:42-159,190-226, tracking.cpp:42-336,
// DenseSLAMSystem.cpp:143-189.  The float reductions of the reference are OpenMP reductions
// (order not defined); here they are serial per row block, so this oracle is deterministic and
// the GPU path is compared with a tolerance, not bit for bit.
// ----------------------------------------------------------------------------
struct TrackData { int result; float error; float J[6]; };      // commons.h:249-253

constexpr float kEDelta = 0.1f;              // constant_parameters.h:17
constexpr int   kRadius = 2;                 // :18
constexpr float kDistThreshold = 0.1f;       // :19
constexpr float kNormalThreshold = 0.8f;     // :20
constexpr float kTrackThreshold = 0.15f;     // :21
constexpr float kGaussDelta = 4.0f;          // :34

// DenseSLAMSystem.cpp:111-118
static inline void make_gaussian(float g[5]) { for (int i = 0; i < 5; ++i) { const int x = i - 2; g[i] = std::exp(-(float)(x * x) / (2 * kGaussDelta * kGaussDelta)); } }

// preprocessing.cpp:42-87
static inline void bilateral_filter(std::vector<float>& out, const std::vector<float>& in, int W, int H) {
  float g[5]; make_gaussian(g);
  const float e_d_squared_2 = kEDelta * kEDelta * 2;
  const int r = kRadius;
#pragma omp parallel for
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int pos = x + y * W;
      if (in[pos] == 0) { out[pos] = 0; continue; }
      float sum = 0.f, t = 0.f;
      const float center = in[pos];
      for (int i = -r; i <= r; ++i)
        for (int j = -r; j <= r; ++j) {
          const int cx = std::max(0, std::min(x + i, W - 1)), cy = std::max(0, std::min(y + j, H - 1));
          const float curPix = in[cx + cy * W];
          if (curPix > 0) {
            const float mod = (curPix - center) * (curPix - center);
            const float factor = g[i + r] * g[j + r] * std::exp(-mod / e_d_squared_2);
            t += factor * curPix;
            sum += factor;
          }
        }
      out[pos] = t / sum;
    }
}
// preprocessing.cpp:190-226 (r = 1, e_d = 3 * e_delta at the call site, DenseSLAMSystem.cpp:151)
// inW = in.width(): the parent level's own row stride (2 outW + 1 when its width is odd); the clamp stays at 2 out - 1 (:214-216)
static inline void half_sample_robust(std::vector<float>& out, const std::vector<float>& in, int outW, int outH, int inW, float e_d, int r) {
#pragma omp parallel for
  for (int y = 0; y < outH; ++y)
    for (int x = 0; x < outW; ++x) {
      const int cx = 2 * x, cy = 2 * y;
      float sum = 0.f, t = 0.f;
      const float center = in[cx + cy * inW];
      for (int i = -r + 1; i <= r; ++i)
        for (int j = -r + 1; j <= r; ++j) {
          const int px = std::min(std::max(cx + j, 0), 2 * outW - 1), py = std::min(std::max(cy + i, 0), 2 * outH - 1);
          const float current = in[px + py * inW];
          if (std::fabs(current - center) < e_d) { sum += 1.0f; t += current; }
        }
      out[x + y * outW] = t / sum;
    }
}
// preprocessing.cpp:89-109 : depth * invK * (x, y, 1, 0)
static inline void depth2vertex(std::vector<V3>& vertex, const std::vector<float>& depth, int W, int H, const M4& invK) {
#pragma omp parallel for
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float d = depth[x + y * W];
      if (d > 0) {
        // (depth * invK) * v : the scalar multiplies the matrix first (Eigen evaluates left to right)
        V3 r;
        const float vx = (float)x, vy = (float)y;
        r.x = ((d * invK.at(0,0)) * vx + (d * invK.at(0,1)) * vy) + (d * invK.at(0,2)) * 1.f;
        r.y = ((d * invK.at(1,0)) * vx + (d * invK.at(1,1)) * vy) + (d * invK.at(1,2)) * 1.f;
        r.z = ((d * invK.at(2,0)) * vx + (d * invK.at(2,1)) * vy) + (d * invK.at(2,2)) * 1.f;
        vertex[x + y * W] = r;
      } else vertex[x + y * W] = {0, 0, 0};
    }
}
static inline V3 cross3(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
// preprocessing.cpp:111-159 ; only .x is written for invalid pixels (the rest keeps its old value)
static inline void vertex2normal(std::vector<V3>& out, const std::vector<V3>& in, int W, int H, bool negY) {
#pragma omp parallel for
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const V3 center = in[x + W * y];
      if (center.z == 0.f) { out[x + y * W].x = kInvalid; continue; }
      const int xl = std::max(x - 1, 0), xr = std::min(x + 1, W - 1);
      int yu, yd;
      if (negY) { yu = std::max(y - 1, 0); yd = std::min(y + 1, H - 1); }
      else { yd = std::max(y - 1, 0); yu = std::min(y + 1, H - 1); }
      const V3 left = in[xl + W * y], right = in[xr + W * y], up = in[x + W * yu], down = in[x + W * yd];
      if (left.z == 0 || right.z == 0 || up.z == 0 || down.z == 0) { out[x + y * W].x = kInvalid; continue; }
      out[x + y * W] = normalized3(cross3(right - left, up - down));
    }
}
// tracking.cpp:226-300
static inline void track_kernel(TrackData* output, const std::vector<V3>& inVertex, const std::vector<V3>& inNormal, int inW, int inH,
                                const std::vector<V3>& refVertex, const std::vector<V3>& refNormal, int refW, int refH,
                                const M4& Ttrack, const M4& view, float dist_threshold, float normal_threshold) {
#pragma omp parallel for
  for (int py = 0; py < inH; ++py)
    for (int px = 0; px < inW; ++px) {
      TrackData& row = output[px + py * refW];
      const V3 inN = inNormal[px + py * inW];
      if (inN.x == kInvalid) { row.result = -1; continue; }
      const V3 projectedVertex = xform3(Ttrack, inVertex[px + py * inW]);
      const V3 projectedPos = xform3(view, projectedVertex);
      const float ppx = projectedPos.x / projectedPos.z + 0.5f, ppy = projectedPos.y / projectedPos.z + 0.5f;
      if (ppx < 0 || ppx > refW - 1 || ppy < 0 || ppy > refH - 1) { row.result = -2; continue; }
      const int rx = (int)ppx, ry = (int)ppy;
      const V3 referenceNormal = refNormal[rx + ry * refW];
      if (referenceNormal.x == kInvalid) { row.result = -3; continue; }
      const V3 diff = refVertex[rx + ry * refW] - projectedVertex;
      const V3 projectedNormal = rot3(Ttrack, inN);
      if (norm3(diff) > dist_threshold) { row.result = -4; continue; }
      if (dot3(projectedNormal, referenceNormal) < normal_threshold) { row.result = -5; continue; }
      row.result = 1;
      row.error = dot3(referenceNormal, diff);
      row.J[0] = referenceNormal.x; row.J[1] = referenceNormal.y; row.J[2] = referenceNormal.z;
      const V3 c = cross3(projectedVertex, referenceNormal);
      row.J[3] = c.x; row.J[4] = c.y; row.J[5] = c.z;
    }
}
// tracking.cpp:66-224 : 8 interleaved row blocks, 32 sums each, then the 8 rows added into row 0
static inline void reduce_kernel(float* out /*8*32*/, const TrackData* J, int JW, int W, int H) {
#pragma omp parallel for
  for (int blockIndex = 0; blockIndex < 8; ++blockIndex) {
    float s[32];
    for (float& v : s) v = 0.f;
    for (int y = blockIndex; y < H; y += 8)
      for (int x = 0; x < W; ++x) {
        const TrackData& row = J[x + y * JW];
        if (row.result < 1) {
          s[29] += row.result == -4 ? 1 : 0;
          s[30] += row.result == -5 ? 1 : 0;
          s[31] += row.result > -4 ? 1 : 0;
          continue;
        }
        s[0] += row.error * row.error;
        for (int i = 0; i < 6; ++i) s[1 + i] += row.error * row.J[i];
        int k = 7;
        for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) s[k++] += row.J[i] * row.J[j];
        s[28] += 1;
      }
    for (int i = 0; i < 32; ++i) out[blockIndex * 32 + i] = s[i];
  }
  for (int j = 1; j < 8; ++j) for (int i = 0; i < 32; ++i) out[i] += out[j * 32 + i];
}
// 6x6 Cholesky solve (tracking.cpp:57-64 uses Eigen::LLT); false when not positive definite
static inline bool solve6(const float* vals /*27: b[6], upper JTJ[21]*/, float x[6]) {
  float C[6][6], L[6][6] = {};
  int k = 6;
  for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { C[i][j] = vals[k]; C[j][i] = vals[k]; ++k; }
  for (int j = 0; j < 6; ++j) {
    float d = C[j][j];
    for (int t = 0; t < j; ++t) d -= L[j][t] * L[j][t];
    if (!(d > 0.f)) return false;
    L[j][j] = std::sqrt(d);
    for (int i = j + 1; i < 6; ++i) {
      float v = C[i][j];
      for (int t = 0; t < j; ++t) v -= L[i][t] * L[j][t];
      L[i][j] = v / L[j][j];
    }
  }
  float y[6];
  for (int i = 0; i < 6; ++i) { float v = vals[i]; for (int t = 0; t < i; ++t) v -= L[i][t] * y[t]; y[i] = v / L[i][i]; }
  for (int i = 5; i >= 0; --i) { float v = y[i]; for (int t = i + 1; t < 6; ++t) v -= L[t][i] * x[t]; x[i] = v / L[i][i]; }
  return true;
}
// Sophus::SE3f::exp(x), x = (upsilon, omega): closed form (Rodrigues + V matrix)
static inline M4 se3_exp(const float x[6]) {
  const float wx = x[3], wy = x[4], wz = x[5];
  const float theta2 = wx * wx + wy * wy + wz * wz, theta = std::sqrt(theta2);
  float A, B, Cc;        // sin(t)/t, (1-cos t)/t^2, (t - sin t)/t^3
  if (theta < 1e-4f) { A = 1.f - theta2 / 6.f; B = 0.5f - theta2 / 24.f; Cc = 1.f / 6.f - theta2 / 120.f; }
  else { A = std::sin(theta) / theta; B = (1.f - std::cos(theta)) / theta2; Cc = (theta - std::sin(theta)) / (theta2 * theta); }
  const float W[3][3] = {{0, -wz, wy}, {wz, 0, -wx}, {-wy, wx, 0}};
  float W2[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { W2[i][j] = 0; for (int t = 0; t < 3; ++t) W2[i][j] += W[i][t] * W[t][j]; }
  M4 T{};
  for (int i = 0; i < 3; ++i) {
    float tv = 0;
    for (int j = 0; j < 3; ++j) {
      const float I = i == j ? 1.f : 0.f;
      T.at(i, j) = I + A * W[i][j] + B * W2[i][j];
      tv += (I + B * W[i][j] + Cc * W2[i][j]) * x[j];
    }
    T.at(i, 3) = tv;
  }
  T.at(3, 3) = 1.f;
  return T;
}

// ----------------------------------------------------------------------------
// the pipeline state the hot path touches (DenseSLAMSystem.h:61-93)
// ----------------------------------------------------------------------------
// ----------------------------------------------------------------------------
// N4: se::algorithms::marching_cube (algorithms/meshing.hpp:158-208) with the functors DenseSLAMSystem::dump_mesh passes
// (DenseSLAMSystem.cpp:302-322: inside = val.x < 0, select = val.x).
// `table` = 256 rows of 16 edge indices, -1 terminated, in the reference's edge numbering (the role of triTable,
// algorithms/edge_tables.h; the tests pass a table generated by tests/mc_table_ref.py -- see DESIGN.md "N4" for how it
// relates to the reference's).  The reference visits the block list under `omp parallel for` and appends under a mutex, so
// its triangle order is schedule dependent; here blocks are visited serially in ascending key order.
// Output: 9 floats per triangle (vertexes[0..2], commons.h:166-168).
// ----------------------------------------------------------------------------
namespace meshing {
// compute_intersection, meshing.hpp:46-55: s + (0.0 - v1) * (d - s) / (v2 - v1); the double (0.0 - v1) rounds back to -v1
template <class F>
inline V3 compute_intersection(const Octree<F>& vol, const V3i& source, const V3i& dest) {
  const float voxelSize = vol.dim / vol.size;
  const V3 s{source.x * voxelSize, source.y * voxelSize, source.z * voxelSize};
  const V3 d{dest.x * voxelSize, dest.y * voxelSize, dest.z * voxelSize};
  const float v1 = vol.get_fine(source.x, source.y, source.z).x;
  const float v2 = vol.get_fine(dest.x, dest.y, dest.z).x;
  const float a = (float)(0.0 - v1), den = v2 - v1;
  return V3{s.x + (a * (d.x - s.x)) / den, s.y + (a * (d.y - s.y)) / den, s.z + (a * (d.z - s.z)) / den};
}
// interp_vertexes, meshing.hpp:57-91: edge -> (source, dest) corner offsets
template <class F>
inline V3 interp_vertexes(const Octree<F>& vol, int x, int y, int z, int edge) {
  static const int ends[12][6] = {
      {0, 0, 0, 1, 0, 0}, {1, 0, 0, 1, 0, 1}, {1, 0, 1, 0, 0, 1}, {0, 0, 0, 0, 0, 1},
      {0, 1, 0, 1, 1, 0}, {1, 1, 0, 1, 1, 1}, {1, 1, 1, 0, 1, 1}, {0, 1, 0, 0, 1, 1},
      {0, 0, 0, 0, 1, 0}, {1, 0, 0, 1, 1, 0}, {1, 0, 1, 1, 1, 1}, {0, 0, 1, 0, 1, 1}};
  if (edge < 0 || edge > 11) return V3{0, 0, 0};
  const int* e = ends[edge];
  return compute_intersection(vol, V3i{x + e[0], y + e[1], z + e[2]}, V3i{x + e[3], y + e[4], z + e[5]});
}
// compute_index, meshing.hpp:120-149 (both gather_points variants read the same voxels: the cached block for interior
// cells, get_fine for cells on a +face; get_fine alone covers both)
template <class F>
inline uint8_t compute_index(const Octree<F>& vol, int x, int y, int z) {
  static const int off[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 0, 1}, {0, 0, 1}, {0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {0, 1, 1}};   // :93-103
  F p[8];
  for (int i = 0; i < 8; ++i) p[i] = vol.get_fine(x + off[i][0], y + off[i][1], z + off[i][2]);
  for (int i = 0; i < 8; ++i) if (p[i].y == 0.f) return 0;
  uint8_t index = 0;
  for (int i = 0; i < 8; ++i) if (p[i].x < 0.f) index |= (uint8_t)(1u << i);
  return index;
}
inline bool check_vertex(const V3& v, float dim) { return v.x <= 0 || v.y <= 0 || v.z <= 0 || v.x > dim || v.y > dim || v.z > dim; }   // :151-153
}  // namespace meshing

template <class F>
inline void marching_cube(const Octree<F>& vol, const int8_t* table, std::vector<float>& triangles) {
  std::vector<int> order(vol.blocks.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return vol.blocks[a].code < vol.blocks[b].code; });
  const int size = vol.size;
  const float dim = vol.dim;
  for (int b : order) {
    const V3i start = vol.blocks[b].coords;
    const int tx = std::min(start.x + kBlockSide, size - 1), ty = std::min(start.y + kBlockSide, size - 1), tz = std::min(start.z + kBlockSide, size - 1);
    for (int x = start.x; x < tx; ++x)
      for (int y = start.y; y < ty; ++y)
        for (int z = start.z; z < tz; ++z) {
          const uint8_t index = meshing::compute_index(vol, x, y, z);
          const int8_t* edges = table + 16 * index;
          for (unsigned e = 0; edges[e] != -1 && e < 16; e += 3) {
            const V3 v1 = meshing::interp_vertexes(vol, x, y, z, edges[e]);
            const V3 v2 = meshing::interp_vertexes(vol, x, y, z, edges[e + 1]);
            const V3 v3 = meshing::interp_vertexes(vol, x, y, z, edges[e + 2]);
            if (meshing::check_vertex(v1, dim) || meshing::check_vertex(v2, dim) || meshing::check_vertex(v3, dim)) continue;
            const float t[9] = {v1.x, v1.y, v1.z, v2.x, v2.y, v2.z, v3.x, v3.y, v3.z};
            triangles.insert(triangles.end(), t, t + 9);
          }
        }
  }
}

template <class F> struct Pipeline {
  int W, H;
  Octree<F> map;
  std::vector<float> depth;          // float_depth_
  std::vector<V3> vertex, normal;    // vertex_, normal_
  std::vector<uint64_t> alloc_list;
  Counters ctr;
  bool count = false;                // enable sample counters (slows raycast a little)

  Pipeline(int size, float dim, int w, int h) : W(w), H(h) {
    map.init(size, dim);
    depth.assign((size_t)w * h, 0.f);
    vertex.assign((size_t)w * h, V3{0, 0, 0});
    normal.assign((size_t)w * h, V3{0, 0, 0});
  }

  // a1: preprocessing.cpp:161-188 ; returns false where the reference exit(1)s
  bool mm2meters(const uint16_t* in, int inW, int inH) {
    if (inW < W || inH < H) return false;
    if (inW % W != 0 || inH % H != 0) return false;
    if (inW / W != inH / H) return false;
    const int ratio = inW / W;
#pragma omp parallel for
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        depth[x + W * y] = in[x * ratio + inW * y * ratio] / 1000.0f;
    return true;
  }

  // a3: kfusion/alloc_impl.hpp:53-118
  unsigned build_allocation_list_sdf(const M4& pose, const M4& K_unused, const float k[4], float voxelSize, float band) {
    (void)K_unused;
    const int size = map.size;
    const size_t reserved = (size_t)(int)(map.dim / ((float)kBlockSide * voxelSize)) * (size_t)W * (size_t)H; // DenseSLAMSystem.cpp:212-215
    if (alloc_list.size() < reserved) alloc_list.resize(reserved);
    const float inverseVoxelSize = 1 / voxelSize;
    const int block_scale = map.leaves_level;
    const M4 invK = inverse_camera_matrix(k);
    const M4 kPose = mul44(pose, invK);
    const V3 camera = { pose.at(0,3), pose.at(1,3), pose.at(2,3) };
    const int numSteps = (int)std::ceil(band * inverseVoxelSize);
    std::atomic<unsigned> voxelCount{0};
    uint64_t* out = alloc_list.data();
#pragma omp parallel for
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) {
        if (depth[x + y * W] == 0) continue;
        const float d = depth[x + y * W];
        const V3 worldVertex = xform3(kPose, { (x + 0.5f) * d, (y + 0.5f) * d, d });
        const V3 direction = normalized3(camera - worldVertex);
        const V3 origin = worldVertex - (band * 0.5f) * direction;
        const V3 step = (direction * band) / (float)numSteps;
        V3 voxelPos = origin;
        for (int i = 0; i < numSteps; ++i) {
          const V3 vs = { std::floor(voxelPos.x * inverseVoxelSize), std::floor(voxelPos.y * inverseVoxelSize), std::floor(voxelPos.z * inverseVoxelSize) };
          if (vs.x < size && vs.y < size && vs.z < size && vs.x >= 0 && vs.y >= 0 && vs.z >= 0) {
            const int vx = (int)vs.x, vy = (int)vs.y, vz = (int)vs.z;
            const int b = map.fetch(vx, vy, vz);
            if (b < 0) {
              const uint64_t key = map.hash(vx, vy, vz, block_scale);
              const unsigned idx = voxelCount++;
              if (idx < reserved) out[idx] = key; else break;
            } else {
              map.blocks[b].active = true;
            }
          }
          voxelPos = voxelPos + step;
        }
      }
    }
    const unsigned written = voxelCount;
    return written >= reserved ? (unsigned)reserved : written;
  }

  // a4: bfusion/alloc_impl.hpp:37-129
  static float compute_stepsize(float dist_travelled, float hf_band, float voxelSize) {
    float half = hf_band * 0.5f;
    if (dist_travelled < hf_band) return voxelSize;
    else if (dist_travelled < hf_band + half) return 10.f * voxelSize;
    return 30.f * voxelSize;
  }
  static int step_to_depth(float step, int max_depth, float voxelsize) {
    return (int)(std::floor(std::log2(voxelsize / step)) + max_depth);
  }
  size_t build_octant_list_ofusion(const M4& pose, const float k[4], float voxelSize, float band) {
    const int size = map.size;
    const size_t reserved = (size_t)(int)(map.dim / ((float)kBlockSide * voxelSize)) * (size_t)W * (size_t)H;
    if (alloc_list.size() < reserved) alloc_list.resize(reserved);
    const float inverseVoxelSize = 1.f / voxelSize;
    const M4 invK = inverse_camera_matrix(k);
    const M4 kPose = mul44(pose, invK);
    const int max_depth = map.max_level;
    const int leaves_depth = map.leaves_level;
    const V3 camera = { pose.at(0,3), pose.at(1,3), pose.at(2,3) };
    std::atomic<unsigned> voxelCount{0};
    uint64_t* out = alloc_list.data();
#pragma omp parallel for
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) {
        if (depth[x + y * W] == 0) continue;
        int tree_depth = max_depth;
        float stepsize = voxelSize;
        const float d = depth[x + y * W];
        const V3 worldVertex = xform3(kPose, { (x + 0.5f) * d, (y + 0.5f) * d, d });
        const V3 direction = normalized3(camera - worldVertex);
        const V3 origin = worldVertex - (band * 0.5f) * direction;
        const float dist = norm3(camera - origin);
        V3 step = direction * stepsize;
        V3 voxelPos = origin;
        float travelled = 0.f;
        for (; travelled < dist; travelled += stepsize) {
          const V3 vs = { std::floor(voxelPos.x * inverseVoxelSize), std::floor(voxelPos.y * inverseVoxelSize), std::floor(voxelPos.z * inverseVoxelSize) };
          if (vs.x < size && vs.y < size && vs.z < size && vs.x >= 0 && vs.y >= 0 && vs.z >= 0) {
            const int vx = (int)vs.x, vy = (int)vs.y, vz = (int)vs.z;
            bool is_block = false;
            const int n = map.fetch_octant(vx, vy, vz, tree_depth, &is_block);
            if (n < 0) {
              const uint64_t key = map.hash(vx, vy, vz, std::min(tree_depth, leaves_depth));
              const unsigned idx = voxelCount++;
              if (idx < reserved) out[idx] = key;
            } else if (tree_depth >= leaves_depth) {
              map.blocks[n].active = true;
            }
          }
          stepsize = compute_stepsize(travelled, band, voxelSize);
          tree_depth = step_to_depth(stepsize, max_depth, voxelSize);
          step = direction * stepsize;
          voxelPos = voxelPos + step;
        }
      }
    }
    const size_t written = voxelCount;
    return written >= reserved ? reserved : written;
  }

  // a8: algorithms/filter.hpp:37-49
  bool in_frustum(const Block<F>& b, float voxelSize, const M4& cam) const {
    const V3 p = { (float)b.coords.x * voxelSize, (float)b.coords.y * voxelSize, (float)b.coords.z * voxelSize };
    const V3 v = xform3(cam, p);
    const float qx = v.x / v.z, qy = v.y / v.z;
    // (int) of a float outside int range (or nan) is undefined in the reference; we call it outside
    if (!(qx > -2147483648.f && qx < 2147483648.f && qy > -2147483648.f && qy < 2147483648.f)) return false;
    const int px = (int)qx, py = (int)qy;
    return px >= 0 && px < W && py >= 0 && py < H;
  }

  // a10: kfusion/mapping_impl.hpp:35-65
  static inline void field_update(SDF& data, const float* depth, int W, V3 pos, float pixx, float pixy, float mu, float /*timestamp*/, float /*voxelsize*/) {
    const int px = (int)pixx, py = (int)pixy;
    const float depthSample = depth[px + W * py];
    if (depthSample <= 0) return;
    const float a = pos.x / pos.z, b = pos.y / pos.z;
    const float diff = (depthSample - pos.z) * std::sqrt((1 + a * a) + b * b);
    if (diff > -mu) {
      const float sdf = std::fmin(1.f, diff / mu);
      data.x = std::max(-1.f, std::min((data.y * data.x + sdf) / (data.y + 1.f), 1.f));
      data.y = std::fmin(data.y + 1, (float)kMaxWeight);
    }
  }
  // a11: bfusion/mapping_impl.hpp:157-191
  static inline void field_update(OFusion& data, const float* depth, int W, V3 pos, float pixx, float pixy, float noiseFactor, float timestamp, float voxelsize) {
    const int px = (int)pixx, py = (int)pixy;
    const float depthSample = depth[px + W * py];
    if (depthSample <= 0) return;
    const float a = pos.x / pos.z, b = pos.y / pos.z;
    const float diff = (pos.z - depthSample) * std::sqrt((1 + a * a) + b * b);
    const float sigma = std::max(2 * voxelsize, std::min(noiseFactor * (pos.z * pos.z), 0.05f));
    float sample = h_new(diff / sigma);
    if (sample == 0.5f) return;
    sample = std::max(0.03f, std::min(sample, 0.97f));
    const double delta_t = (double)timestamp - data.y;
    // applyWindow (:150-155): delta_t narrows to float at the call
    float fraction = 1.f / (1.f + ((float)delta_t / 4.f));
    fraction = std::max(0.5f, fraction);
    data.x = data.x * fraction;
    // updateLogs (:145-148): unqualified log2 on a float argument.  perfstats.h includes <math.h>, whose libstdc++ wrapper
    // brings std::log2's float overload into the global namespace, so this is log2f and a float sum (confirmed by the
    // reference build in oracle/_ref, which the oracle matches bit for bit with this reading and not with the double one)
    const float upd = data.x + std::log2(sample / (1.f - sample));
    data.x = std::max(-1000.f, std::min(upd, 1000.f));
    data.y = (double)timestamp;
  }

  // a9: functors/projective_functor.hpp:73-111
  void update_block(Block<F>& block, float voxel_size, const M4& Tcw, const M4& K, float mu, float timestamp) {
    const V3 delta = rot3(Tcw, { voxel_size, 0.f, 0.f });
    const V3 cameraDelta = rot3(K, delta);
    bool is_visible = false;
    const int bx = block.coords.x, by = block.coords.y, bz = block.coords.z;
    for (int z = bz; z < bz + 8; ++z)
      for (int y = by; y < by + 8; ++y) {
        const V3 start = xform3(Tcw, { (float)bx * voxel_size, (float)y * voxel_size, (float)z * voxel_size });
        const V3 camerastart = rot3(K, start);
        for (int x = 0; x < 8; ++x) {
          const V3 cv = camerastart + ((float)x * cameraDelta);
          const V3 pos = start + ((float)x * delta);
          if (pos.z < 0.0001f) continue;
          const float inverse_depth = 1.f / cv.z;
          const float pixx = cv.x * inverse_depth + 0.5f, pixy = cv.y * inverse_depth + 0.5f;
          if (pixx < 0.5f || pixx > (float)W - 1.5f || pixy < 0.5f || pixy > (float)H - 1.5f) continue;
          is_visible = true;
          field_update(block.data[x + (y - by) * 8 + (z - bz) * 64], depth.data(), W, pos, pixx, pixy, mu, timestamp, voxel_size);
        }
      }
    block.active = is_visible;
  }
  // a12: functors/projective_functor.hpp:113-137
  void update_node(Node<F>& node, float voxel_size, const M4& Tcw, const M4& K, float mu, float timestamp) {
    const V3i voxel = morton_decode(node.code);     // unpack_morton(code_) -- level bits included, as in the reference
    const float hs = 0.5f * voxel_size * (float)node.side;
    const V3 delta = rot3(Tcw, { hs, hs, hs });
    const V3 delta_c = rot3(K, delta);
    const V3 base_cam = xform3(Tcw, { voxel_size * (float)voxel.x, voxel_size * (float)voxel.y, voxel_size * (float)voxel.z });
    const V3 basepix_hom = rot3(K, base_cam);
    for (int i = 0; i < 8; ++i) {
      const V3 dirf = { (float)((i & 1) > 0), (float)((i & 2) > 0), (float)((i & 4) > 0) };
      const V3 vox_cam = { base_cam.x + dirf.x * delta.x, base_cam.y + dirf.y * delta.y, base_cam.z + dirf.z * delta.z };
      const V3 pix_hom = { basepix_hom.x + dirf.x * delta_c.x, basepix_hom.y + dirf.y * delta_c.y, basepix_hom.z + dirf.z * delta_c.z };
      if (vox_cam.z < 0.0001f) continue;
      const float inverse_depth = 1.f / pix_hom.z;
      const float pixx = pix_hom.x * inverse_depth + 0.5f, pixy = pix_hom.y * inverse_depth + 0.5f;
      if (pixx < 0.5f || pixx > (float)W - 1.5f || pixy < 0.5f || pixy > (float)H - 1.5f) continue;
      field_update(node.value[i], depth.data(), W, vox_cam, pixx, pixy, mu, timestamp, voxel_size);
    }
  }

  // DenseSLAMSystem::integration body (DenseSLAMSystem.cpp:211-253) without the frame gate.
  // Returns the number of raw allocation requests.
  unsigned integrate(const M4& pose, const float k[4], float mu, unsigned frame) {
    const float voxelsize = map.dim / (float)map.size;
    const M4 K = camera_matrix(k);
    unsigned allocated;
    if (traits<F>::is_sdf) allocated = build_allocation_list_sdf(pose, K, k, voxelsize, 2 * mu);
    else allocated = (unsigned)build_octant_list_ofusion(pose, k, voxelsize, 6 * mu);
    ctr.n_keys_raw += allocated;
    map.allocate(alloc_list.data(), (int)allocated, &ctr);

    const M4 Tcw = rigid_inverse(pose);
    const float timestamp = (1.f / 30.f) * (float)frame;
    // build_active_list: projective_functor.hpp:54-71 + filter.hpp:61-118
    const M4 cam = mul44(K, Tcw);
    const int nb = (int)map.blocks.size();
    std::vector<int> active;
    {
      std::vector<uint8_t> flag((size_t)nb);
#pragma omp parallel for
      for (int i = 0; i < nb; ++i) flag[i] = map.blocks[i].active || in_frustum(map.blocks[i], voxelsize, cam);
      for (int i = 0; i < nb; ++i) if (flag[i]) active.push_back(i);
    }
    ctr.n_active += active.size();
    const int na = (int)active.size();
#pragma omp parallel for
    for (int i = 0; i < na; ++i) update_block(map.blocks[active[i]], voxelsize, Tcw, K, mu, timestamp);
    const int nn = (int)map.nodes.size();
    ctr.n_nodes += (uint64_t)nn;
#pragma omp parallel for
    for (int i = 0; i < nn; ++i) update_node(map.nodes[i], voxelsize, Tcw, K, mu, timestamp);
    return allocated;
  }

  // ---- N1: DenseSLAMSystem::preprocessing (filter) + ::tracking (DenseSLAMSystem.cpp:128-189) ----
  std::vector<std::vector<float>> scaled_depth;
  std::vector<std::vector<V3>> input_vertex, input_normal;
  std::vector<TrackData> tracking_result;
  float reduction[8 * 32];
  void ensure_pyramid(int levels) {
    if ((int)scaled_depth.size() == levels) return;
    scaled_depth.assign(levels, {}); input_vertex.assign(levels, {}); input_normal.assign(levels, {});
    for (int i = 0; i < levels; ++i) {
      const size_t n = (size_t)(W >> i) * (H >> i);
      scaled_depth[i].assign(n, 0.f); input_vertex[i].assign(n, V3{0, 0, 0}); input_normal[i].assign(n, V3{0, 0, 0});
    }
    tracking_result.assign((size_t)W * H, TrackData{0, 0.f, {0, 0, 0, 0, 0, 0}});
  }
  // second half of preprocessing(): bilateral filter or plain copy into scaled_depth_[0]
  void filter_depth(bool filter, int levels) {
    ensure_pyramid(levels);
    if (filter) bilateral_filter(scaled_depth[0], depth, W, H);
    else scaled_depth[0] = depth;
  }
  // tracking(): pose is updated in place; returns checkPoseKernel's verdict.  raycast_pose = pose of the
  // vertex/normal maps (raycast_pose_).  iterations[level] as Configuration::pyramid.
  bool tracking(M4& pose, const M4& raycast_pose, const float k[4], float icp_threshold, const int* iterations, int levels) {
    ensure_pyramid(levels);
    for (int i = 1; i < levels; ++i) half_sample_robust(scaled_depth[i], scaled_depth[i - 1], W >> i, H >> i, W >> (i - 1), kEDelta * 3, 1);
    for (int i = 0; i < levels; ++i) {
      const float ks[4] = { k[0] / (float)(1 << i), k[1] / (float)(1 << i), k[2] / (float)(1 << i), k[3] / (float)(1 << i) };
      depth2vertex(input_vertex[i], scaled_depth[i], W >> i, H >> i, inverse_camera_matrix(ks));
      vertex2normal(input_normal[i], input_vertex[i], W >> i, H >> i, k[1] < 0);
    }
    const M4 old_pose = pose;
    const M4 projectReference = mul44(camera_matrix(k), rigid_inverse(raycast_pose));
    for (int level = levels - 1; level >= 0; --level) {
      const int lw = W / (1 << level), lh = H / (1 << level);
      for (int i = 0; i < iterations[level]; ++i) {
        track_kernel(tracking_result.data(), input_vertex[level], input_normal[level], lw, lh, vertex, normal, W, H, pose, projectReference,
                     kDistThreshold, kNormalThreshold);
        reduce_kernel(reduction, tracking_result.data(), W, lw, lh);
        float x[6];
        if (!solve6(reduction + 1, x)) for (float& v : x) v = 0.f;
        pose = mul44(se3_exp(x), pose);
        float n2 = 0; for (float v : x) n2 += v * v;
        if (std::sqrt(n2) < icp_threshold) break;
      }
    }
    // checkPoseKernel (tracking.cpp:320-336)
    if ((std::sqrt(reduction[0] / reduction[28]) > 2e-2) || (reduction[28] / (float)(W * H) < kTrackThreshold)) { pose = old_pose; return false; }
    return true;
  }

  // a13: rendering.cpp:50-90 ; view = raycast_pose * K^-1
  void raycast(const M4& pose, const float k[4], float mu) {
    const M4 view = mul44(pose, inverse_camera_matrix(k));
    const float step = map.dim / (float)map.size;        // DenseSLAMSystem.cpp:197
    const float largestep = step * (float)kBlockSide;
    uint64_t ng = 0, ni = 0, ngr = 0;
#pragma omp parallel for reduction(+:ng,ni,ngr)
    for (int y = 0; y < H; ++y) {
      Counters local; Counters* c = count ? &local : nullptr;
      for (int x = 0; x < W; ++x) {
        const V3 dir = normalized3(rot3(view, { (float)x, (float)y, 1.f }));
        const V3 transl = { view.at(0,3), view.at(1,3), view.at(2,3) };
        RayIterator<F> ray(map, transl, dir, kNearPlane, kFarPlane);
        ray.next();
        const float t_min = ray.tcmin();
        const V4 hit = t_min > 0.f ? raycast_field(map, transl, dir, t_min, ray.tmax(), mu, step, largestep, c) : V4{0, 0, 0, 0};
        if (hit.w > 0.0) {
          vertex[x + y * W] = { hit.x, hit.y, hit.z };
          const V3 surfNorm = vol_grad(map, { hit.x, hit.y, hit.z }, c);
          if (norm3(surfNorm) == 0) normal[x + y * W] = { kInvalid, 0, 0 };
          else normal[x + y * W] = traits<F>::is_sdf ? normalized3(-1.f * surfNorm) : normalized3(surfNorm);
        } else {
          vertex[x + y * W] = { 0, 0, 0 };
          normal[x + y * W] = { kInvalid, 0, 0 };
        }
      }
      ng += local.n_get; ni += local.n_interp; ngr += local.n_grad;
    }
    ctr.n_get += ng; ctr.n_interp += ni; ctr.n_grad += ngr;
  }

  // a18: rendering.cpp:214-283 ; `render` == view pose differs from the raycast pose
  void render_volume(uint8_t* out, const M4& viewPose, const float k[4], float mu, float largestep, bool render) {
    const M4 view = mul44(viewPose, inverse_camera_matrix(k));
    const float step = map.dim / (float)map.size;        // DenseSLAMSystem.cpp:282
    const V3 light = { viewPose.at(0,3), viewPose.at(1,3), viewPose.at(2,3) };
    const float farP = kFarPlane * 2.0f;
#pragma omp parallel for
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        V3 test{0, 0, 0}, surfNorm;
        const int o = (x + W * y) * 4;
        if (render) {
          const V3 dir = normalized3(rot3(view, { (float)x, (float)y, 1.f }));
          const V3 transl = { view.at(0,3), view.at(1,3), view.at(2,3) };
          RayIterator<F> ray(map, transl, dir, kNearPlane, farP);
          ray.next();
          const float t_min = ray.tmin();
          const V4 hit = t_min > 0.f ? raycast_field(map, transl, dir, t_min, ray.tmax(), mu, step, largestep, nullptr) : V4{0, 0, 0, 0};
          if (hit.w > 0) {
            test = { hit.x, hit.y, hit.z };
            surfNorm = vol_grad(map, test, nullptr);
            if (traits<F>::is_sdf) surfNorm = -1.f * surfNorm;
          } else surfNorm = { kInvalid, 0, 0 };
        } else {
          test = vertex[x + W * y];
          surfNorm = normal[x + W * y];
        }
        if (surfNorm.x != kInvalid && norm3(surfNorm) > 0) {
          const V3 diff = normalized3(test - light);
          const float dirv = std::fmax(dot3(normalized3(surfNorm), diff), 0.f);
          float col = dirv + kAmbient;
          col = std::min(std::max(col, 0.f), 1.f);
          col *= 255.f;
          out[o + 0] = (uint8_t)col; out[o + 1] = (uint8_t)col; out[o + 2] = (uint8_t)col; out[o + 3] = 0;
        } else {
          out[o + 0] = 0; out[o + 1] = 0; out[o + 2] = 0; out[o + 3] = 0;
        }
      }
  }

  // rendering.cpp:111-152 + commons.h:105-164
  void render_depth(uint8_t* out) const {
    const float rangeScale = 1 / (kFarPlane - kNearPlane);
#pragma omp parallel for
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const int pos = y * W + x, o = pos * 4;
        const float d = depth[pos];
        if (d < kNearPlane) { out[o] = 255; out[o+1] = 255; out[o+2] = 255; out[o+3] = 0; }
        else if (d > kFarPlane) { out[o] = 0; out[o+1] = 0; out[o+2] = 0; out[o+3] = 0; }
        else {
          double h = (double)((d - kNearPlane) * rangeScale);
          const double v = 0.75, m = 0.25, sv = 0.6667;
          h *= 6.0;
          const int sextant = (int)h;
          const double fract = h - sextant, vsf = v * sv * fract, mid1 = m + vsf, mid2 = v - vsf;
          double r = 0, g = 0, b = 0;
          switch (sextant) {
            case 0: r = v; g = mid1; b = m; break;
            case 1: r = mid2; g = v; b = m; break;
            case 2: r = m; g = v; b = mid1; break;
            case 3: r = m; g = mid2; b = v; break;
            case 4: r = mid1; g = m; b = v; break;
            case 5: r = v; g = m; b = mid2; break;
            default: break;
          }
          out[o] = (uint8_t)(r * 255); out[o+1] = (uint8_t)(g * 255); out[o+2] = (uint8_t)(b * 255); out[o+3] = 0;
        }
      }
  }
};

// rendering.cpp:154-212 : colour per TrackData::result
static inline void render_track(uint8_t* out, const int* result, int stride_ints, int W, int H) {
  for (int i = 0; i < W * H; ++i) {
    uint8_t r, g, b;
    switch (result[(size_t)i * stride_ints]) {
      case 1: r = 128; g = 128; b = 128; break;
      case -1: r = 0; g = 0; b = 0; break;
      case -2: r = 255; g = 0; b = 0; break;
      case -3: r = 0; g = 255; b = 0; break;
      case -4: r = 0; g = 0; b = 255; break;
      case -5: r = 255; g = 255; b = 0; break;
      default: r = 255; g = 128; b = 128; break;
    }
    out[4*i] = r; out[4*i+1] = g; out[4*i+2] = b; out[4*i+3] = 0;
  }
}

inline BsplineLut::BsplineLut() {
  for (int i = 0; i < 1000; ++i) {
    const double t = -3.0 + 6.0 * (double)i / 999.0;
    v[i] = (float)bspline_closed(t);
  }
}

}  // namespace seo
