// N4: marching cubes over the VoxelBlock pool -- se::algorithms::marching_cube (se_core/include/se/algorithms/meshing.hpp:158-208)
// with the `inside` / `select` functors DenseSLAMSystem::dump_mesh passes (se_denseslam/src/DenseSLAMSystem.cpp:302-322:
// inside = val.x < 0, select = val.x).
//
// Case table.  Which triangles a cube configuration emits is the classic 256 x 16 marching-cubes tabulation (Lorensen & Cline 1987,
// P. Bourke's public-domain list), which the reference ships as edge_tables.h.  The order of the triangles inside a case and the
// diagonals that split its polygons are a convention that cannot be re-derived, so a mesh that is to equal the reference's
// triangle for triangle has to use the same list: se_mc_table.cuh holds it as data, in this library's own encoding (256 strings
// of hexadecimal edge indices; scripts/make_mc_table.py).  tests/mc_table_ref.py is a first-principles generator of the
// table's geometry -- on every cube face, walked counter-clockwise from outside, each run of inside corners is cut off by one
// directed segment, segments chain into polygons -- and tests/test_meshing.py requires the list's 256 cases to have exactly those
// directed polygon boundaries.  (Round 1 meshed with the generator's own fans: same cells, vertices and triangle count as the
// reference, other diagonals in 158 cases.)
#pragma once
#include "se_map.cuh"

#include <cstdint>

namespace se_b200 {

constexpr int kMcRow = 16;          // 5 triangles x 3 edges + terminator, the reference's row width

// corner c of the cell sits at voxel + kMcCorner[c]; edge e joins kMcEdge[e][0] (source) -> [1] (dest)   (meshing.hpp:58-104)
SE_HD void mc_corner(int c, int& dx, int& dy, int& dz) {
  dx = (c == 1) | (c == 2) | (c == 5) | (c == 6);
  dy = c >> 2;
  dz = (c == 2) | (c == 3) | (c == 6) | (c == 7);
}
SE_HD void mc_edge(int e, int& a, int& b) {
  if (e < 4) { a = e == 3 ? 0 : e; b = e == 3 ? 3 : e + 1; }
  else if (e < 8) { a = e == 7 ? 4 : e; b = e == 7 ? 7 : e + 1; }
  else { a = e - 8; b = e - 4; }
}

// host: the 256 x 16 case table (-1 terminated rows of edge indices, three per triangle)
inline void mc_case_table(int8_t table[256 * kMcRow]) {
  static const char* const kCases[256] = {
#include "se_mc_table.cuh"
  };
  for (int idx = 0; idx < 256; ++idx) {
    int8_t* row = table + idx * kMcRow;
    int n = 0;
    for (const char* c = kCases[idx]; *c && n < kMcRow - 1; ++c) row[n++] = (int8_t)(*c <= '9' ? *c - '0' : *c - 'a' + 10);
    while (n < kMcRow) row[n++] = -1;
  }
}

#ifdef __CUDACC__

constexpr int kMeshThreads = 512;     // one thread per cell of the block

__global__ void k_iota(int* __restrict__ v, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = i; }

__device__ __forceinline__ float2 mc_pack(const SdfVoxel& v) { return make_float2(v.x, v.y == 0.f ? 0.f : 1.f); }
__device__ __forceinline__ float2 mc_pack(const OfuVoxel& v) { return make_float2(v.x, v.y == 0.0 ? 0.f : 1.f); }

// compute_intersection (meshing.hpp:46-55): s + (0 - v1) * (d - s) / (v2 - v1), per component, fp32
__device__ __forceinline__ float mc_lerp(int s, int d, float voxel, float v1, float v2) {
  const float fs = (float)s * voxel, fd = (float)d * voxel;
  return fs + ((0.f - v1) * (fd - fs)) / (v2 - v1);
}

// One CTA per VoxelBlock, taken in ascending key order (`order`), one thread per cell in the reference's loop order
// (x outer, z inner, meshing.hpp:181-183).  The 9^3 corner samples the block's cells touch are staged in shared memory
// (the block itself plus one layer of up to 7 neighbours, Octree::get_fine: unallocated => initValue, weight 0 => no
// triangles, meshing.hpp:131-138).  WRITE = false counts the triangles the block emits, WRITE = true stores them at
// offsets[sorted position] + the cell's rank, which reproduces the order of a serial run over the sorted block list.
template <class V, bool WRITE>
__global__ void __launch_bounds__(kMeshThreads) k_mesh_blocks(MapView<V> map, const int* __restrict__ order, const int8_t* __restrict__ table,
                                                              const unsigned long long* __restrict__ offsets, unsigned int* __restrict__ counts,
                                                              float* __restrict__ out) {
  __shared__ float2 s_c[9 * 9 * 9];               // [x][y][z]: (field value, 1 if the voxel was ever updated)
  __shared__ int s_tab[256 * kMcRow / 4];
  __shared__ unsigned int s_warp[kMeshThreads / 32];
  const int tid = threadIdx.x;
  const int b = order[blockIdx.x];
  const int4 base = map.block_coord[b];
  for (int i = tid; i < 256 * kMcRow / 4; i += kMeshThreads) s_tab[i] = reinterpret_cast<const int*>(table)[i];
  for (int i = tid; i < 729; i += kMeshThreads) {
    const int cx = i / 81, cy = (i / 9) % 9, cz = i % 9;
    const int gx = base.x + cx, gy = base.y + cy, gz = base.z + cz;
    float2 val = mc_pack(FieldTraits<V>::init());
    if (gx < map.size && gy < map.size && gz < map.size) {
      const int id = ((cx | cy | cz) < 8) ? b : fetch_block(map, gx, gy, gz);
      if (id >= 0) val = mc_pack(load_voxel(map.block_data + (size_t)id * kBlockVoxels + (gx & 7) + ((gy & 7) << 3) + ((gz & 7) << 6)));
    }
    s_c[i] = val;
  }
  __syncthreads();
  const int8_t* tab = reinterpret_cast<const int8_t*>(s_tab);
  const int x = tid >> 6, y = (tid >> 3) & 7, z = tid & 7;
  const int gx = base.x + x, gy = base.y + y, gz = base.z + z;
  float tri[5][9];
  int kept = 0;
  // top = min(coordinates + side, size - 1) (meshing.hpp:178-180): the last voxel layer of the volume starts no cell
  if (gx < map.size - 1 && gy < map.size - 1 && gz < map.size - 1) {
    bool all_seen = true;
    int index = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      int dx, dy, dz; mc_corner(c, dx, dy, dz);
      const float2 p = s_c[((x + dx) * 9 + (y + dy)) * 9 + (z + dz)];
      all_seen &= p.y != 0.f;                      // compute_index returns 0 when any corner has y == 0 (meshing.hpp:131-138)
      index |= (p.x < 0.f) << c;                   // inside(points[c]) (meshing.hpp:140-147, DenseSLAMSystem.cpp:305-313)
    }
    if (all_seen) {
      const float voxel = map.dim / (float)map.size;
      const float dim = map.dim;
      const int8_t* edges = tab + index * kMcRow;
      for (int e = 0; e < kMcRow - 1 && edges[e] != -1; e += 3) {
        float* t = tri[kept];
        bool drop = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          int a, c2; mc_edge(edges[e + k], a, c2);
          int ax, ay, az, bx, by, bz; mc_corner(a, ax, ay, az); mc_corner(c2, bx, by, bz);
          const float v1 = s_c[((x + ax) * 9 + (y + ay)) * 9 + (z + az)].x, v2 = s_c[((x + bx) * 9 + (y + by)) * 9 + (z + bz)].x;   // get_fine(source), get_fine(dest)
          const float px = mc_lerp(gx + ax, gx + bx, voxel, v1, v2);
          const float py = mc_lerp(gy + ay, gy + by, voxel, v1, v2);
          const float pz = mc_lerp(gz + az, gz + bz, voxel, v1, v2);
          t[3 * k] = px; t[3 * k + 1] = py; t[3 * k + 2] = pz;
          drop |= (px <= 0.f) | (py <= 0.f) | (pz <= 0.f) | (px > dim) | (py > dim) | (pz > dim);   // checkVertex (meshing.hpp:151-153)
        }
        if (!drop) ++kept;
      }
    }
  }
  // rank of this cell's first triangle inside the block, cells in thread order
  const unsigned lane = tid & 31, warp = tid >> 5;
  unsigned incl = kept;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (unsigned)d) incl += o; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kMeshThreads / 32; ++w) { const unsigned c = s_warp[w]; if ((unsigned)w < warp) before += c; total += c; }
  if (!WRITE) { if (tid == 0) counts[blockIdx.x] = total; return; }
  float* dst = out + (offsets[blockIdx.x] + before + incl - kept) * 9ull;
  for (int j = 0; j < kept; ++j)
#pragma unroll
    for (int q = 0; q < 9; ++q) dst[j * 9 + q] = tri[j][q];
}

#endif  // __CUDACC__
}  // namespace se_b200
