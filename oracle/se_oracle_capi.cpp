// se_oracle_capi.cpp -- plain-C entry points over oracle/se_oracle.hpp so that the
// Python tests (ctypes) and bench.py's cpu_baseline leg can drive the CPU oracle.
// TEST INFRASTRUCTURE ONLY -- see the header of se_oracle.hpp.
#include "se_oracle.hpp"
#include <numeric>

using namespace seo;

namespace {
struct Handle {
  int field;   // 0 SDF, 1 OFusion
  Pipeline<SDF>* s = nullptr;
  Pipeline<OFusion>* o = nullptr;
};
M4 to_m4(const float* p) { M4 m; std::memcpy(m.m, p, sizeof(m.m)); return m; }
}  // namespace

#define DISPATCH(h, expr) do { Handle* H_ = (Handle*)(h); if (H_->field == 0) { auto& P = *H_->s; expr; } else { auto& P = *H_->o; expr; } } while (0)

extern "C" {

// ---- key codec / key ops --------------------------------------------------
uint64_t seo_morton_encode(int x, int y, int z) { return morton_encode((uint64_t)x, (uint64_t)y, (uint64_t)z); }
void seo_morton_decode(uint64_t c, int out[3]) { V3i v = morton_decode(c); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
uint64_t seo_level_mask(int i) { return level_mask(i); }
uint64_t seo_key_encode(int x, int y, int z, int level, int max_depth) { return key_encode(x, y, z, level, max_depth); }
int seo_key_descendant(uint64_t o, uint64_t a, int max_depth) { return key_descendant(o, a, max_depth); }
uint64_t seo_key_parent(uint64_t o, int max_depth) { return key_parent(o, max_depth); }
int seo_key_child_id(uint64_t o, int level, int max_depth) { return key_child_id(o, level, max_depth); }
void seo_key_far_corner(uint64_t o, int level, int max_depth, int out[3]) { V3i v = key_far_corner(o, level, max_depth); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
void seo_key_face_neighbour(uint64_t o, unsigned face, unsigned l, unsigned max_depth, int out[3]) { V3i v = key_face_neighbour(o, face, l, max_depth); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
void seo_key_exterior_neighbours(uint64_t out[7], uint64_t o, int level, int max_depth) { key_exterior_neighbours(out, o, level, max_depth); }
void seo_key_siblings(uint64_t out[8], uint64_t o, int max_depth) { key_siblings(out, o, max_depth); }
int seo_keys_unique(uint64_t* k, int n) { return keys_unique(k, n); }
int seo_keys_filter_ancestors(uint64_t* k, int n, int max_depth) { return keys_filter_ancestors(k, n, max_depth); }
int seo_keys_unique_multiscale(uint64_t* k, int n, unsigned level) { return keys_unique_multiscale(k, n, level); }
float seo_bspline_lut(int i) { return BsplineLut::get().v[i]; }

int seo_omp_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void seo_set_omp_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// ---- pipeline -------------------------------------------------------------
void* seo_create(int field, int size, float dim, int W, int H) {
  Handle* h = new Handle;
  h->field = field;
  if (field == 0) h->s = new Pipeline<SDF>(size, dim, W, H);
  else h->o = new Pipeline<OFusion>(size, dim, W, H);
  return h;
}
void seo_destroy(void* hh) { Handle* h = (Handle*)hh; delete h->s; delete h->o; delete h; }

int seo_preprocess(void* h, const uint16_t* in, int inW, int inH) { int r = 0; DISPATCH(h, r = P.mm2meters(in, inW, inH) ? 0 : 1); return r; }
void seo_set_depth(void* h, const float* d) { DISPATCH(h, std::memcpy(P.depth.data(), d, sizeof(float) * P.depth.size())); }
void seo_get_depth(void* h, float* d) { DISPATCH(h, std::memcpy(d, P.depth.data(), sizeof(float) * P.depth.size())); }
unsigned seo_integrate(void* h, const float* pose, const float* k, float mu, unsigned frame) { unsigned r = 0; DISPATCH(h, r = P.integrate(to_m4(pose), k, mu, frame)); return r; }
void seo_raycast(void* h, const float* pose, const float* k, float mu) { DISPATCH(h, P.raycast(to_m4(pose), k, mu)); }
void seo_render_volume(void* h, uint8_t* out, const float* viewpose, const float* k, float mu, float largestep, int render) { DISPATCH(h, P.render_volume(out, to_m4(viewpose), k, mu, largestep, render != 0)); }
void seo_render_depth(void* h, uint8_t* out) { DISPATCH(h, P.render_depth(out)); }
void seo_render_track(uint8_t* out, const int* result, int stride_ints, int W, int H) { render_track(out, result, stride_ints, W, H); }
void seo_get_vertex(void* h, float* out) { DISPATCH(h, std::memcpy(out, P.vertex.data(), sizeof(V3) * P.vertex.size())); }
void seo_get_normal(void* h, float* out) { DISPATCH(h, std::memcpy(out, P.normal.data(), sizeof(V3) * P.normal.size())); }
void seo_set_vertex_normal(void* h, const float* v, const float* n) { DISPATCH(h, { std::memcpy(P.vertex.data(), v, sizeof(V3) * P.vertex.size()); std::memcpy(P.normal.data(), n, sizeof(V3) * P.normal.size()); }); }

int seo_block_count(void* h) { int r = 0; DISPATCH(h, r = (int)P.map.blocks.size()); return r; }
int seo_node_count(void* h) { int r = 0; DISPATCH(h, r = (int)P.map.nodes.size()); return r; }
int seo_voxel_bytes(void* h) { return ((Handle*)h)->field == 0 ? (int)sizeof(SDF) : (int)sizeof(OFusion); }

// blocks sorted by key; data is n*512 voxels in the field's native struct layout
void seo_get_blocks_sorted(void* h, uint64_t* keys, int* coords, uint8_t* active, void* data) {
  DISPATCH(h, {
    const int n = (int)P.map.blocks.size();
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return P.map.blocks[a].code < P.map.blocks[b].code; });
    using VT = std::remove_reference_t<decltype(P.map.blocks[0].data[0])>;
    VT* d = (VT*)data;
    for (int i = 0; i < n; ++i) {
      const auto& b = P.map.blocks[order[i]];
      if (keys) keys[i] = b.code;
      if (coords) { coords[3*i] = b.coords.x; coords[3*i+1] = b.coords.y; coords[3*i+2] = b.coords.z; }
      if (active) active[i] = b.active ? 1 : 0;
      if (d) std::memcpy(d + (size_t)i * 512, b.data, sizeof(VT) * 512);
    }
  });
}
void seo_get_nodes_sorted(void* h, uint64_t* codes, uint32_t* side, uint8_t* mask, void* values) {
  DISPATCH(h, {
    const int n = (int)P.map.nodes.size();
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return P.map.nodes[a].code < P.map.nodes[b].code; });
    using VT = std::remove_reference_t<decltype(P.map.nodes[0].value[0])>;
    VT* d = (VT*)values;
    for (int i = 0; i < n; ++i) {
      const auto& nd = P.map.nodes[order[i]];
      if (codes) codes[i] = nd.code;
      if (side) side[i] = nd.side;
      if (mask) mask[i] = nd.children_mask;
      if (d) std::memcpy(d + (size_t)i * 8, nd.value, sizeof(VT) * 8);
    }
  });
}

int seo_allocate(void* h, const uint64_t* keys, int n) {
  std::vector<uint64_t> k(keys, keys + n);
  int r = 0;
  DISPATCH(h, r = P.map.allocate(k.data(), n, &P.ctr) ? 1 : 0);
  return r;
}
int seo_fetch(void* h, int x, int y, int z) { int r = 0; DISPATCH(h, r = P.map.fetch(x, y, z) >= 0); return r; }
int seo_fetch_octant(void* h, int x, int y, int z, int depth) { int r = 0; DISPATCH(h, r = P.map.fetch_octant(x, y, z, depth) >= 0); return r; }
// code of the octant found by fetch_octant, or ~0 when absent
uint64_t seo_fetch_octant_code(void* h, int x, int y, int z, int depth) {
  uint64_t r = ~0ull;
  DISPATCH(h, { bool isb = false; int n = P.map.fetch_octant(x, y, z, depth, &isb); if (n >= 0) r = isb ? P.map.blocks[n].code : P.map.nodes[n].code; });
  return r;
}
void seo_get_fine(void* h, int x, int y, int z, double out[2]) { DISPATCH(h, { auto v = P.map.get_fine(x, y, z); out[0] = v.x; out[1] = (double)v.y; }); }
void seo_get_coarse(void* h, int x, int y, int z, double out[2]) { DISPATCH(h, { auto v = P.map.get(x, y, z); out[0] = v.x; out[1] = (double)v.y; }); }
void seo_set_voxel(void* h, int x, int y, int z, double vx, double vy) {
  DISPATCH(h, { auto v = P.map.get_fine(x, y, z); v.x = (float)vx; v.y = (decltype(v.y))vy; P.map.set(x, y, z, v); });
}
// set value_[slot] of the node found by fetch_octant(x,y,z,depth); returns 0 when absent or a block
int seo_set_node_value(void* h, int x, int y, int z, int depth, int slot, double vx, double vy) {
  int r = 0;
  DISPATCH(h, { bool isb = false; int n = P.map.fetch_octant(x, y, z, depth, &isb); if (n >= 0 && !isb) { P.map.nodes[n].value[slot].x = (float)vx; P.map.nodes[n].value[slot].y = (decltype(P.map.nodes[n].value[slot].y))vy; r = 1; } });
  return r;
}
float seo_interp(void* h, float x, float y, float z) { float r = 0; DISPATCH(h, r = P.map.interp({x, y, z})); return r; }
void seo_grad(void* h, float x, float y, float z, float out[3]) { DISPATCH(h, { V3 g = P.map.grad({x, y, z}); out[0] = g.x; out[1] = g.y; out[2] = g.z; }); }
void seo_gather(void* h, int x, int y, int z, float out[8]) { DISPATCH(h, P.map.gather_points(x, y, z, out)); }

// codes of the blocks ray_iterator::next() returns, in order; also tmin/tmax/tcmin after the first next()
int seo_ray_blocks(void* h, const float* origin, const float* dir, float nearP, float farP, uint64_t* out, int max_out, float tinfo[3]) {
  int n = 0;
  DISPATCH(h, {
    using FT = std::remove_reference_t<decltype(P.map.blocks[0].data[0])>;
    RayIterator<FT> it(P.map, {origin[0], origin[1], origin[2]}, {dir[0], dir[1], dir[2]}, nearP, farP);
    int b;
    bool first = true;
    while ((b = it.next()) >= 0) {
      if (first && tinfo) { tinfo[0] = it.tmin(); tinfo[1] = it.tmax(); tinfo[2] = it.tcmin(); }
      first = false;
      if (n < max_out) out[n] = P.map.blocks[b].code;
      ++n;
    }
    if (first && tinfo) { tinfo[0] = it.tmin(); tinfo[1] = it.tmax(); tinfo[2] = it.tcmin(); }
  });
  return n;
}

// ---- N1 ----------------------------------------------------------------------
void seo_filter_depth(void* h, int filter, int levels) { DISPATCH(h, P.filter_depth(filter != 0, levels)); }
int seo_tracking(void* h, float* pose_io, const float* raycast_pose, const float* k, float icp_threshold, const int* iterations, int levels) {
  int r = 0;
  DISPATCH(h, { M4 p = to_m4(pose_io); r = P.tracking(p, to_m4(raycast_pose), k, icp_threshold, iterations, levels) ? 1 : 0; std::memcpy(pose_io, p.m, sizeof(p.m)); });
  return r;
}
void seo_get_pyramid(void* h, int level, float* depth, float* vertex, float* normal) {
  DISPATCH(h, {
    if (depth) std::memcpy(depth, P.scaled_depth[level].data(), sizeof(float) * P.scaled_depth[level].size());
    if (vertex) std::memcpy(vertex, P.input_vertex[level].data(), sizeof(V3) * P.input_vertex[level].size());
    if (normal) std::memcpy(normal, P.input_normal[level].data(), sizeof(V3) * P.input_normal[level].size());
  });
}
// N4: marching cubes; returns the triangle count, copies min(count, capacity) triangles (9 floats each)
long long seo_marching_cube(void* h, const int8_t* table, float* out, long long capacity) {
  std::vector<float> tri;
  DISPATCH(h, { marching_cube(P.map, table, tri); });
  const long long n = (long long)(tri.size() / 9);
  if (out) std::memcpy(out, tri.data(), sizeof(float) * 9 * (size_t)std::min(n, capacity));
  return n;
}
void seo_get_tracking(void* h, void* track_data /*W*H*32 B*/, float* reduction /*32*/) {
  DISPATCH(h, {
    if (track_data) std::memcpy(track_data, P.tracking_result.data(), sizeof(TrackData) * P.tracking_result.size());
    if (reduction) std::memcpy(reduction, P.reduction, sizeof(float) * 32);
  });
}
void seo_se3_exp(const float x[6], float out[16]) { M4 t = se3_exp(x); std::memcpy(out, t.m, sizeof(t.m)); }
int seo_solve6(const float vals[27], float x[6]) { return solve6(vals, x) ? 1 : 0; }

void seo_set_counting(void* h, int on) { DISPATCH(h, P.count = on != 0); }
void seo_reset_counters(void* h) { DISPATCH(h, P.ctr = Counters()); }
void seo_get_counters(void* h, uint64_t out[9]) {
  DISPATCH(h, {
    out[0] = P.ctr.n_get; out[1] = P.ctr.n_interp; out[2] = P.ctr.n_grad; out[3] = P.ctr.n_active; out[4] = P.ctr.n_nodes;
    out[5] = P.ctr.n_new_blocks; out[6] = P.ctr.n_new_nodes; out[7] = P.ctr.n_unique_keys; out[8] = P.ctr.n_keys_raw;
  });
}

}  // extern "C"

// vectorised checks used by the known-answer tests (a scalar ctypes call per voxel would be too slow)
extern "C" long long seo_morton_roundtrip_mismatches(int x0, int x1, int y0, int y1, int z0, int z1) {
  long long bad = 0;
  for (int z = z0; z < z1; ++z)
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const V3i v = morton_decode(morton_encode((uint64_t)x, (uint64_t)y, (uint64_t)z));
        bad += (v.x != x) | (v.y != y) | (v.z != z);
      }
  return bad;
}
