// ref_capi.cpp -- TEST INFRASTRUCTURE ONLY.
// The reference's OWN pipeline (se_denseslam/src/DenseSLAMSystem.cpp and everything it includes, compiled where it lies
// under /root/reference, unmodified) behind the same `seo_*` C entry points as the oracle (se_oracle_capi.cpp), so the
// tests can run it beside the oracle and the CUDA path, and bench.py --impl reference can time it.
//
// The image has neither Eigen nor Sophus; the build uses the stand-in headers in oracle/ref_standin/ (see the header of
// ref_standin/Eigen/Dense for what that means for bit-level results).  Built by oracle/Makefile into oracle/_ref/ with
// -fno-access-control so this file can reach DenseSLAMSystem's private images (vertex_, normal_, float_depth_, ...); one
// library per field type, as in the reference (SE_FIELD_TYPE, se_denseslam/CMakeLists.txt:31-50).
// Nothing in the product may link or load this.
#include <perfstats.h>
PerfStats Stats;                       // the application defines it in the reference (DenseSLAMSystem.cpp:53 declares it extern)

#include <../src/DenseSLAMSystem.cpp>  // resolved through -I<reference>/se_denseslam/include

#include <omp.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

namespace {
constexpr int kField = std::is_same<FieldType, SDF>::value ? 0 : 1;
using Voxel = se::Octree<FieldType>::value_type;

struct Handle {
  DenseSLAMSystem* sys = nullptr;
  int W = 0, H = 0;
  std::vector<uint16_t> last_depth;
  int last_w = 0, last_h = 0;
  Eigen::Matrix4f view;                // storage for viewPose_
  se::Octree<FieldType>* tree = nullptr;   // seo_load_map: a stand-alone octree read by the reference's Octree::load (no pipeline around it)
};
Eigen::Matrix4f to_m4(const float* p) { Eigen::Matrix4f m; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m(r, c) = p[4 * r + c]; return m; }
Eigen::Vector4f to_k(const float* k) { return Eigen::Vector4f(k[0], k[1], k[2], k[3]); }
se::Octree<FieldType>& map_of(Handle* h) { return h->tree ? *h->tree : *h->sys->volume_._map_index; }
auto select_x = [](const Voxel& v) { return v.x; };
}  // namespace

extern "C" {

int seo_ref_field() { return kField; }

// SEO_REF_PYRAMID_LEVELS (environment, read at creation): pyramid depth of the pipelines created afterwards; default: the 3
// levels {10, 5, 4} of the reference's default configuration
void* seo_create(int field, int size, float dim, int W, int H) {
  if (field != kField) return nullptr;
  Handle* h = new Handle;
  h->W = W; h->H = H;
  Configuration config;
  config.mu = 0.1f;
  std::vector<int> pyramid = {10, 5, 4};
  if (const char* e = std::getenv("SEO_REF_PYRAMID_LEVELS")) pyramid.assign((size_t)std::max(1, std::atoi(e)), 4);
  Eigen::Matrix4f init = Eigen::Matrix4f::Identity();
  h->sys = new DenseSLAMSystem(Eigen::Vector2i(W, H), Eigen::Vector3i::Constant(size), Eigen::Vector3f::Constant(dim), init, pyramid, config);
  return h;
}
void seo_destroy(void* hh) { Handle* h = (Handle*)hh; delete h->sys; delete h->tree; delete h; }

// ---- N3: the reference's own map file writer / reader (octree.hpp:897-950, io/se_serialise.hpp:54-99) --------------
int seo_save_map(void* hh, const char* path) { map_of((Handle*)hh).save(path); return 0; }
// Octree::load into a fresh, stand-alone octree; the handle answers the block / node / point queries below.
// NB the reference reads the file's float `dim` into an int (octree.hpp:921-923), so the loaded tree's dim_ is the float's
// bit pattern converted to float; `dim_fix` > 0 repairs it (private member, this file is built with -fno-access-control)
// so that metric queries on the loaded tree are meaningful.
void* seo_load_map(const char* path, float dim_fix) {
  Handle* h = new Handle;
  h->tree = new se::Octree<FieldType>;
  h->tree->load(path);
  if (dim_fix > 0.f) h->tree->dim_ = dim_fix;
  return h;
}
int seo_map_size(void* hh) { return map_of((Handle*)hh).size(); }
float seo_map_dim(void* hh) { return map_of((Handle*)hh).dim(); }
void seo_set_omp_threads(int n) { omp_set_num_threads(n); }

// mm2metersKernel exit(1)s on a bad ratio (preprocessing.cpp:166-176): report it instead, as the oracle does
int seo_preprocess(void* hh, const uint16_t* in, int inW, int inH) {
  Handle* h = (Handle*)hh;
  if (inW < h->W || inH < h->H || inW % h->W != 0 || inH % h->H != 0 || inW / h->W != inH / h->H) return 1;
  h->last_depth.assign(in, in + (size_t)inW * inH); h->last_w = inW; h->last_h = inH;
  h->sys->preprocessing(in, Eigen::Vector2i(inW, inH), false);
  return 0;
}
void seo_filter_depth(void* hh, int filter, int levels) {
  Handle* h = (Handle*)hh;
  (void)levels;
  h->sys->preprocessing(h->last_depth.data(), Eigen::Vector2i(h->last_w, h->last_h), filter != 0);
}
void seo_set_depth(void* hh, const float* d) { Handle* h = (Handle*)hh; std::memcpy(h->sys->float_depth_.data(), d, sizeof(float) * h->W * h->H); }
void seo_get_depth(void* hh, float* d) { Handle* h = (Handle*)hh; std::memcpy(d, h->sys->float_depth_.data(), sizeof(float) * h->W * h->H); }

unsigned seo_integrate(void* hh, const float* pose, const float* k, float mu, unsigned frame) {
  Handle* h = (Handle*)hh;
  h->sys->pose_ = to_m4(pose);
  h->sys->integration(to_k(k), 1, mu, frame);
  return 0;
}
void seo_raycast(void* hh, const float* pose, const float* k, float mu) {
  Handle* h = (Handle*)hh;
  h->sys->pose_ = to_m4(pose);
  h->sys->raycasting(to_k(k), mu, 3);                          // frame > 2 (DenseSLAMSystem.cpp:194)
}
// DenseSLAMSystem::renderVolume (DenseSLAMSystem.cpp:274-290) with the `render` decision made by the caller
void seo_render_volume(void* hh, uint8_t* out, const float* viewpose, const float* k, float mu, float largestep, int render) {
  Handle* h = (Handle*)hh;
  DenseSLAMSystem& s = *h->sys;
  const Eigen::Matrix4f view = to_m4(viewpose);
  const float step = s.volume_dimension_.x() / s.volume_resolution_.x();
  renderVolumeKernel(s.volume_, out, Eigen::Vector2i(h->W, h->H), view * getInverseCameraMatrix(to_k(k)), nearPlane, farPlane * 2.0f, mu, step,
                     largestep, view.topRightCorner<3, 1>(), ambient, render != 0, s.vertex_, s.normal_);
}
void seo_render_depth(void* hh, uint8_t* out) { Handle* h = (Handle*)hh; h->sys->renderDepth(out, Eigen::Vector2i(h->W, h->H)); }
void seo_render_track(uint8_t* out, const int* result, int stride_ints, int W, int H) {
  (void)stride_ints;                                           // TrackData records, 8 ints each
  renderTrackKernel(out, reinterpret_cast<const TrackData*>(result), Eigen::Vector2i(W, H));
}
void seo_get_vertex(void* hh, float* out) { Handle* h = (Handle*)hh; std::memcpy(out, h->sys->vertex_.data(), sizeof(float) * 3 * h->W * h->H); }
void seo_get_normal(void* hh, float* out) { Handle* h = (Handle*)hh; std::memcpy(out, h->sys->normal_.data(), sizeof(float) * 3 * h->W * h->H); }
void seo_set_vertex_normal(void* hh, const float* v, const float* n) {
  Handle* h = (Handle*)hh;
  std::memcpy(h->sys->vertex_.data(), v, sizeof(float) * 3 * h->W * h->H);
  std::memcpy(h->sys->normal_.data(), n, sizeof(float) * 3 * h->W * h->H);
}

int seo_block_count(void* hh) { return (int)map_of((Handle*)hh).getBlockBuffer().size(); }
int seo_node_count(void* hh) { return (int)map_of((Handle*)hh).getNodesBuffer().size(); }
int seo_voxel_bytes(void*) { return (int)sizeof(Voxel); }

void seo_get_blocks_sorted(void* hh, uint64_t* keys, int* coords, uint8_t* active, void* data) {
  auto& buf = map_of((Handle*)hh).getBlockBuffer();
  const int n = (int)buf.size();
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return buf[a]->code_ < buf[b]->code_; });
  Voxel* d = (Voxel*)data;
  for (int i = 0; i < n; ++i) {
    se::VoxelBlock<FieldType>* b = buf[order[i]];
    if (keys) keys[i] = b->code_;
    const Eigen::Vector3i c = b->coordinates();
    if (coords) { coords[3 * i] = c.x(); coords[3 * i + 1] = c.y(); coords[3 * i + 2] = c.z(); }
    if (active) active[i] = b->active() ? 1 : 0;
    if (d) std::memcpy(d + (size_t)i * 512, b->getBlockRawPtr(), sizeof(Voxel) * 512);
  }
}
void seo_get_nodes_sorted(void* hh, uint64_t* codes, uint32_t* side, uint8_t* mask, void* values) {
  auto& buf = map_of((Handle*)hh).getNodesBuffer();
  const int n = (int)buf.size();
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return buf[a]->code_ < buf[b]->code_; });
  Voxel* d = (Voxel*)values;
  for (int i = 0; i < n; ++i) {
    se::Node<FieldType>* nd = buf[order[i]];
    if (codes) codes[i] = nd->code_;
    if (side) side[i] = nd->side_;
    if (mask) mask[i] = nd->children_mask_;
    if (d) std::memcpy(d + (size_t)i * 8, nd->value_, sizeof(Voxel) * 8);
  }
}

int seo_allocate(void* hh, const uint64_t* keys, int n) {
  std::vector<se::key_t> k(keys, keys + n);
  return map_of((Handle*)hh).allocate(k.data(), n) ? 1 : 0;
}
int seo_fetch(void* hh, int x, int y, int z) { return map_of((Handle*)hh).fetch(x, y, z) != nullptr; }
int seo_fetch_octant(void* hh, int x, int y, int z, int depth) { return map_of((Handle*)hh).fetch_octant(x, y, z, depth) != nullptr; }
uint64_t seo_fetch_octant_code(void* hh, int x, int y, int z, int depth) {
  se::Node<FieldType>* n = map_of((Handle*)hh).fetch_octant(x, y, z, depth);
  return n ? (uint64_t)n->code_ : ~0ull;
}
void seo_get_fine(void* hh, int x, int y, int z, double out[2]) { const Voxel v = map_of((Handle*)hh).get_fine(x, y, z); out[0] = v.x; out[1] = (double)v.y; }
void seo_get_coarse(void* hh, int x, int y, int z, double out[2]) { const Voxel v = map_of((Handle*)hh).get(x, y, z); out[0] = v.x; out[1] = (double)v.y; }
void seo_set_voxel(void* hh, int x, int y, int z, double vx, double vy) {
  Voxel v = map_of((Handle*)hh).get_fine(x, y, z);
  v.x = (float)vx; v.y = (decltype(v.y))vy;
  map_of((Handle*)hh).set(x, y, z, v);
}
float seo_interp(void* hh, float x, float y, float z) { return map_of((Handle*)hh).interp(Eigen::Vector3f(x, y, z), select_x); }
void seo_grad(void* hh, float x, float y, float z, float out[3]) {
  const Eigen::Vector3f g = map_of((Handle*)hh).grad(Eigen::Vector3f(x, y, z), select_x);
  out[0] = g.x(); out[1] = g.y(); out[2] = g.z();
}

int seo_ray_blocks(void* hh, const float* origin, const float* dir, float nearP, float farP, uint64_t* out, int max_out, float tinfo[3]) {
  se::ray_iterator<FieldType> it(map_of((Handle*)hh), Eigen::Vector3f(origin[0], origin[1], origin[2]), Eigen::Vector3f(dir[0], dir[1], dir[2]), nearP, farP);
  int n = 0;
  bool first = true;
  while (se::VoxelBlock<FieldType>* b = it.next()) {
    if (first && tinfo) { tinfo[0] = it.tmin(); tinfo[1] = it.tmax(); tinfo[2] = it.tcmin(); }
    first = false;
    if (n < max_out) out[n] = b->code_;
    ++n;
  }
  if (first && tinfo) { tinfo[0] = it.tmin(); tinfo[1] = it.tmax(); tinfo[2] = it.tcmin(); }
  return n;
}

// ---- N1 ------------------------------------------------------------------------------------------------
int seo_tracking(void* hh, float* pose_io, const float* raycast_pose, const float* k, float icp_threshold, const int* iterations, int levels) {
  Handle* h = (Handle*)hh;
  DenseSLAMSystem& s = *h->sys;
  s.pose_ = to_m4(pose_io);
  s.raycast_pose_ = to_m4(raycast_pose);
  for (int i = 0; i < levels && i < (int)s.iterations_.size(); ++i) s.iterations_[i] = iterations[i];
  const bool ok = s.tracking(to_k(k), icp_threshold, 1, 0);
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) pose_io[4 * r + c] = s.pose_(r, c);
  return ok ? 1 : 0;
}
void seo_get_pyramid(void* hh, int level, float* depth, float* vertex, float* normal) {
  DenseSLAMSystem& s = *((Handle*)hh)->sys;
  const size_t n = (size_t)s.scaled_depth_[level].width() * s.scaled_depth_[level].height();
  if (depth) std::memcpy(depth, s.scaled_depth_[level].data(), sizeof(float) * n);
  if (vertex) std::memcpy(vertex, s.input_vertex_[level].data(), sizeof(float) * 3 * n);
  if (normal) std::memcpy(normal, s.input_normal_[level].data(), sizeof(float) * 3 * n);
}
void seo_get_tracking(void* hh, void* track_data, float* reduction) {
  DenseSLAMSystem& s = *((Handle*)hh)->sys;
  if (track_data) std::memcpy(track_data, s.tracking_result_.data(), sizeof(TrackData) * s.tracking_result_.size());
  if (reduction) std::memcpy(reduction, s.reduction_output_.data(), sizeof(float) * 32);
}

// ---- N4: the reference's marching_cube with its own edge_tables.h; `table` is ignored ------------------------
long long seo_marching_cube(void* hh, const int8_t* table, float* out, long long capacity) {
  (void)table;
  std::vector<Triangle> mesh;
  auto inside = [](const Voxel& val) { return val.x < 0.f; };   // DenseSLAMSystem.cpp:305-313
  se::algorithms::marching_cube(map_of((Handle*)hh), select_x, inside, mesh);
  const long long n = (long long)mesh.size();
  for (long long i = 0; out && i < std::min(n, capacity); ++i)
    for (int v = 0; v < 3; ++v)
      for (int c = 0; c < 3; ++c) out[i * 9 + v * 3 + c] = mesh[i].vertexes[v](c);
  return n;
}

}  // extern "C"
