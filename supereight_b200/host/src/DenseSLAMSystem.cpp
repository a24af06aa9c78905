// DenseSLAMSystem.cpp -- the reference's pipeline object (se_denseslam/src/DenseSLAMSystem.cpp)
// as a thin host-side shim: stage gating, pose state and argument checks here, all computation
// behind the C ABI (include/se_b200.h) on the GPU.  No CPU fallback: if the library reports an
// error the process stops with its message, like the reference's exit(1) paths.
#include "se/DenseSLAMSystem.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <type_traits>

#include "se_b200.h"
#include <iostream>
#include <cmath>
#include <algorithm>

namespace {
constexpr int kFieldType = std::is_same<FieldType, SDF>::value ? SE_B200_SDF : SE_B200_OFUSION;
static_assert(sizeof(SDF) == sizeof(se_b200_sdf_voxel) && sizeof(OFusion) == sizeof(se_b200_ofusion_voxel), "voxel layout");

void pack(const Eigen::Matrix4f& m, float out[16]) {
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[4 * r + c] = m(r, c);
}
void pack(const Eigen::Vector4f& k, float out[4]) { for (int i = 0; i < 4; ++i) out[i] = k(i); }
[[noreturn]] void die(const char* where) {
  std::fprintf(stderr, "%s: %s\n", where, se_b200_last_error());
  std::exit(1);
}
#define SE_CHECK(call, where) do { if ((call) != SE_B200_OK) die(where); } while (0)

Eigen::Matrix4f translation(const Eigen::Vector3f& t) {      // se::math::toMatrix4f (math_utils.h:90-97)
  Eigen::Matrix4f m = Eigen::Matrix4f::Identity();
  m(0, 3) = t.x(); m(1, 3) = t.y(); m(2, 3) = t.z();
  return m;
}
se_b200_map* g_last_map = nullptr;
}  // namespace

DenseSLAMSystem::DenseSLAMSystem(const Eigen::Vector2i& inputSize, const Eigen::Vector3i& volumeResolution,
                                 const Eigen::Vector3f& volumeDimensions, const Eigen::Vector3f& initPose,
                                 std::vector<int>& pyramid, const Configuration& config)
    : DenseSLAMSystem(inputSize, volumeResolution, volumeDimensions, translation(initPose), pyramid, config) {}

DenseSLAMSystem::DenseSLAMSystem(const Eigen::Vector2i& inputSize, const Eigen::Vector3i& volumeResolution,
                                 const Eigen::Vector3f& volumeDimensions, const Eigen::Matrix4f& initPose,
                                 std::vector<int>& pyramid, const Configuration& config)
    : computation_size_(inputSize), config_(config) {
  init_pose_ = Eigen::Vector3f(initPose(0, 3), initPose(1, 3), initPose(2, 3));      // DenseSLAMSystem.cpp:77
  volume_dimension_ = volumeDimensions;
  volume_resolution_ = volumeResolution;
  mu_ = config.mu;
  pose_ = initPose;
  raycast_pose_ = initPose;
  iterations_ = pyramid;
  viewPose_ = &pose_;
  // device and pool sizes are not part of the reference's constructor: environment overrides, library defaults otherwise
  const char* dev = std::getenv("SE_B200_DEVICE");
  const char* mb = std::getenv("SE_B200_MAX_BLOCKS");
  SE_CHECK(se_b200_create(&map_, kFieldType, volumeResolution.x(), volumeDimensions.x(), inputSize.x(), inputSize.y(),
                          mb ? std::atoll(mb) : 0, 0, dev ? std::atoi(dev) : 0), "DenseSLAMSystem");
  g_last_map = map_;
}

DenseSLAMSystem::~DenseSLAMSystem() {
  if (g_last_map == map_) g_last_map = nullptr;
  se_b200_destroy(map_);
}

bool DenseSLAMSystem::preprocessing(const unsigned short* inputDepth, const Eigen::Vector2i& inputSize, bool filterInput) {
  // mm2metersKernel (preprocessing.cpp:161-188); a bad ratio prints "Invalid ratio." and exits, as there
  SE_CHECK(se_b200_preprocess_depth_host(map_, inputDepth, inputSize.x(), inputSize.y()), "preprocessing");
  // bilateralFilterKernel or a plain copy into scaled_depth_[0] (DenseSLAMSystem.cpp:132-139)
  if (!iterations_.empty()) SE_CHECK(se_b200_filter_depth(map_, filterInput ? 1 : 0, (int)iterations_.size()), "preprocessing");
  return true;
}

bool DenseSLAMSystem::tracking(const Eigen::Vector4f& k, float icp_threshold, unsigned tracking_rate, unsigned frame) {
  if (frame % tracking_rate != 0) return false;          // DenseSLAMSystem.cpp:146-147
  if (iterations_.empty()) return false;
  float p[16], rp[16], kk[4];
  pack(pose_, p); pack(raycast_pose_, rp); pack(k, kk);
  int ok = 0;
  SE_CHECK(se_b200_track(map_, p, rp, kk, icp_threshold, iterations_.data(), (int)iterations_.size(), &ok), "tracking");
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) pose_(r, c) = p[4 * r + c];
  tracked_ = ok != 0;
  return tracked_;
}

bool DenseSLAMSystem::integration(const Eigen::Vector4f& k, unsigned integration_rate, float mu, unsigned frame) {
  if (((frame % integration_rate) == 0) || (frame <= 3)) {   // DenseSLAMSystem.cpp:209
    float p[16], kk[4];
    pack(pose_, p); pack(k, kk);
    SE_CHECK(se_b200_integrate(map_, p, kk, mu, frame), "integration");
    integrated_ = true;
    return true;
  }
  integrated_ = false;
  return false;
}

bool DenseSLAMSystem::raycasting(const Eigen::Vector4f& k, float mu, unsigned frame) {
  if (frame > 2) {                                            // DenseSLAMSystem.cpp:195
    raycast_pose_ = pose_;
    float p[16], kk[4];
    pack(raycast_pose_, p); pack(k, kk);
    SE_CHECK(se_b200_raycast(map_, p, kk, mu), "raycasting");
    return true;
  }
  return false;
}

void DenseSLAMSystem::renderVolume(unsigned char* out, const Eigen::Vector2i&, int frame, int rate,
                                   const Eigen::Vector4f& k, float largestep) {
  if (frame % rate == 0) {                                    // DenseSLAMSystem.cpp:281
    float p[16], kk[4];
    pack(*viewPose_, p); pack(k, kk);
    const int reraycast = !viewPose_->isApprox(raycast_pose_);   // :287
    SE_CHECK(se_b200_render_volume_host(map_, out, p, kk, mu_, largestep, reraycast), "renderVolume");
  }
}

void DenseSLAMSystem::renderTrack(unsigned char* out, const Eigen::Vector2i&) {
  SE_CHECK(se_b200_render_track_host(map_, out, nullptr, 0), "renderTrack");     // the TrackData of the last tracking(), on the device
}

void DenseSLAMSystem::renderDepth(unsigned char* out, const Eigen::Vector2i&) {
  SE_CHECK(se_b200_render_depth_host(map_, out), "renderDepth");
}

void DenseSLAMSystem::dump_volume(const std::string) {}      // empty in the reference too (DenseSLAMSystem.cpp:270-272)

// DenseSLAMSystem.cpp:302-322: marching cubes on the device (se_b200_extract_mesh), then the file writeVtkMesh produces
// (commons.h:325-391, without point / cell data): one POINTS entry per triangle corner, "3 i i+1 i+2" polygons.
void DenseSLAMSystem::dump_mesh(const std::string filename) {
  int64_t n = 0;
  SE_CHECK(se_b200_extract_mesh(map_, &n), "dump_mesh");
  std::vector<float> tri((size_t)n * 9);
  SE_CHECK(se_b200_download_mesh(map_, tri.data(), n), "dump_mesh");
  std::ofstream f(filename.c_str());
  f << "# vtk DataFile Version 1.0" << std::endl;
  f << "vtk mesh generated from KFusion" << std::endl;
  f << "ASCII" << std::endl;
  f << "DATASET POLYDATA" << std::endl;
  f << "POINTS " << n * 3 << " FLOAT" << std::endl;
  for (int64_t i = 0; i < n * 3; ++i) f << tri[3 * i] << " " << tri[3 * i + 1] << " " << tri[3 * i + 2] << std::endl;
  f << "POLYGONS " << n << " " << n * 4 << std::endl;
  for (int64_t i = 0; i < n; ++i) f << "3 " << 3 * i << " " << 3 * i + 1 << " " << 3 * i + 2 << std::endl;
  f << std::endl;
}

void DenseSLAMSystem::getMap(std::shared_ptr<se::MapSnapshot>& out) {
  auto s = std::make_shared<se::MapSnapshot>();
  s->size = volume_resolution_.x();
  s->dim = volume_dimension_.x();
  int nb = 0, nn = 0;
  SE_CHECK(se_b200_block_count(map_, &nb), "getMap");
  SE_CHECK(se_b200_node_count(map_, &nn), "getMap");
  s->block_keys.resize(nb); s->block_coords.resize((size_t)nb * 3); s->block_voxels.resize((size_t)nb * 512);
  s->node_codes.resize(nn); s->node_sides.resize(nn); s->node_values.resize((size_t)nn * 8);
  SE_CHECK(se_b200_download_blocks_sorted(map_, s->block_keys.data(), s->block_coords.data(), nullptr, s->block_voxels.data()), "getMap");
  SE_CHECK(se_b200_download_nodes_sorted(map_, s->node_codes.data(), s->node_sides.data(), nullptr, s->node_values.data()), "getMap");
  out = s;
}

void DenseSLAMSystem::setMap(const se::MapSnapshot& in) {
  SE_CHECK(se_b200_upload_nodes(map_, in.node_codes.data(), in.node_values.data(), (int)in.node_codes.size()), "setMap");
  SE_CHECK(se_b200_upload_blocks(map_, in.block_keys.data(), in.block_voxels.data(), (int)in.block_keys.size()), "setMap");
}

bool se::MapSnapshot::save(const std::string& filename) const {
  std::ofstream os(filename, std::ios::binary);
  if (!os) return false;
  os.write((const char*)&size, sizeof(int));
  os.write((const char*)&dim, sizeof(float));
  size_t n = node_codes.size();
  os.write((const char*)&n, sizeof(size_t));
  for (size_t i = 0; i < n; ++i) {
    os.write((const char*)&node_codes[i], sizeof(uint64_t));
    os.write((const char*)&node_sides[i], sizeof(uint32_t));
    os.write((const char*)&node_values[8 * i], sizeof(FieldType) * 8);
  }
  n = block_keys.size();
  os.write((const char*)&n, sizeof(size_t));
  for (size_t i = 0; i < n; ++i) {
    os.write((const char*)&block_keys[i], sizeof(uint64_t));
    os.write((const char*)&block_coords[3 * i], sizeof(int32_t) * 3);
    os.write((const char*)&block_voxels[512 * i], sizeof(FieldType) * 512);
  }
  return (bool)os;
}

bool se::MapSnapshot::load(const std::string& filename) {
  std::ifstream is(filename, std::ios::binary);
  if (!is) return false;
  is.read((char*)&size, sizeof(int));
  is.read((char*)&dim, sizeof(float));       // the reference reads dim into an int (octree.hpp:921-923); the file holds a float
  size_t n = 0;
  is.read((char*)&n, sizeof(size_t));
  node_codes.resize(n); node_sides.resize(n); node_values.resize(8 * n);
  for (size_t i = 0; i < n; ++i) {
    is.read((char*)&node_codes[i], sizeof(uint64_t));
    is.read((char*)&node_sides[i], sizeof(uint32_t));
    is.read((char*)&node_values[8 * i], sizeof(FieldType) * 8);
  }
  is.read((char*)&n, sizeof(size_t));
  block_keys.resize(n); block_coords.resize(3 * n); block_voxels.resize(512 * n);
  for (size_t i = 0; i < n; ++i) {
    is.read((char*)&block_keys[i], sizeof(uint64_t));
    is.read((char*)&block_coords[3 * i], sizeof(int32_t) * 3);
    is.read((char*)&block_voxels[512 * i], sizeof(FieldType) * 512);
  }
  return (bool)is;
}

// ---- host mirror of Octree::fetch / get_fine / interp over a snapshot -------------------------------------------------
namespace {
uint64_t spread3(uint64_t v) {                       // morton_utils.hpp:37-49: one coordinate's bits to every third position
  uint64_t x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
int log2i(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
}  // namespace

FieldType se::MapSnapshot::initValue() {              // volume_traits.hpp:41-72
  FieldType v{};
  v.x = std::is_same<FieldType, SDF>::value ? 1.f : 0.f;
  v.y = 0;
  return v;
}

void se::MapSnapshot::build_index() const {
  index_.resize(block_keys.size());
  for (size_t i = 0; i < block_keys.size(); ++i) index_[i] = std::make_pair(block_keys[i], (int)i);
  std::sort(index_.begin(), index_.end());
}

int se::MapSnapshot::fetch(int x, int y, int z) const {
  if (x < 0 || y < 0 || z < 0 || x >= size || y >= size || z >= size) return -1;
  if (index_.size() != block_keys.size()) build_index();
  const int max_level = log2i(size), level = max_level - 3;              // blocks live at the leaves level (octree.hpp:205-212)
  const uint64_t morton = spread3((uint64_t)(x & ~7)) | (spread3((uint64_t)(y & ~7)) << 1) | (spread3((uint64_t)(z & ~7)) << 2);
  const uint64_t key = morton | (uint64_t)level;                         // (the block's low corner has no bits below the level's mask)
  auto it = std::lower_bound(index_.begin(), index_.end(), std::make_pair(key, -1));
  return (it != index_.end() && it->first == key) ? it->second : -1;
}

FieldType se::MapSnapshot::get_fine(int x, int y, int z) const {
  const int b = fetch(x, y, z);
  if (b < 0) return initValue();
  return block_voxels[(size_t)b * 512 + (x & 7) + 8 * (y & 7) + 64 * (z & 7)];
}

float se::MapSnapshot::interp(float px, float py, float pz) const {
  const float flx = std::floor(px), fly = std::floor(py), flz = std::floor(pz);
  const float fx = px - flx, fy = py - fly, fz = pz - flz;
  const int bx = std::max((int)flx, 0), by = std::max((int)fly, 0), bz = std::max((int)flz, 0);
  float p[8];
  for (int i = 0; i < 8; ++i) p[i] = get_fine(bx + (i & 1), by + ((i >> 1) & 1), bz + ((i >> 2) & 1)).x;    // empty().x == initValue().x
  return (((p[0] * (1 - fx) + p[1] * fx) * (1 - fy) + (p[2] * (1 - fx) + p[3] * fx) * fy) * (1 - fz)
        + ((p[4] * (1 - fx) + p[5] * fx) * (1 - fy) + (p[6] * (1 - fx) + p[7] * fx) * fy) * fz);
}

void DenseSLAMSystem::getVertexNormal(std::vector<float>& vertex, std::vector<float>& normal) {
  const size_t n = (size_t)computation_size_.x() * computation_size_.y() * 3;
  vertex.resize(n); normal.resize(n);
  SE_CHECK(se_b200_download_vertex_normal(map_, vertex.data(), normal.data()), "getVertexNormal");
}

void DenseSLAMSystem::setRenderTarget(unsigned char* out) {
  SE_CHECK(se_b200_set_render_target(map_, out), "setRenderTarget");
}

void DenseSLAMSystem::registerHostBuffer(void* ptr, size_t bytes) {
  if (se_b200_register_host_buffer(ptr, bytes) != SE_B200_OK) { std::cerr << "DenseSLAMSystem: " << se_b200_last_error() << std::endl; exit(1); }
}
void DenseSLAMSystem::unregisterHostBuffer(void* ptr) { se_b200_unregister_host_buffer(ptr); }
void DenseSLAMSystem::enableStageTiming(bool on) { SE_CHECK(se_b200_set_stage_timing(map_, on ? 1 : 0), "enableStageTiming"); }

float DenseSLAMSystem::stageMilliseconds(int stage) {
  float ms = 0.f;
  SE_CHECK(se_b200_elapsed_ms(map_, stage, &ms), "stageMilliseconds");
  return ms;
}

void synchroniseDevices() { if (g_last_map) se_b200_sync(g_last_map); }
