// DenseSLAMSystem.h -- the reference's pipeline class
// (se_denseslam/include/se/DenseSLAMSystem.h:58-411) with the same public methods, argument
// meaning and return conventions, implemented on top of the C ABI of include/se_b200.h: the map,
// the images and every kernel live on the GPU.
//
// As in the reference the field type is a compile-time choice: build with
// -DSE_FIELD_TYPE=SDF or -DSE_FIELD_TYPE=OFusion (se_denseslam/CMakeLists.txt:31-50).
//
// What differs from the reference, by design:
//  * getMap() cannot hand out a host se::Octree (the map is in HBM): it returns a snapshot
//    (se::MapSnapshot: blocks and nodes sorted by key, the comparable form of Octree::save,
//    se_core/include/se/octree.hpp:897-915).
//  * tracking() runs the ICP residual / reduction kernels on the GPU and the 6x6 solve + SE3
//    exponential on the host; its float reductions have a fixed order, so results are
//    reproducible but equal to the CPU reference only to rounding.
//  * errors the reference handles with exit(1) (bad size ratio, preprocessing.cpp:165-176)
//    do the same here, with the library's message.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "config.h"
#include "eigen_lite.h"

struct se_b200_map;

#ifndef SE_FIELD_TYPE
#define SE_FIELD_TYPE SDF
#endif
struct SDF { float x; float y; };                      // volume_traits.hpp:41-44
struct OFusion { float x; double y; };                 // volume_traits.hpp:62-65
typedef SE_FIELD_TYPE FieldType;

namespace se {
struct MapSnapshot {                                   // what Octree::save writes, sorted by key
  int size = 0;
  float dim = 0.f;
  std::vector<uint64_t> block_keys;
  std::vector<int32_t> block_coords;                   // 3 per block
  std::vector<FieldType> block_voxels;                 // 512 per block, x + 8 y + 64 z
  std::vector<uint64_t> node_codes;
  std::vector<uint32_t> node_sides;
  std::vector<FieldType> node_values;                  // 8 per node
  // The file format of Octree::save / Octree::load (se_core/include/se/octree.hpp:897-950,
  // io/se_serialise.hpp:54-99): int size, float dim, size_t n_nodes, {u64 code, u32 side, value_[8]} x n,
  // size_t n_blocks, {u64 code, int3 coords, voxel[512]} x n.  Records are written in key order (the
  // reference writes pool order, which is arbitrary there too).
  bool save(const std::string& filename) const;
  bool load(const std::string& filename);

  // A host mirror of the read accessors a caller of getMap() uses on the reference's se::Octree, over the records held here:
  //   fetch     Octree::fetch (octree.hpp:440-458): index of the block containing voxel (x, y, z) in block_keys, -1 if none
  //   get_fine  Octree::get_fine (:356-377): the voxel, initValue() where nothing is allocated
  //   interp    Octree::interp with select = .x (:541-563, interpolation/interp_gather.hpp:105-237): trilinear, voxel units
  // (voxels outside [0, size) read initValue(): the reference indexes out of bounds there).  Records need not be sorted.
  int fetch(int x, int y, int z) const;
  FieldType get_fine(int x, int y, int z) const;
  float interp(float x, float y, float z) const;
  static FieldType initValue();

 private:
  mutable std::vector<std::pair<uint64_t, int>> index_;      // (key, position in block_keys), sorted; built on the first query
  void build_index() const;
};
}  // namespace se

class DenseSLAMSystem {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  // DenseSLAMSystem.h:111-134.  Only .x() of the resolution / dimensions is used (cubic volume,
  // DenseSLAMSystem.cpp:123-125).
  DenseSLAMSystem(const Eigen::Vector2i& inputSize, const Eigen::Vector3i& volumeResolution,
                  const Eigen::Vector3f& volumeDimensions, const Eigen::Vector3f& initPose,
                  std::vector<int>& pyramid, const Configuration& config);
  DenseSLAMSystem(const Eigen::Vector2i& inputSize, const Eigen::Vector3i& volumeResolution,
                  const Eigen::Vector3f& volumeDimensions, const Eigen::Matrix4f& initPose,
                  std::vector<int>& pyramid, const Configuration& config);
  ~DenseSLAMSystem();
  DenseSLAMSystem(const DenseSLAMSystem&) = delete;
  DenseSLAMSystem& operator=(const DenseSLAMSystem&) = delete;

  // :147  mm -> m (+ sub-sampling to the computation size); `filterInput` only feeds tracking (N1)
  bool preprocessing(const unsigned short* inputDepth, const Eigen::Vector2i& inputSize, bool filterInput);
  // :173  ICP against the last raycast; returns checkPoseKernel's verdict (pose restored when it fails)
  bool tracking(const Eigen::Vector4f& k, float icp_threshold, unsigned tracking_rate, unsigned frame);
  // :193  runs when frame % integration_rate == 0 or frame <= 3 (DenseSLAMSystem.cpp:209)
  bool integration(const Eigen::Vector4f& k, unsigned integration_rate, float mu, unsigned frame);
  // :212  runs when frame > 2 (DenseSLAMSystem.cpp:195); latches raycast_pose_
  bool raycasting(const Eigen::Vector4f& k, float mu, unsigned frame);
  // :219-224
  void dump_volume(const std::string filename);
  void dump_mesh(const std::string filename);
  // :241  out = W*H*4 bytes RGBA, caller-owned; runs when frame % rate == 0 (DenseSLAMSystem.cpp:281)
  void renderVolume(unsigned char* out, const Eigen::Vector2i& outputSize, int frame, int rate,
                    const Eigen::Vector4f& k, float mu);
  void renderTrack(unsigned char* out, const Eigen::Vector2i& outputSize);   // :268
  void renderDepth(unsigned char* out, const Eigen::Vector2i& outputSize);   // :285

  void getMap(std::shared_ptr<se::MapSnapshot>& out);                         // :295 (see header comment)
  // inverse of getMap(): rebuild the device map from a snapshot (the role of Octree::load, octree.hpp:917-950)
  void setMap(const se::MapSnapshot& in);
  bool getTracked() { return tracked_; }
  bool getIntegrated() { return integrated_; }
  Eigen::Vector3f getPosition() {                                             // :318-325
    return Eigen::Vector3f(pose_(0, 3) - init_pose_.x(), pose_(1, 3) - init_pose_.y(), pose_(2, 3) - init_pose_.z());
  }
  Eigen::Vector3f getInitPos() { return init_pose_; }
  Eigen::Matrix4f getPose() { return pose_; }
  void setPose(const Eigen::Matrix4f pose) {                                  // :353-356: adds init_pose_
    pose_ = pose;
    pose_(0, 3) += init_pose_.x(); pose_(1, 3) += init_pose_.y(); pose_(2, 3) += init_pose_.z();
  }
  void setViewPose(Eigen::Matrix4f* value = nullptr) {                        // :363-372: keeps the caller's pointer
    if (value == nullptr) { viewPose_ = &pose_; need_render_ = false; }
    else { viewPose_ = value; need_render_ = true; }
  }
  Eigen::Matrix4f* getViewPose() { return viewPose_; }
  Eigen::Vector3f getModelDimensions() { return volume_dimension_; }
  Eigen::Vector3i getModelResolution() { return volume_resolution_; }
  Eigen::Vector2i getComputationResolution() { return computation_size_; }

  // -- additions (not in the reference) ----------------------------------------------------------
  // vertex / normal maps of the last raycast (W*H*3 floats each), for callers that consume them
  void getVertexNormal(std::vector<float>& vertex, std::vector<float>& normal);
  // extension (se_b200_set_render_target): raycasting() also renders the reuse-path image into `out` (device memory or
  // page-locked host memory, W*H*4 bytes; nullptr turns it off); renderVolume(out, ...) with the same pointer then only waits
  void setRenderTarget(unsigned char* out);
  // page-lock a caller-owned host buffer (se_b200_register_host_buffer): the reference application malloc()s its depth and
  // RGBA buffers (se_apps/src/benchmark.cpp:90-97), which makes every transfer a staged synchronous copy; registered once,
  // preprocessing() uploads asynchronously and renderVolume() / the render target write the image in place.  Unregister
  // before the memory is freed.
  static void registerHostBuffer(void* ptr, size_t bytes);
  static void unregisterHostBuffer(void* ptr);
  // per-stage device timing for stageMilliseconds() (off by default: the event records cost ~10 % of a frame)
  void enableStageTiming(bool on);
  // device time of the last run of a stage in ms (stands in for the TICK/TOCK samples, se_shared/timings.h)
  float stageMilliseconds(int stage);
  se_b200_map* handle() { return map_; }

 private:
  Eigen::Vector2i computation_size_;
  Eigen::Matrix4f pose_;
  Eigen::Matrix4f* viewPose_;
  Eigen::Vector3f volume_dimension_;
  Eigen::Vector3i volume_resolution_;
  std::vector<int> iterations_;
  bool tracked_ = false;
  bool integrated_ = false;
  Eigen::Vector3f init_pose_;
  float mu_;
  bool need_render_ = false;
  Configuration config_;
  Eigen::Matrix4f raycast_pose_;
  se_b200_map* map_ = nullptr;
};

// DenseSLAMSystem.h:418: declared by the reference, never defined there; here it really synchronises
void synchroniseDevices();
