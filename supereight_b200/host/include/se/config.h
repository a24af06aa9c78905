// config.h -- struct Configuration with the fields of the reference's
// se_denseslam/include/se/config.h:39-214 (same names, types and meaning), so that application
// code filling it (se_apps/include/default_parameters.h:195-466) compiles unchanged.
#pragma once
#include <string>
#include <vector>

#include "eigen_lite.h"

struct Configuration {
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  int compute_size_ratio = 1;        // -c
  int tracking_rate = 1;             // -t
  int integration_rate = 2;          // -r
  int rendering_rate = 4;            // -z
  Eigen::Vector3i volume_resolution; // -v
  Eigen::Vector3f volume_size;       // -s
  int voxel_block_size = 8;
  Eigen::Vector3f initial_pos_factor;  // -p
  std::vector<int> pyramid;          // -y
  std::string dump_volume_file;
  std::string input_file;
  std::string log_file;
  std::string groundtruth_file;
  Eigen::Matrix4f gt_transform;
  Eigen::Vector4f camera;            // -k
  bool camera_overrided = false;
  float mu = 0.1f;                   // -m
  int fps = 0;
  bool blocking_read = false;
  float icp_threshold = 1e-5f;
  bool no_gui = false;
  bool render_volume_fullsize = false;
  bool bilateralFilter = false;
  bool colouredVoxels = false;
  bool multiResolution = false;
  bool bayesian = false;
};
