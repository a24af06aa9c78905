// se_b200_benchmark.cpp -- the frame loop of the reference's se_apps/src/benchmark.cpp:101-177 driving
// the GPU-backed DenseSLAMSystem: same stage order, same integration gate, same 14-column TSV log.
// Depth comes from a SLAMBench 1.0 ".raw" file (per frame: uint32 w, h, uint16 depth[w*h], uint32 w, h,
// uchar3 rgb[w*h]; se_apps/include/interface.h:384-426), poses from a text file with one row-major 4x4
// camera-to-world matrix per line (the role of the ground-truth file in `-g` mode,
// se_apps/src/mainQt.cpp:257-265; poses are relative to the initial position, as setPose expects).
//
//   se-denseslam-{sdf,ofusion}-b200-benchmark -i scene.raw -g poses.txt [-v 512] [-s 4.8] [-m 0.1] [-c 1]
//        [-r 1] [-z 1] [-p 0,0,0] [-k fx,fy,cx,cy] [-o log.tsv] [-d dump.bin] [-b map.bin] [-M mesh.vtk] [-n max_frames] [-t 0|1] [-f 0|1]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "se/DenseSLAMSystem.h"

namespace {
struct RawReader {
  FILE* f = nullptr;
  uint32_t w = 0, h = 0;
  bool open(const std::string& path) {
    f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    uint32_t s[2];
    if (std::fread(s, sizeof(uint32_t), 2, f) != 2) return false;
    w = s[0]; h = s[1];
    std::rewind(f);
    return w > 0 && h > 0;
  }
  bool next(uint16_t* depth) {
    uint32_t s[2];
    if (std::fread(s, sizeof(uint32_t), 2, f) != 2) return false;
    if (s[0] != w || s[1] != h) return false;
    if (std::fread(depth, sizeof(uint16_t), (size_t)w * h, f) != (size_t)w * h) return false;
    if (std::fread(s, sizeof(uint32_t), 2, f) != 2) return false;
    return std::fseek(f, (long)((size_t)s[0] * s[1] * 3), SEEK_CUR) == 0;      // RGB is not used by this path
  }
  ~RawReader() { if (f) std::fclose(f); }
};
std::vector<float> parse_floats(const char* s) {
  std::vector<float> v; std::stringstream ss(s); std::string tok;
  while (std::getline(ss, tok, ',')) v.push_back(std::strtof(tok.c_str(), nullptr));
  return v;
}
template <class T> void put(std::ofstream& os, const std::vector<T>& v) {
  const uint64_t n = v.size();
  os.write((const char*)&n, sizeof(n));
  os.write((const char*)v.data(), (std::streamsize)(n * sizeof(T)));
}
}  // namespace

int main(int argc, char** argv) {
  Configuration config;
  config.volume_resolution = Eigen::Vector3i(256, 256, 256);        // default_parameters.h:25-49
  config.volume_size = Eigen::Vector3f(2.f, 2.f, 2.f);
  config.initial_pos_factor = Eigen::Vector3f(0.f, 0.f, 0.f);
  config.camera = Eigen::Vector4f(481.2f, 480.f, 320.f, 240.f);
  config.pyramid = {10, 5, 4};
  config.integration_rate = 2; config.rendering_rate = 4; config.mu = 0.1f; config.compute_size_ratio = 1;
  std::string poses_file, dump_file, map_file, mesh_file, load_file;
  int max_frames = -1;
  bool use_tracking = false;      // -t 1: track with ICP after the first 4 frames instead of reading the pose file
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string a = argv[i]; const char* v = argv[i + 1];
    if (a == "-i") config.input_file = v;
    else if (a == "-g") poses_file = v;
    else if (a == "-v") { const int s = std::atoi(v); config.volume_resolution = Eigen::Vector3i(s, s, s); }
    else if (a == "-s") { const float s = std::strtof(v, nullptr); config.volume_size = Eigen::Vector3f(s, s, s); }
    else if (a == "-m") config.mu = std::strtof(v, nullptr);
    else if (a == "-c") config.compute_size_ratio = std::atoi(v);
    else if (a == "-r") config.integration_rate = std::atoi(v);
    else if (a == "-z") config.rendering_rate = std::atoi(v);
    else if (a == "-o") config.log_file = v;
    else if (a == "-d") dump_file = v;
    else if (a == "-M") mesh_file = v;
    else if (a == "-b") map_file = v;
    else if (a == "-L") load_file = v;                      // a map file (e.g. one written by the reference's Octree::save) to load, raycast and dump
    else if (a == "-n") max_frames = std::atoi(v);
    else if (a == "-t") use_tracking = std::atoi(v) != 0;
    else if (a == "-f") config.bilateralFilter = std::atoi(v) != 0;
    else if (a == "-p") { auto p = parse_floats(v); if (p.size() == 3) config.initial_pos_factor = Eigen::Vector3f(p[0], p[1], p[2]); }
    else if (a == "-k") { auto p = parse_floats(v); if (p.size() == 4) { config.camera = Eigen::Vector4f(p[0], p[1], p[2], p[3]); config.camera_overrided = true; } }
    else { std::cerr << "unknown option " << a << std::endl; return 2; }
  }
  RawReader reader;
  if (!reader.open(config.input_file)) { std::cerr << "cannot read " << config.input_file << std::endl; return 1; }
  std::ifstream poses(poses_file);
  if (!poses) { std::cerr << "cannot read poses file " << poses_file << std::endl; return 1; }
  std::ofstream logfile;
  std::ostream* logstream = &std::cout;
  if (!config.log_file.empty()) { logfile.open(config.log_file); logstream = &logfile; }
  logstream->precision(10);

  const Eigen::Vector3f init_pose(config.initial_pos_factor.x() * config.volume_size.x(), config.initial_pos_factor.y() * config.volume_size.y(),
                                  config.initial_pos_factor.z() * config.volume_size.z());
  const int cw = (int)reader.w / config.compute_size_ratio, ch = (int)reader.h / config.compute_size_ratio;
  const Eigen::Vector4f camera = config.camera / (float)config.compute_size_ratio;
  std::vector<uint16_t> inputDepth((size_t)reader.w * reader.h);
  std::vector<unsigned char> depthRender((size_t)cw * ch * 4), trackRender((size_t)cw * ch * 4), volumeRender((size_t)cw * ch * 4);

  DenseSLAMSystem pipeline(Eigen::Vector2i(cw, ch), config.volume_resolution, config.volume_size, init_pose, config.pyramid, config);
  // The reference's driver malloc()s these buffers (benchmark.cpp:90-97).  Page-locked once, the depth upload is an
  // asynchronous DMA and the images are written in place; with an image wanted every frame the raycast kernel renders
  // straight into volumeRender (setRenderTarget) and renderVolume() only waits for it.
  DenseSLAMSystem::registerHostBuffer(inputDepth.data(), inputDepth.size() * sizeof(uint16_t));
  for (auto* v : {&depthRender, &trackRender, &volumeRender}) DenseSLAMSystem::registerHostBuffer(v->data(), v->size());
  if (config.rendering_rate == 1) pipeline.setRenderTarget(volumeRender.data());

  using clk = std::chrono::steady_clock;
  std::chrono::time_point<clk> t[7];
  t[0] = clk::now();
  *logstream << "frame\tacquisition\tpreprocessing\ttracking\tintegration\traycasting\trendering\tcomputation\ttotal    \tX          \tY          \tZ         \ttracked   \tintegrated" << std::endl;
  logstream->setf(std::ios::fixed, std::ios::floatfield);
  unsigned frame = 0;
  while ((max_frames < 0 || (int)frame < max_frames) && reader.next(inputDepth.data())) {
    Eigen::Matrix4f gt = Eigen::Matrix4f::Identity();
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) if (!(poses >> gt(r, c))) { std::cerr << "poses file ended at frame " << frame << std::endl; return 1; }
    t[1] = clk::now();
    pipeline.preprocessing(inputDepth.data(), Eigen::Vector2i((int)reader.w, (int)reader.h), config.bilateralFilter);
    synchroniseDevices();
    t[2] = clk::now();
    bool tracked = true;
    if (use_tracking && frame > 3) tracked = pipeline.tracking(camera, config.icp_threshold, config.tracking_rate, frame);   // benchmark.cpp:124-125
    else pipeline.setPose(gt);                              // ground-truth mode (mainQt.cpp:257-265): the pose is given
    t[3] = clk::now();
    const Eigen::Matrix4f pose = pipeline.getPose();
    const float xt = pose(0, 3) - init_pose.x(), yt = pose(1, 3) - init_pose.y(), zt = pose(2, 3) - init_pose.z();
    bool integrated = false;
    if (tracked || frame <= 3) integrated = pipeline.integration(camera, config.integration_rate, config.mu, frame);
    synchroniseDevices();
    t[4] = clk::now();
    pipeline.raycasting(camera, config.mu, frame);
    synchroniseDevices();
    t[5] = clk::now();
    pipeline.renderDepth(depthRender.data(), Eigen::Vector2i(cw, ch));
    pipeline.renderTrack(trackRender.data(), Eigen::Vector2i(cw, ch));
    pipeline.renderVolume(volumeRender.data(), Eigen::Vector2i(cw, ch), (int)frame, config.rendering_rate, camera, 0.75f * config.mu);
    t[6] = clk::now();
    auto s = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    *logstream << frame << "\t" << s(t[0], t[1]) << "\t" << s(t[1], t[2]) << "\t" << s(t[2], t[3]) << "\t" << s(t[3], t[4]) << "\t" << s(t[4], t[5]) << "\t"
               << s(t[5], t[6]) << "\t" << s(t[1], t[5]) << "\t" << s(t[0], t[6]) << "\t" << xt << "\t" << yt << "\t" << zt << "\t" << tracked << "        \t" << integrated << std::endl;
    frame++;
    t[0] = clk::now();
  }
  if (!map_file.empty()) {                                  // benchmark.cpp:179-181 writes "test.bin" the same way
    std::shared_ptr<se::MapSnapshot> map;
    pipeline.getMap(map);
    if (!map->save(map_file)) { std::cerr << "cannot write " << map_file << std::endl; return 1; }
    // round trip: load the file into a fresh pipeline and check it reproduces the same snapshot
    se::MapSnapshot back;
    if (!back.load(map_file)) { std::cerr << "cannot read back " << map_file << std::endl; return 1; }
    DenseSLAMSystem second(Eigen::Vector2i(cw, ch), config.volume_resolution, config.volume_size, init_pose, config.pyramid, config);
    second.setMap(back);
    std::shared_ptr<se::MapSnapshot> again;
    second.getMap(again);
    const bool same = again->block_keys == map->block_keys && again->node_codes == map->node_codes &&
                      std::memcmp(again->block_voxels.data(), map->block_voxels.data(), map->block_voxels.size() * sizeof(FieldType)) == 0 &&
                      std::memcmp(again->node_values.data(), map->node_values.data(), map->node_values.size() * sizeof(FieldType)) == 0;
    std::cerr << "map file round trip: " << (same ? "identical" : "DIFFERENT") << " (" << map->block_keys.size() << " blocks, " << map->node_codes.size() << " nodes)" << std::endl;
    if (!same) return 3;
  }
  if (!mesh_file.empty()) pipeline.dump_mesh(mesh_file);   // what mainQt.cpp:166-169 does with --dump-volume
  if (!dump_file.empty()) {                                 // parity artefact for tests/test_gpu_host_shim.py
    std::shared_ptr<se::MapSnapshot> map;
    pipeline.getMap(map);
    std::vector<float> vertex, normal;
    pipeline.getVertexNormal(vertex, normal);
    std::ofstream os(dump_file, std::ios::binary);
    put(os, map->block_keys); put(os, map->block_voxels); put(os, map->node_codes); put(os, map->node_values);
    put(os, vertex); put(os, normal); put(os, volumeRender); put(os, depthRender); put(os, trackRender);
    if (!load_file.empty()) {
      // N3, the other direction: a map file from elsewhere -> se::MapSnapshot::load -> setMap() on a fresh pipeline -> raycast from
      // the stream's last pose; its content as re-exported by the device and the vertex / normal maps go to the dump
      se::MapSnapshot in;
      if (!in.load(load_file)) { std::cerr << "cannot read " << load_file << std::endl; return 1; }
      DenseSLAMSystem third(Eigen::Vector2i(cw, ch), config.volume_resolution, config.volume_size, init_pose, config.pyramid, config);
      third.setMap(in);
      Eigen::Matrix4f last = pipeline.getPose();
      last(0, 3) -= init_pose.x(); last(1, 3) -= init_pose.y(); last(2, 3) -= init_pose.z();
      third.setPose(last);
      third.raycasting(camera, config.mu, 3);
      std::shared_ptr<se::MapSnapshot> back;
      third.getMap(back);
      third.getVertexNormal(vertex, normal);
      put(os, back->block_keys); put(os, back->block_voxels); put(os, back->node_codes); put(os, back->node_values);
      put(os, vertex); put(os, normal);
    }
  }
  pipeline.setRenderTarget(nullptr);
  DenseSLAMSystem::unregisterHostBuffer(inputDepth.data());
  for (auto* v : {&depthRender, &trackRender, &volumeRender}) DenseSLAMSystem::unregisterHostBuffer(v->data());
  return 0;
}
