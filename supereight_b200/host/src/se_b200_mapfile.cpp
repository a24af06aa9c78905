// se-denseslam-<field>-b200-mapfile -- the map file without a device: reads and writes the format of the
// reference's Octree::save / Octree::load (se_core/include/se/octree.hpp:897-950, io/se_serialise.hpp:54-99) through the
// shim's se::MapSnapshot, the same code DenseSLAMSystem::getMap() / setMap() and se_b200_benchmark -b use.
//   copy <in> <out>    load, then save: a file written by the reference comes out byte for byte (records keep their order)
//   sort <in> <out>    load, sort nodes and blocks by key (the order getMap() exports), save
//   info <in>          size, dim, node and block counts
//   query <in> <points> <out>   points: lines "x y z" (floats, voxel units); out: "get_fine.x get_fine.y interp" per line, with
//                      get_fine at the truncated coordinates -- se::MapSnapshot's host mirror of Octree::get_fine / interp
// Needs no GPU: nothing here touches the device.
#include <algorithm>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <string>

#include "se/DenseSLAMSystem.h"

static void sort_by_key(se::MapSnapshot& s) {
  auto permute = [](auto& v, const std::vector<size_t>& order, size_t per) {
    std::remove_reference_t<decltype(v)> out(v.size());
    for (size_t i = 0; i < order.size(); ++i) std::copy(v.begin() + order[i] * per, v.begin() + (order[i] + 1) * per, out.begin() + i * per);
    v.swap(out);
  };
  std::vector<size_t> order(s.node_codes.size());
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return s.node_codes[a] < s.node_codes[b]; });
  permute(s.node_codes, order, 1); permute(s.node_sides, order, 1); permute(s.node_values, order, 8);
  order.resize(s.block_keys.size());
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return s.block_keys[a] < s.block_keys[b]; });
  permute(s.block_keys, order, 1); permute(s.block_coords, order, 3); permute(s.block_voxels, order, 512);
}

int main(int argc, char** argv) {
  const std::string cmd = argc > 1 ? argv[1] : "";
  if (!((cmd == "info" && argc == 3) || ((cmd == "copy" || cmd == "sort") && argc == 4) || (cmd == "query" && argc == 5))) {
    std::cerr << "usage: " << argv[0] << " copy|sort <in> <out>  |  info <in>  |  query <in> <points> <out>" << std::endl;
    return 2;
  }
  se::MapSnapshot s;
  if (!s.load(argv[2])) { std::cerr << "cannot read " << argv[2] << std::endl; return 1; }
  if (cmd == "info") {
    std::cout << "size " << s.size << " dim " << s.dim << " nodes " << s.node_codes.size() << " blocks " << s.block_keys.size() << std::endl;
    return 0;
  }
  if (cmd == "query") {
    std::ifstream pts(argv[3]);
    std::ofstream out(argv[4]);
    out.precision(17);
    float x, y, z;
    while (pts >> x >> y >> z) {
      const FieldType v = s.get_fine((int)x, (int)y, (int)z);
      out << v.x << " " << (double)v.y << " " << s.interp(x, y, z) << "\n";
    }
    return out ? 0 : 1;
  }
  if (cmd == "sort") sort_by_key(s);
  if (!s.save(argv[3])) { std::cerr << "cannot write " << argv[3] << std::endl; return 1; }
  return 0;
}
