// Stand-in for cub::DeviceScan (CPU test tier, see ../../cuda_runtime.h).
#pragma once
#include <cuda_runtime.h>
namespace cub {
struct DeviceScan {
  template <class In, class Out>
  static cudaError_t ExclusiveSum(void* tmp, size_t& bytes, const In* in, Out* out, int n, cudaStream_t = nullptr) {
    if (!tmp) { bytes = 1; return cudaSuccess; }
    Out acc = 0;
    for (int i = 0; i < n; ++i) { const Out v = (Out)in[i]; out[i] = acc; acc += v; }
    return cudaSuccess;
  }
};
}  // namespace cub
