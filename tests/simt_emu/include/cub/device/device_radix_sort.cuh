// Stand-in for cub::DeviceRadixSort (CPU test tier, see ../../cuda_runtime.h): same two-call protocol, std::stable_sort.
#pragma once
#include <cuda_runtime.h>
#include <numeric>
namespace cub {
struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n,
                               int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!tmp) { bytes = 1; return cudaSuccess; }
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    const K mask = end_bit - begin_bit >= (int)sizeof(K) * 8 ? ~K(0) : (((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (kin[a] & mask) < (kin[b] & mask); });
    std::vector<K> k(n); std::vector<V> v(n);
    for (int i = 0; i < n; ++i) { k[i] = kin[idx[i]]; v[i] = vin[idx[i]]; }
    for (int i = 0; i < n; ++i) { kout[i] = k[i]; vout[i] = v[i]; }
    return cudaSuccess;
  }
};
}  // namespace cub
