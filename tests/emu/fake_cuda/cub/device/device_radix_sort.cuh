// fake_cuda/cub: host stand-ins for the three CUB device algorithms the library calls (stable LSD radix sort on a bit
// range, exclusive prefix sum).  Same calling convention: a null temp-storage pointer asks for the size.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {
namespace emu_detail {
template <class K>
inline K key_bits(K k, int begin_bit, int end_bit) {
  const int w = end_bit - begin_bit;
  const K shifted = (K)(k >> begin_bit);
  return w >= (int)(8 * sizeof(K)) ? shifted : (K)(shifted & (((K)1 << w) - 1));
}
}  // namespace emu_detail

struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out,
                               int n, int begin_bit = 0, int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!tmp) {
      tmp_bytes = 16;
      return cudaSuccess;
    }
    std::vector<int> order((size_t)std::max(n, 0));
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      return emu_detail::key_bits(keys_in[a], begin_bit, end_bit) < emu_detail::key_bits(keys_in[b], begin_bit, end_bit);
    });
    std::vector<K> ko(order.size());
    std::vector<V> vo(order.size());
    for (size_t i = 0; i < order.size(); ++i) {
      ko[i] = keys_in[order[i]];
      if (vals_in) vo[i] = vals_in[order[i]];
    }
    for (size_t i = 0; i < order.size(); ++i) {
      keys_out[i] = ko[i];
      if (vals_in && vals_out) vals_out[i] = vo[i];
    }
    return cudaSuccess;
  }
  template <class K>
  static cudaError_t SortKeys(void* tmp, size_t& tmp_bytes, const K* keys_in, K* keys_out, int n, int begin_bit = 0,
                              int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!tmp) {
      tmp_bytes = 16;
      return cudaSuccess;
    }
    std::vector<K> k(keys_in, keys_in + std::max(n, 0));
    std::stable_sort(k.begin(), k.end(), [&](K a, K b) {
      return emu_detail::key_bits(a, begin_bit, end_bit) < emu_detail::key_bits(b, begin_bit, end_bit);
    });
    std::copy(k.begin(), k.end(), keys_out);
    return cudaSuccess;
  }
};
}  // namespace cub
