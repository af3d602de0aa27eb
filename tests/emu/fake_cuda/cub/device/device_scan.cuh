#pragma once
#include <cuda_runtime.h>

namespace cub {
struct DeviceScan {
  template <class In, class Out>
  static cudaError_t ExclusiveSum(void* tmp, size_t& tmp_bytes, In in, Out out, int n, cudaStream_t = nullptr) {
    if (!tmp) {
      tmp_bytes = 16;
      return cudaSuccess;
    }
    auto run = decltype(+in[0]){};
    for (int i = 0; i < n; ++i) {  // in place is allowed (in == out)
      const auto v = in[i];
      out[i] = run;
      run += v;
    }
    return cudaSuccess;
  }
};
}  // namespace cub
