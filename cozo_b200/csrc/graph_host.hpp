// graph_host.hpp — host-side state of a staged graph (opaque cozo_gpu_graph_t).
#pragma once
#include <mutex>

#include "common.cuh"

namespace cozo {
struct PrState;  // PageRank's own layouts (pagerank.cu), built lazily by the first cozo_gpu_pagerank call
void pr_state_free(PrState* p);

// scoped device buffer
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  template <class T>
  T* as() {
    return static_cast<T*>(p);
  }
};

inline bool poisoned(const volatile int* p) { return p && *p != 0; }
}  // namespace cozo

struct cozo_gpu_graph {
  uint32_t n = 0;
  uint64_t m = 0;
  bool weighted = false;
  // CSR exactly as GraphBuilder::csr_layout(Sorted) lays it out (fixed_rule/mod.rs:192-195, 318-321)
  uint32_t *out_ptr = nullptr, *out_idx = nullptr, *in_ptr = nullptr, *in_idx = nullptr;
  float* out_w = nullptr;
  std::mutex pr_mu;
  cozo::PrState* pr = nullptr;
};
