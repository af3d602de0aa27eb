// hnsw_host.hpp — host-side state of a staged HNSW index (opaque cozo_gpu_hnsw_t).
#pragma once
#include <condition_variable>
#include <mutex>
#include <vector>

#include "hnsw_device.cuh"

struct HnswWorkspace {
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  uint32_t* vis = nullptr;
  size_t vis_words = 0;  // total words allocated
  uint32_t* vlog = nullptr;
  size_t vlog_words = 0;
  uint32_t* counter = nullptr;
  // staging buffers of the host-pointer API
  float* q = nullptr;
  size_t q_floats = 0;
  uint32_t* ids = nullptr;
  float* dist = nullptr;
  uint32_t* count = nullptr;
  uint32_t* qstats = nullptr;
  size_t out_rows = 0, out_k = 0;
  uint32_t* mask = nullptr;  // row filter verdicts of the host-pointer call
  size_t mask_words = 0;
  // pinned mirrors
  float* h_q = nullptr;
  size_t h_q_floats = 0;
};

struct cozo_gpu_hnsw {
  cozo::HnswDev dev{};
  // owned device buffers
  float* d_vec = nullptr;
  bool vec_owned = true;
  // F64 vector index (VecElementType::F64): payloads as doubles, searched by hnsw_f64.cu only
  double* d_vec64 = nullptr;
  uint32_t ld64 = 0;
  bool f64 = false;
  uint32_t* d_adj0 = nullptr;
  uint32_t* d_upper_off = nullptr;
  uint32_t* d_adj_up = nullptr;
  uint64_t up_rows = 0;
  // builder-only companions (distances of the stored edges, hnsw.rs:281-318 `dist`)
  float* d_adj0_dist = nullptr;
  float* d_adj_up_dist = nullptr;
  uint8_t* d_node_level = nullptr;
  uint32_t *d_deg0 = nullptr, *d_deg_up = nullptr;  // live out-degree of every row
  uint32_t* d_up_owner = nullptr;                   // node owning each upper-layer row
  uint8_t* d_dead = nullptr;                        // removed nodes (hnsw_remove)
  uint32_t cap_n = 0;                               // allocated rows (>= dev.n); 0 = staged, exact fit
  uint64_t cap_up = 0;
  bool build_state_ready = false;
  uint32_t ef_construction = 0;
  int keep_pruned = 0;
  int extend_candidates = 0;  // hnsw.rs:499-511: fidelity mode, one node per batch (hnsw_build.cu)
  uint64_t rng_state = 0x5EED0003ull;
  uint32_t n_live = 0;
  uint32_t borrowed_rows = 0;  // rows available in a borrowed vector buffer
  std::vector<uint8_t> live;
  // host copies
  std::vector<uint8_t> node_level;  // top layer index of each node (0 = layer 0 only)
  uint32_t n_levels = 1;
  uint32_t m_max0 = 0, m_max = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<HnswWorkspace*> pool;
  uint32_t n_workspaces = 0;  // created so far (bounded by the "hnsw.max_workspaces" option)
};

namespace cozo {
struct ScatterDest {
  uint32_t n_dest, slot;
  uint32_t* ids[COZO_GPU_MAX_PEERS];
  float* dist[COZO_GPU_MAX_PEERS];
};
int hnsw_launch_search(cozo_gpu_hnsw* h, HnswWorkspace* ws, const float* d_q, uint32_t B, uint32_t k, uint32_t ef,
                       double radius, uint32_t* d_ids, float* d_dist, uint32_t* d_count, uint32_t* d_qstats,
                       cudaStream_t stream, const ScatterDest* scatter, const uint32_t* d_filter_mask = nullptr);
HnswWorkspace* hnsw_acquire_ws(cozo_gpu_hnsw* h);
void hnsw_release_ws(cozo_gpu_hnsw* h, HnswWorkspace* ws);
int hnsw_ws_reserve(HnswWorkspace* ws, size_t vis_words, size_t vlog_words, cudaStream_t stream);
int hnsw_ensure_build_state(cozo_gpu_hnsw* h);
}  // namespace cozo
