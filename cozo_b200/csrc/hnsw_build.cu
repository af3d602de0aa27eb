// hnsw_build.cu — index construction on the device.
//
// Batched variant of SessionTx::hnsw_put_vector (runtime/hnsw.rs:155-375):
// nodes are inserted in id order like create_hnsw_index's bulk insert
// (runtime/relation.rs:1176-1185), a batch at a time.  Every node of a batch
//   K1  searches the graph as it stood before the batch: ef=1 on the layers
//       above its own (hnsw.rs:219-229), ef_construction on its own layers
//       (hnsw.rs:242-256; found_nn carries over between layers);
//   K2  picks neighbours with hnsw_select_neighbours_heuristic (hnsw.rs:470-538),
//       capped at m_max(layer) (hnsw.rs:243-247), writes its own adjacency row
//       (the out edges, hnsw.rs:281-298) and queues the in edges (hnsw.rs:300-318);
//   K3  sorts the queued in edges by (layer, target);
//   K4  appends them to the target rows, shrinking an over-full row with the
//       same heuristic over the stored distances (hnsw_shrink_neighbour,
//       hnsw.rs:376-469).
// Differences from the reference, all consequences of batching: nodes of one
// batch do not see each other during K1, and a row that receives several in
// edges in one batch is shrunk once per chunk of 32 arrivals rather than once
// per arrival.  Soft-deleted (`ignore_link`) rows are not materialised: the
// staged layout only holds what hnsw_get_neighbours would return.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "hnsw_host.hpp"
#include "hnsw_build_kernels.cuh"

namespace cozo {

// thin __global__ wrappers around the bodies in hnsw_build_kernels.cuh
template <int NV, int METRIC>
__global__ void __launch_bounds__(128, 4) build_search_kernel(HnswDev g, BuildDev b, BatchParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  build_search_body<NV, METRIC>(g, b, p, smem);
}
template <int NV, int METRIC>
__global__ void __launch_bounds__(128) build_select_kernel(HnswDev g, BuildDev b, BatchParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  build_select_body<NV, METRIC>(g, b, p, smem);
}
template <int NV, int METRIC>
__global__ void __launch_bounds__(128) build_link_kernel(HnswDev g, BuildDev b, const unsigned long long* keys,
                                                         const uint32_t* perm, const uint32_t* req_src,
                                                         const float* req_d, uint32_t nreq, const uint32_t* heads,
                                                         uint32_t nheads) {
  extern __shared__ __align__(128) uint8_t smem[];
  build_link_body<NV, METRIC>(g, b, keys, perm, req_src, req_d, nreq, heads, nheads, smem);
}

using K1Fn = void (*)(HnswDev, BuildDev, BatchParams);
using K4Fn = void (*)(HnswDev, BuildDev, const unsigned long long*, const uint32_t*, const uint32_t*, const float*,
                      uint32_t, const uint32_t*, uint32_t);
using KdFn = void (*)(HnswDev, BuildDev, const uint32_t*, uint32_t, int);

template <int NV>
struct BuildKernels {
  static void get(int metric, K1Fn& k1, K1Fn& k2, K4Fn& k4, KdFn& kd) {
    switch (metric) {
      case COZO_GPU_L2:
        k1 = build_search_kernel<NV, COZO_GPU_L2>;
        k2 = build_select_kernel<NV, COZO_GPU_L2>;
        k4 = build_link_kernel<NV, COZO_GPU_L2>;
        kd = build_edge_dist_kernel<NV, COZO_GPU_L2>;
        break;
      case COZO_GPU_COSINE:
        k1 = build_search_kernel<NV, COZO_GPU_COSINE>;
        k2 = build_select_kernel<NV, COZO_GPU_COSINE>;
        k4 = build_link_kernel<NV, COZO_GPU_COSINE>;
        kd = build_edge_dist_kernel<NV, COZO_GPU_COSINE>;
        break;
      default:
        k1 = build_search_kernel<NV, COZO_GPU_IP>;
        k2 = build_select_kernel<NV, COZO_GPU_IP>;
        k4 = build_link_kernel<NV, COZO_GPU_IP>;
        kd = build_edge_dist_kernel<NV, COZO_GPU_IP>;
    }
  }
};

static int pick_build_kernels(const HnswDev& g, K1Fn& k1, K1Fn& k2, K4Fn& k4, KdFn& kd) {
  uint32_t need = (g.ld / 4 + 31) / 32;
  if (need <= 1) BuildKernels<1>::get(g.metric, k1, k2, k4, kd);
  else if (need <= 2) BuildKernels<2>::get(g.metric, k1, k2, k4, kd);
  else if (need <= 4) BuildKernels<4>::get(g.metric, k1, k2, k4, kd);
  else if (need <= 6) BuildKernels<6>::get(g.metric, k1, k2, k4, kd);
  else if (need <= 8) BuildKernels<8>::get(g.metric, k1, k2, k4, kd);
  else if (need <= 16) BuildKernels<16>::get(g.metric, k1, k2, k4, kd);
  else return set_error(COZO_GPU_EUNSUP, "vec_dim %u exceeds the supported maximum 2048", g.dim);
  return 0;
}

struct SplitMix64 {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

#define H_CUDA(call)                                                                                \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return set_error(_e == cudaErrorMemoryAllocation ? COZO_GPU_ENOMEM : COZO_GPU_ECUDA,           \
                       "%s failed: %s (line %d)", #call, cudaGetErrorString(_e), __LINE__);          \
  } while (0)

// grow one device array to `new_elems` elements, keeping `old_elems`, filling the tail with `fill`
template <class T>
static int grow(T*& p, size_t old_elems, size_t new_elems, int fill_byte) {
  T* np = nullptr;
  H_CUDA(cudaMalloc(&np, std::max<size_t>(new_elems, 1) * sizeof(T)));
  if (old_elems && p) H_CUDA(cudaMemcpy(np, p, old_elems * sizeof(T), cudaMemcpyDeviceToDevice));
  if (new_elems > old_elems)
    H_CUDA(cudaMemset(reinterpret_cast<char*>(np) + old_elems * sizeof(T), fill_byte, (new_elems - old_elems) * sizeof(T)));
  if (p) cudaFree(p);
  p = np;
  return 0;
}

// capacity management: rows [0,n) / [0,up_rows) are live, the rest is pre-initialised slack
int hnsw_reserve(cozo_gpu_hnsw* h, uint32_t need_n, uint64_t need_up) {
  HnswDev& g = h->dev;
  if (need_n > h->cap_n) {
    uint32_t nc = std::max<uint32_t>(need_n, (uint32_t)std::min<uint64_t>(0x7FFFFFFEull, (uint64_t)h->cap_n * 3 / 2));
    if (!h->vec_owned) {
      if (need_n > h->borrowed_rows) {  // borrowed vectors cannot grow in place: take a private copy
        float* nv = nullptr;
        H_CUDA(cudaMalloc(&nv, (size_t)nc * g.ld * 4));
        if (g.n) H_CUDA(cudaMemcpy(nv, h->d_vec, (size_t)g.n * g.ld * 4, cudaMemcpyDeviceToDevice));
        H_CUDA(cudaMemset(nv + (size_t)g.n * g.ld, 0, (size_t)(nc - g.n) * g.ld * 4));
        h->d_vec = nv;
        h->vec_owned = true;
      }
    } else {
      int rc = grow(h->d_vec, (size_t)h->cap_n * g.ld, (size_t)nc * g.ld, 0);
      if (rc) return rc;
    }
    int rc = grow(h->d_adj0, (size_t)h->cap_n * g.s0, (size_t)nc * g.s0, 0xFF);
    if (!rc) rc = grow(h->d_adj0_dist, (size_t)h->cap_n * g.s0, (size_t)nc * g.s0, 0);
    if (!rc) rc = grow(h->d_deg0, h->cap_n, nc, 0);
    if (!rc) rc = grow(h->d_upper_off, h->cap_n, nc, 0xFF);
    if (!rc) rc = grow(h->d_node_level, h->cap_n, nc, 0);
    if (!rc) rc = grow(h->d_dead, h->cap_n, nc, 0);
    if (rc) return rc;
    h->cap_n = nc;
  }
  if (need_up > h->cap_up || !h->d_adj_up) {
    uint64_t uc = std::max<uint64_t>(std::max<uint64_t>(need_up, 1), h->cap_up * 3 / 2);
    int rc = grow(h->d_adj_up, (size_t)h->cap_up * g.su, (size_t)uc * g.su, 0xFF);
    if (!rc) rc = grow(h->d_adj_up_dist, (size_t)h->cap_up * g.su, (size_t)uc * g.su, 0);
    if (!rc) rc = grow(h->d_deg_up, (size_t)h->cap_up, (size_t)uc, 0);
    if (!rc) rc = grow(h->d_up_owner, (size_t)h->cap_up, (size_t)uc, 0xFF);
    if (rc) return rc;
    h->cap_up = uc;
  }
  g.vec = h->d_vec;
  g.adj0 = h->d_adj0;
  g.adj_up = h->d_adj_up;
  g.upper_off = h->d_upper_off;
  return 0;
}

static BuildDev build_dev(cozo_gpu_hnsw* h) {
  BuildDev b{};
  b.adj0 = h->d_adj0;
  b.adj0_d = h->d_adj0_dist;
  b.deg0 = h->d_deg0;
  b.adj_up = h->d_adj_up;
  b.adj_up_d = h->d_adj_up_dist;
  b.deg_up = h->d_deg_up;
  b.node_level = h->d_node_level;
  b.m_max0 = h->m_max0;
  b.m_max = h->m_max;
  b.keep_pruned = h->keep_pruned;
  b.extend = 0;
  b.upper_off = h->d_upper_off;
  return b;
}

// An index staged from host CSR carries no degrees / edge distances: derive them once.
int hnsw_ensure_build_state(cozo_gpu_hnsw* h) {
  if (h->f64)
    return set_error(COZO_GPU_EUNSUP, "F64 indexes are search-only on the device: re-stage after a mutation");
  if (h->build_state_ready) return 0;
  HnswDev& g = h->dev;
  // staging allocated exactly n rows; move to the growable layout
  if (h->cap_n == 0) {
    h->cap_n = g.n;
    // capacity == live rows exactly: the arrays staging made have no slack, and hnsw_reserve's grow() treats
    // everything below the capacity as initialised (a capacity of 1 with 0 live upper rows would carry one
    // row of uninitialised ids and degrees into the grown arrays)
    h->cap_up = h->up_rows;
    uint32_t* p32 = nullptr;
    float* pf = nullptr;
    uint8_t* p8 = nullptr;
    const size_t up_alloc = std::max<size_t>((size_t)h->cap_up, 1);
    H_CUDA(cudaMalloc(&pf, std::max<size_t>((size_t)h->cap_n * g.s0, 1) * 4));
    h->d_adj0_dist = pf;
    H_CUDA(cudaMalloc(&pf, up_alloc * g.su * 4));
    h->d_adj_up_dist = pf;
    H_CUDA(cudaMalloc(&p32, std::max<size_t>(h->cap_n, 1) * 4));
    h->d_deg0 = p32;
    H_CUDA(cudaMalloc(&p32, up_alloc * 4));
    h->d_deg_up = p32;
    H_CUDA(cudaMalloc(&p32, up_alloc * 4));
    h->d_up_owner = p32;
    H_CUDA(cudaMalloc(&p8, std::max<size_t>(h->cap_n, 1)));
    h->d_node_level = p8;
    H_CUDA(cudaMalloc(&p8, std::max<size_t>(h->cap_n, 1)));
    h->d_dead = p8;
    H_CUDA(cudaMemset(h->d_dead, 0, std::max<size_t>(h->cap_n, 1)));
    if (g.n) H_CUDA(cudaMemcpy(h->d_node_level, h->node_level.data(), g.n, cudaMemcpyHostToDevice));
    std::vector<uint32_t> owner(up_alloc, NONE);
    uint64_t r = 0;
    for (uint32_t i = 0; i < g.n; ++i)
      for (uint32_t L = 0; L < h->node_level[i]; ++L) owner[r++] = i;
    H_CUDA(cudaMemcpy(h->d_up_owner, owner.data(), up_alloc * 4, cudaMemcpyHostToDevice));
    h->live.assign(g.n, 1);
    h->n_live = g.n;
  }
  K1Fn k1, k2;
  K4Fn k4;
  KdFn kd;
  int rc = pick_build_kernels(g, k1, k2, k4, kd);
  if (rc) return rc;
  BuildDev b = build_dev(h);
  if (g.n) kd<<<(g.n + 3) / 4, 128>>>(g, b, nullptr, g.n, 0);
  if (h->up_rows) kd<<<(uint32_t)((h->up_rows + 3) / 4), 128>>>(g, b, h->d_up_owner, (uint32_t)h->up_rows, 1);
  H_CUDA(cudaGetLastError());
  H_CUDA(cudaDeviceSynchronize());
  h->build_state_ready = true;
  return 0;
}

// The batched insertion loop for nodes [begin, end) whose vectors, levels and upper_off are in place.
int hnsw_insert_range(cozo_gpu_hnsw* h, uint32_t begin, uint32_t end, uint32_t max_batch) {
  const DeviceInfo& di = device_info();
  HnswDev& g = h->dev;
  if (begin >= end) return 0;
  K1Fn k1, k2;
  K4Fn k4;
  KdFn kd;
  int rc = pick_build_kernels(g, k1, k2, k4, kd);
  if (rc) return rc;
  BuildDev b = build_dev(h);
  const uint32_t ef_c = h->ef_construction;
  const bool extend = h->extend_candidates != 0;
  if (!max_batch) max_batch = 8192;
  // extend_candidates reads the rows of OTHER nodes while selecting, so its result depends on the order in which
  // the reference links the new node's neighbours: one node per batch, in-edges linked one at a time in selection order
  if (extend) max_batch = 1;
  uint32_t wpc = (uint32_t)get_option("hnsw.warps_per_cta", 4);
  wpc = std::min(4u, std::max(1u, wpc));
  uint32_t ns = (uint32_t)get_option("hnsw.stages", 4);
  ns = std::min(32u, std::max(1u, ns));
  SmemLayout lay = make_layout(ef_c, ns, g.ld);
  size_t smem1 = (size_t)lay.warp_bytes * wpc;
  while (smem1 > di.smem_optin && wpc > 1) {
    wpc >>= 1;
    smem1 = (size_t)lay.warp_bytes * wpc;
  }
  if (smem1 > di.smem_optin) return set_error(COZO_GPU_EUNSUP, "ef_construction=%u does not fit shared memory", ef_c);
  H_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)di.smem_optin));
  int cps = 0;
  H_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, k1, wpc * 32, smem1));
  if (cps < 1) return set_error(COZO_GPU_ECUDA, "build kernel does not fit on an SM");
  const uint32_t max_grid1 = (uint32_t)di.sm_count * (uint32_t)cps;
  const uint32_t mcap = std::max(h->m_max0, h->m_max);
  const size_t smem2 = (size_t)4 * 2 * mcap * 4;
  const size_t smem4 = (size_t)4 * (4 * (mcap + 32) + 2 * mcap) * 4;

  struct Scratch {
    uint32_t *coff = nullptr, *list_node = nullptr, *list_level = nullptr;
    float* cand_d = nullptr;
    uint32_t *cand_id = nullptr, *cand_cnt = nullptr;
    unsigned long long *req_key = nullptr, *req_key2 = nullptr;
    uint32_t *req_src = nullptr, *perm = nullptr, *perm2 = nullptr, *heads = nullptr, *counters = nullptr;
    float* req_d = nullptr;
    void* cub_tmp = nullptr;
    unsigned long long* ext_keys = nullptr;
    float* ext_d = nullptr;
    uint32_t* ext_id = nullptr;
    cozo_gpu_hnsw* h = nullptr;
    HnswWorkspace* ws = nullptr;
    ~Scratch() {
      void* ptrs[] = {coff, list_node, list_level, cand_d, cand_id, cand_cnt, req_key, req_key2,
                      req_src, perm, perm2, heads, counters, req_d, cub_tmp, ext_keys, ext_d, ext_id};
      for (void* p : ptrs)
        if (p) cudaFree(p);
      if (ws) hnsw_release_ws(h, ws);
    }
  } sc;
  sc.h = h;
  sc.ws = hnsw_acquire_ws(h);
  if (!sc.ws) return COZO_GPU_ECUDA;
  HnswWorkspace* ws = sc.ws;
  cudaStream_t st = ws->stream;
  const uint32_t nwords = round_up((g.n + 31) / 32, 4);  // the search may visit ANY indexed id, not only ids < end
  const uint32_t logcap = std::min<uint32_t>(65536u, std::max<uint32_t>(4096u, 64u * ef_c));
  {
    size_t slots = (size_t)max_grid1 * wpc;
    rc = hnsw_ws_reserve(ws, slots * nwords, slots * logcap, st);
    if (rc) return rc;
  }
  max_batch = std::min<uint32_t>(max_batch, end - begin);
  uint32_t max_lvl = 0;
  for (uint32_t i = begin; i < end; ++i) max_lvl = std::max<uint32_t>(max_lvl, h->node_level[i]);
  const uint32_t Tcap = (uint32_t)std::min<uint64_t>((uint64_t)max_batch * (max_lvl + 1), (uint64_t)max_batch + h->up_rows + 16);
  const uint64_t req_cap = std::max<uint64_t>((uint64_t)Tcap * mcap, 1);
  H_CUDA(cudaMalloc(&sc.coff, ((size_t)max_batch + 1) * 4));
  H_CUDA(cudaMalloc(&sc.list_node, (size_t)Tcap * 4));
  H_CUDA(cudaMalloc(&sc.list_level, (size_t)Tcap * 4));
  H_CUDA(cudaMalloc(&sc.cand_d, (size_t)Tcap * ef_c * 4));
  H_CUDA(cudaMalloc(&sc.cand_id, (size_t)Tcap * ef_c * 4));
  H_CUDA(cudaMalloc(&sc.cand_cnt, (size_t)Tcap * 4));
  H_CUDA(cudaMalloc(&sc.req_key, req_cap * 8));
  H_CUDA(cudaMalloc(&sc.req_key2, req_cap * 8));
  H_CUDA(cudaMalloc(&sc.req_src, req_cap * 4));
  H_CUDA(cudaMalloc(&sc.req_d, req_cap * 4));
  H_CUDA(cudaMalloc(&sc.perm, req_cap * 4));
  H_CUDA(cudaMalloc(&sc.perm2, req_cap * 4));
  H_CUDA(cudaMalloc(&sc.heads, req_cap * 4));
  H_CUDA(cudaMalloc(&sc.counters, 64));
  if (extend) {
    uint32_t cap = 32;
    const uint32_t need = std::max(ef_c, mcap + 32) * (1 + std::max(g.s0, g.su));
    while (cap < need) cap <<= 1;
    b.extend = 1;
    b.ext_cap = cap;
    const size_t slices = std::max<size_t>(Tcap, 4);  // K2: one slice per task; the sequential link step uses slice 0
    H_CUDA(cudaMalloc(&sc.ext_keys, slices * cap * 8));
    H_CUDA(cudaMalloc(&sc.ext_d, slices * cap * 4));
    H_CUDA(cudaMalloc(&sc.ext_id, slices * cap * 4));
    b.ext_keys = sc.ext_keys;
    b.ext_d = sc.ext_d;
    b.ext_id = sc.ext_id;
  }
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, sc.req_key, sc.req_key2, sc.perm, sc.perm2, (int)req_cap, 0, 40,
                                  st);
  H_CUDA(cudaMalloc(&sc.cub_tmp, cub_bytes));
  {
    std::vector<uint32_t> iota(req_cap);
    for (uint64_t i = 0; i < req_cap; ++i) iota[i] = (uint32_t)i;
    H_CUDA(cudaMemcpy(sc.perm, iota.data(), req_cap * 4, cudaMemcpyHostToDevice));
  }

  uint32_t inserted = begin;
  if (g.entry == NONE) {  // first vector: fresh self-loops only (hnsw.rs:360-373)
    g.entry = begin;
    g.top_level = h->node_level[begin];
    inserted = begin + 1;
    h->n_live += 1;
  }
  std::vector<uint32_t> coff, lnode, llevel;
  while (inserted < end) {
    // a batch never exceeds 1/16 of what is already linked, so that batch members (which do not see
    // each other) stay a small fraction of every neighbourhood
    uint32_t bs = std::min<uint32_t>(std::min<uint32_t>(max_batch, std::max<uint32_t>(1u, h->n_live / 16)), end - inserted);
    const uint32_t top = g.top_level;
    coff.assign(bs + 1, 0);
    lnode.clear();
    llevel.clear();
    for (uint32_t i = 0; i < bs; ++i) {
      uint32_t id = inserted + i;
      uint32_t nl = std::min<uint32_t>(h->node_level[id], top) + 1;
      coff[i] = (uint32_t)lnode.size();
      for (uint32_t L = 0; L < nl; ++L) {
        lnode.push_back(id);
        llevel.push_back(L);
      }
    }
    coff[bs] = (uint32_t)lnode.size();
    const uint32_t T = (uint32_t)lnode.size();
    if (T > Tcap) return set_error(COZO_GPU_ECUDA, "internal: batch list overflow");
    H_CUDA(cudaMemcpyAsync(sc.coff, coff.data(), ((size_t)bs + 1) * 4, cudaMemcpyHostToDevice, st));
    H_CUDA(cudaMemcpyAsync(sc.list_node, lnode.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st));
    H_CUDA(cudaMemcpyAsync(sc.list_level, llevel.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st));
    H_CUDA(cudaMemsetAsync(sc.counters, 0, 64, st));
    BatchParams p{};
    p.begin = inserted;
    p.count = bs;
    p.top = top;
    p.ef_c = ef_c;
    p.coff = sc.coff;
    p.list_node = sc.list_node;
    p.list_level = sc.list_level;
    p.T = T;
    p.cand_d = sc.cand_d;
    p.cand_id = sc.cand_id;
    p.cand_cnt = sc.cand_cnt;
    p.req_key = sc.req_key;
    p.req_src = sc.req_src;
    p.req_d = sc.req_d;
    p.req_count = sc.counters + 1;
    p.counter = sc.counters + 0;
    p.vis = ws->vis;
    p.nwords = nwords;
    p.vlog = ws->vlog;
    p.logcap = logcap;
    p.ns = ns;
    p.lay = lay;
    uint32_t grid1 = std::min<uint32_t>(max_grid1, (bs + wpc - 1) / wpc);
    k1<<<grid1, wpc * 32, smem1, st>>>(g, b, p);
    H_CUDA(cudaGetLastError());
    k2<<<(T + 3) / 4, 128, smem2, st>>>(g, b, p);
    H_CUDA(cudaGetLastError());
    uint32_t nreq = 0;
    H_CUDA(cudaMemcpyAsync(&nreq, sc.counters + 1, 4, cudaMemcpyDeviceToHost, st));
    H_CUDA(cudaStreamSynchronize(st));
    if (nreq > req_cap) return set_error(COZO_GPU_ECUDA, "internal: in-edge queue overflow");
    if (nreq && extend) {
      // the reference links the selected neighbours one by one, nearest first (hnsw.rs:279-357); a later shrink sees
      // the rows an earlier one left.  Requests sit in the queue in exactly that order (K2 writes a task's requests
      // consecutively), so: one launch per request, in queue order, on this stream.
      for (uint32_t r = 0; r < nreq; ++r)
        k4<<<1, 32, smem4, st>>>(g, b, sc.req_key, sc.perm, sc.req_src, sc.req_d, nreq, sc.perm + r, 1);
      H_CUDA(cudaGetLastError());
    } else if (nreq) {
      size_t tb = cub_bytes;
      cub::DeviceRadixSort::SortPairs(sc.cub_tmp, tb, sc.req_key, sc.req_key2, sc.perm, sc.perm2, (int)nreq, 0, 40, st);
      build_heads_kernel<<<(nreq + 255) / 256, 256, 0, st>>>(sc.req_key2, nreq, sc.heads, sc.counters + 2);
      H_CUDA(cudaGetLastError());
      uint32_t nheads = 0;
      H_CUDA(cudaMemcpyAsync(&nheads, sc.counters + 2, 4, cudaMemcpyDeviceToHost, st));
      H_CUDA(cudaStreamSynchronize(st));
      k4<<<(nheads + 3) / 4, 128, smem4, st>>>(g, b, sc.req_key2, sc.perm2, sc.req_src, sc.req_d, nreq, sc.heads,
                                               nheads);
      H_CUDA(cudaGetLastError());
    }
    // a node above the current top becomes the entry point (hnsw.rs:206-218);
    // the entry is the smallest id on the top layer (hnsw.rs:184-199)
    for (uint32_t i = 0; i < bs; ++i) {
      uint32_t id = inserted + i;
      if (h->node_level[id] > g.top_level) {
        g.top_level = h->node_level[id];
        g.entry = id;
      }
    }
    inserted += bs;
    h->n_live += bs;
  }
  H_CUDA(cudaStreamSynchronize(st));
  h->n_levels = g.top_level + 1;
  return 0;
}

// append `count` vectors as ids [n, n+count): levels from the handle's RNG, rows and capacity
static int append_nodes(cozo_gpu_hnsw* h, const float* vectors, int on_device, uint32_t count) {
  HnswDev& g = h->dev;
  const uint32_t n0 = g.n, n1 = n0 + count;
  if ((uint64_t)n0 + count >= 0x7FFFFFFFull) return set_error(COZO_GPU_EUNSUP, "too many vectors");
  // level law (hnsw.rs:46-52), level_multiplier = 1/ln(m) (relation.rs:1147)
  SplitMix64 rng{h->rng_state};
  const double mult = 1.0 / std::log((double)h->m_max);
  h->node_level.resize(n1, 0);
  h->live.resize(n1, 1);
  std::vector<uint32_t> upper_off(count, NONE);
  std::vector<uint32_t> owners;
  uint64_t up = h->up_rows;
  for (uint32_t i = 0; i < count; ++i) {
    double u = rng.uniform();
    double r = -std::log(u) * mult;
    if (!(r < 15.0)) r = 15.0;
    uint8_t lv = (uint8_t)std::floor(r);
    h->node_level[n0 + i] = lv;
    if (lv) {
      upper_off[i] = (uint32_t)up;
      up += lv;
      for (uint32_t L = 0; L < lv; ++L) owners.push_back(n0 + i);
    }
  }
  h->rng_state = rng.s;
  int rc = hnsw_reserve(h, n1, up);
  if (rc) return rc;
  if (vectors)  // nullptr: the (borrowed) rows are already in place
    H_CUDA(cudaMemcpy2D(h->d_vec + (size_t)n0 * g.ld, (size_t)g.ld * 4, vectors, (size_t)g.dim * 4,
                        (size_t)g.dim * 4, count, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  H_CUDA(cudaMemcpy(h->d_upper_off + n0, upper_off.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
  H_CUDA(cudaMemcpy(h->d_node_level + n0, h->node_level.data() + n0, count, cudaMemcpyHostToDevice));
  if (!owners.empty())
    H_CUDA(cudaMemcpy(h->d_up_owner + h->up_rows, owners.data(), owners.size() * 4, cudaMemcpyHostToDevice));
  h->up_rows = up;
  g.n = n1;
  return 0;
}

}  // namespace cozo

using namespace cozo;

extern "C" int cozo_gpu_hnsw_build(cozo_gpu_hnsw_t** out, const CozoGpuHnswBuildDesc* d) {
  if (!out || !d) return set_error(COZO_GPU_EINVAL, "null argument");
  *out = nullptr;
  int rc = ensure_init();
  if (rc) return rc;
  if (d->dim == 0 || d->n_vectors == 0 || !d->vectors) return set_error(COZO_GPU_EINVAL, "bad build descriptor");
  if (d->metric < 0 || d->metric > 2) return set_error(COZO_GPU_EINVAL, "unknown distance %d", d->metric);
  if (d->m_neighbours < 2 || d->m_neighbours > 64)
    return set_error(COZO_GPU_EUNSUP, "m_neighbours must be in [2,64] for the device builder");
  if (d->ef_construction == 0) return set_error(COZO_GPU_EINVAL, "ef_construction must be set");  // sys.rs:603
  if (d->n_vectors >= 0x7FFFFFFFu) return set_error(COZO_GPU_EUNSUP, "too many vectors");
  const uint32_t n = d->n_vectors;
  const uint32_t m = d->m_neighbours;

  auto* h = new cozo_gpu_hnsw();
  HnswDev& g = h->dev;
  g.n = 0;
  g.dim = d->dim;
  g.ld = round_up(d->dim, 4);
  g.metric = d->metric;
  g.entry = NONE;
  g.top_level = 0;
  h->m_max = m;       // relation.rs:1145
  h->m_max0 = 2 * m;  // relation.rs:1146
  g.s0 = round_up(h->m_max0, 32);
  g.su = round_up(h->m_max, 32);
  h->ef_construction = d->ef_construction;
  h->keep_pruned = d->keep_pruned_connections;
  h->extend_candidates = d->extend_candidates != 0;
  h->rng_state = d->level_seed;
  h->build_state_ready = true;
  auto fail = [&](int code) {
    cozo_gpu_hnsw_free(h);
    return code;
  };
  const bool borrow = d->vectors_on_device && d->borrow_vectors && g.ld == g.dim;
  if (borrow) {  // search straight out of the caller's buffer; a later insert takes a private copy
    h->d_vec = const_cast<float*>(d->vectors);
    h->vec_owned = false;
    h->borrowed_rows = n;
  }
  rc = append_nodes(h, borrow ? nullptr : d->vectors, d->vectors_on_device, n);
  if (rc) return fail(rc);
  rc = hnsw_insert_range(h, 0, n, d->max_batch);
  if (rc) return fail(rc);
  *out = h;
  return 0;
}

// hnsw_put for rows that sort after every indexed key (query/stored.rs:332 -> hnsw.rs:679-727).
extern "C" int cozo_gpu_hnsw_insert(cozo_gpu_hnsw_t* h, const float* vectors, uint32_t count, int32_t vectors_on_device,
                                    uint32_t ef_construction, int32_t keep_pruned_connections, uint32_t* first_id) {
  if (!h || (count && !vectors)) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (first_id) *first_id = h->dev.n;
  if (count == 0) return 0;
  // A handle staged from the canary row alone (n = 0) grows from empty arrays: hnsw_ensure_build_state gives it
  // capacities equal to its live rows (0), so every row the first insert creates is initialised by grow().
  if (h->m_max < 2 || h->m_max > 64) return set_error(COZO_GPU_EUNSUP, "m_neighbours must be in [2,64] for the device builder");
  if (ef_construction) h->ef_construction = ef_construction;
  if (h->ef_construction == 0) return set_error(COZO_GPU_EINVAL, "ef_construction must be set");
  if (keep_pruned_connections >= 0) h->keep_pruned = keep_pruned_connections;
  // a staged handle learns the manifest's extend_candidates through the option (a built handle keeps its own)
  if (get_option("hnsw.extend_candidates", -1) >= 0) h->extend_candidates = get_option("hnsw.extend_candidates", 0) != 0;
  rc = hnsw_ensure_build_state(h);
  if (rc) return rc;
  const uint32_t n0 = h->dev.n;
  rc = append_nodes(h, vectors, vectors_on_device, count);
  if (rc) return rc;
  return hnsw_insert_range(h, n0, n0 + count, 0);
}

// hnsw_remove (hnsw.rs:728-868) for a batch of ids.
extern "C" int cozo_gpu_hnsw_remove(cozo_gpu_hnsw_t* h, const uint32_t* ids, uint32_t count) {
  if (!h || (count && !ids)) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (count == 0) return 0;
  HnswDev& g = h->dev;
  for (uint32_t i = 0; i < count; ++i)
    if (ids[i] >= g.n) return set_error(COZO_GPU_EINVAL, "id %u out of range", ids[i]);
  rc = hnsw_ensure_build_state(h);
  if (rc) return rc;
  uint32_t* d_ids = nullptr;
  H_CUDA(cudaMalloc(&d_ids, (size_t)count * 4));
  cudaMemcpy(d_ids, ids, (size_t)count * 4, cudaMemcpyHostToDevice);
  mark_dead_kernel<<<(count + 255) / 256, 256>>>(d_ids, count, h->d_dead);
  remove_compact_kernel<<<(g.n + 3) / 4, 128>>>(h->d_adj0, h->d_adj0_dist, h->d_deg0, g.s0, g.n, nullptr, h->d_dead);
  if (h->up_rows)
    remove_compact_kernel<<<(uint32_t)((h->up_rows + 3) / 4), 128>>>(h->d_adj_up, h->d_adj_up_dist, h->d_deg_up, g.su,
                                                                     (uint32_t)h->up_rows, h->d_up_owner, h->d_dead);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(d_ids);
  if (e != cudaSuccess) return set_error(COZO_GPU_ECUDA, "remove failed: %s", cudaGetErrorString(e));
  for (uint32_t i = 0; i < count; ++i)
    if (h->live[ids[i]]) {
      h->live[ids[i]] = 0;
      h->n_live--;
    }
  // the entry point was removed: re-point to the first remaining row in key order (hnsw.rs:828-865)
  if (g.entry != NONE && !h->live[g.entry]) {
    uint32_t best = NONE;
    uint32_t top = 0;
    for (uint32_t i = 0; i < g.n; ++i)
      if (h->live[i] && (best == NONE || h->node_level[i] > top)) {
        best = i;
        top = h->node_level[i];
      }
    g.entry = best;
    g.top_level = best == NONE ? 0 : top;
    h->n_levels = g.top_level + 1;
  }
  return 0;
}

__global__ void clear_dead_kernel(const uint32_t* ids, uint32_t count, uint8_t* dead) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dead[ids[i]] = 0;
}

// hnsw_put of a CHANGED vector under an existing key (hnsw.rs:175-182: the hash differs, so the old
// vector is removed and the new one inserted).  The node keeps its id and its layer (the reference
// draws a fresh level; the level law is the same either way).  Also revives removed ids.
extern "C" int cozo_gpu_hnsw_update(cozo_gpu_hnsw_t* h, const uint32_t* ids, const float* vectors, uint32_t count,
                                    uint32_t ef_construction, int32_t keep_pruned_connections) {
  if (!h || (count && (!ids || !vectors))) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (count == 0) return 0;
  HnswDev& g = h->dev;
  std::vector<uint32_t> order(count);
  for (uint32_t i = 0; i < count; ++i) {
    if (ids[i] >= g.n) return set_error(COZO_GPU_EINVAL, "id %u out of range", ids[i]);
    order[i] = i;
  }
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ids[a] < ids[b]; });
  for (uint32_t i = 1; i < count; ++i)
    if (ids[order[i]] == ids[order[i - 1]]) return set_error(COZO_GPU_EINVAL, "id %u given twice", ids[order[i]]);
  if (ef_construction) h->ef_construction = ef_construction;
  if (h->ef_construction == 0) return set_error(COZO_GPU_EINVAL, "ef_construction must be set");
  if (keep_pruned_connections >= 0) h->keep_pruned = keep_pruned_connections;
  rc = hnsw_ensure_build_state(h);
  if (rc) return rc;
  // 1. drop the old vectors' links (no-op for ids that were already removed)
  std::vector<uint32_t> live_ids;
  for (uint32_t i = 0; i < count; ++i)
    if (h->live[ids[i]]) live_ids.push_back(ids[i]);
  if (!live_ids.empty()) {
    rc = cozo_gpu_hnsw_remove(h, live_ids.data(), (uint32_t)live_ids.size());
    if (rc) return rc;
  }
  // 2. new payloads, revive
  if (!h->vec_owned) {  // never write into a borrowed buffer
    rc = hnsw_reserve(h, h->borrowed_rows + 1, h->up_rows);
    if (rc) return rc;
  }
  for (uint32_t i = 0; i < count; ++i)
    H_CUDA(cudaMemcpy(h->d_vec + (size_t)ids[i] * g.ld, vectors + (size_t)i * g.dim, (size_t)g.dim * 4,
                      cudaMemcpyHostToDevice));
  uint32_t* d_ids = nullptr;
  H_CUDA(cudaMalloc(&d_ids, (size_t)count * 4));
  cudaMemcpy(d_ids, ids, (size_t)count * 4, cudaMemcpyHostToDevice);
  clear_dead_kernel<<<(count + 255) / 256, 256>>>(d_ids, count, h->d_dead);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(d_ids);
  if (e != cudaSuccess) return set_error(COZO_GPU_ECUDA, "update failed: %s", cudaGetErrorString(e));
  for (uint32_t i = 0; i < count; ++i) h->live[ids[i]] = 1;
  // 3. re-insert, one contiguous id run at a time (hnsw_insert_range links [begin,end) into the graph)
  for (uint32_t i = 0; i < count;) {
    uint32_t j = i + 1;
    while (j < count && ids[order[j]] == ids[order[j - 1]] + 1) ++j;
    rc = hnsw_insert_range(h, ids[order[i]], ids[order[j - 1]] + 1, 0);
    if (rc) return rc;
    i = j;
  }
  return 0;
}
