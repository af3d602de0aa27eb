// graph.cu — FixedRule graph algorithms on an HBM-resident CSR:
//   PageRank              (fixed_rule/algos/pagerank.rs:29-56 -> graph::page_rank)
//   multi-source Dijkstra (fixed_rule/algos/shortest_path_dijkstra.rs:274-339)
//   ClosenessCentrality   (fixed_rule/algos/all_pairs_shortest_path.rs:97-143)
//   BetweennessCentrality (fixed_rule/algos/all_pairs_shortest_path.rs:29-95)
// The CSR is what GraphBuilder::csr_layout(Sorted) produces for the edge stream
// of as_directed_graph / as_directed_weighted_graph (fixed_rule/mod.rs:136-328).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

struct cozo_gpu_graph {
  uint32_t n = 0;
  uint64_t m = 0;
  bool weighted = false;
  uint32_t *out_ptr = nullptr, *out_idx = nullptr, *in_ptr = nullptr, *in_idx = nullptr;
  float* out_w = nullptr;
  // ---- PageRank works on a RELABELLED copy of the in-CSR ("slot space") --------------------------
  // slot[v] = rank of v by out-degree, descending.  A source is gathered once per out-edge, so in
  // slot space the few ten-thousand sources that carry almost half of an R-MAT graph's edges are the
  // first entries of the contribution vector: the pull kernel keeps them in shared memory.  Rows are
  // processed in slot order (scores, contributions and out-degrees are all slot-indexed: coalesced,
  // no scatter), and each row's in-neighbours are sorted by slot, so neighbouring lanes tend to
  // touch the same cache line.  Scores are un-permuted once, when the result is copied out.
  uint32_t *slot = nullptr;                                 // [n]   original id -> slot
  uint32_t *pr_in_ptr = nullptr, *pr_in_idx = nullptr;      // in-CSR in slot space ([n+1], [m])
  uint32_t *pr_od = nullptr;                                // [n]   out-degree by slot
  uint32_t* hubs = nullptr;  // slot-space rows with in-degree > HUB_T
  uint32_t n_hubs = 0;
  uint32_t* blk_start = nullptr;  // [2*n_blk] row mini-blocks [r0,r1) of the pull kernel (<=32 rows, <=256 in-edges)
  uint32_t n_blk = 0;
  uint32_t* med_rows = nullptr;   // rows with BLK_CAP < in-degree <= HUB_T: one warp each
  uint32_t n_med = 0;
  // hub rows are cut into chunks of <= HUB_CHUNK in-edges, one warp per chunk
  uint32_t *hub_chunk_ptr = nullptr, *chunk_beg = nullptr, *chunk_end = nullptr;
  uint32_t n_chunks = 0;
};

namespace cozo {

constexpr uint32_t HUB_T = 4096;   // rows longer than this are cut into HUB_CHUNK pieces, one warp each
constexpr uint32_t BLK_CAP = 256;  // in-edges of a warp's mini-block (8 gathers per lane in flight)
constexpr uint32_t BLK_ROWS = 32;  // rows per mini-block: one lane sums one row
constexpr uint32_t HUB_CHUNK = 4096; // in-edges of a hub row summed by one warp

__global__ void edge_check_kernel(const uint32_t* src, const uint32_t* dst, const float* w, uint64_t m, uint32_t n,
                                  int* bad) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  if (src[e] >= n || dst[e] >= n) *bad = 1;
  if (w) {
    float x = w[e];
    if (!(x >= 0.f) || isinf(x)) *bad = 2;  // finite and non-negative (fixed_rule/mod.rs:258-286)
  }
}

__global__ void make_keys_kernel(const uint32_t* hi, const uint32_t* lo, uint64_t m, unsigned long long* keys,
                                 uint32_t* deg) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  keys[e] = ((unsigned long long)hi[e] << 32) | lo[e];
  atomicAdd(&deg[hi[e]], 1u);
}

__global__ void split_keys_kernel(const unsigned long long* keys, uint64_t m, uint32_t* lo) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  lo[e] = (uint32_t)(keys[e] & 0xFFFFFFFFull);
}

__global__ void gather_w_kernel(const float* w, const uint32_t* perm, uint64_t m, float* out) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  out[e] = w[perm[e]];
}

__global__ void iota_kernel(uint32_t* p, uint64_t m) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m) p[e] = (uint32_t)e;
}

// ---- PageRank ----------------------------------------------------------------
// GAP-style pull iteration (graph 0.3.1 page_rank, un-vendored):
//   new[u] = base + d * sum_{v in in(u)} contrib[v];  err += |new[u]-old[u]| (f64)
//   contrib'[u] = new[u] / out_degree(u)
// Everything below is indexed in slot space (see struct cozo_gpu_graph).
__global__ void pr_init_kernel(const uint32_t* __restrict__ od, uint32_t n, float init, float* scores, float* contrib) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  scores[r] = init;
  const uint32_t d = od[r];
  contrib[r] = d ? init / (float)d : 0.f;  // d==0: value is never read (no out edge leads anywhere)
}
__global__ void pr_slot_kernel(const uint32_t* perm, uint32_t n, uint32_t* slot) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) slot[perm[r]] = r;
}
__global__ void pr_degkey_kernel(const uint32_t* out_ptr, uint32_t n, uint32_t* key, uint32_t* val) {
  uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n) {
    key[u] = 0xFFFFFFFFu - (out_ptr[u + 1] - out_ptr[u]);  // ascending sort => descending out-degree
    val[u] = u;
  }
}
// per original row u: degrees by slot
__global__ void pr_degrees_kernel(const uint32_t* out_ptr, const uint32_t* in_ptr, const uint32_t* slot, uint32_t n,
                                  uint32_t* od_slot, uint32_t* in_cnt_slot) {
  uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const uint32_t r = slot[u];
  od_slot[r] = out_ptr[u + 1] - out_ptr[u];
  in_cnt_slot[r] = in_ptr[u + 1] - in_ptr[u];
}
// one warp per original destination row: key = (slot[dst] << 32) | slot[src] for each in-edge
__global__ void pr_edge_keys_kernel(const uint32_t* __restrict__ in_ptr, const uint32_t* __restrict__ in_idx,
                                    const uint32_t* __restrict__ slot, uint32_t n, unsigned long long* keys) {
  const uint32_t w = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const unsigned long long hi = (unsigned long long)slot[w] << 32;
  for (uint32_t e = in_ptr[w] + lane; e < in_ptr[w + 1]; e += 32) keys[e] = hi | slot[in_idx[e]];
}
__global__ void pr_unpermute_kernel(const float* __restrict__ scores_slot, const uint32_t* __restrict__ slot, uint32_t n,
                                    float* out) {
  uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n) out[u] = scores_slot[slot[u]];
}

__device__ __forceinline__ void block_add_err(double e, double* out) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if (lane == 0) sh[warp] = e;
  __syncthreads();
  if (warp == 0) {
    int nw = (blockDim.x + 31) >> 5;
    double v = lane < nw ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v != 0.0) atomicAdd(out, v);
  }
}

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float ldg_f32_hint(const float* a, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u32_hint(const uint32_t* a, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}

struct PrArgs {
  const uint32_t *in_ptr, *in_idx, *od;            // slot-space in-CSR and out-degrees
  const uint32_t *blk_start, *med_rows, *chunk_beg, *chunk_end;
  uint32_t n_blk, n_med, n_chunks;
  uint32_t dynamic;                                // 1: warps draw work items from device counters
  float base, damping;
  const float* contrib_old;
  float *contrib_new, *scores, *partial;
  double* err;                                     // err[0]; the three work counters follow at err + 1
};

// a warp sums contrib over in_idx[b, en): 8 independent gathers per lane in flight, shuffle tree
__device__ __forceinline__ float pr_warp_row_sum(const PrArgs& a, uint32_t b, uint32_t en, int lane, uint64_t keep,
                                                 uint64_t stream) {
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t k = b + lane;
  for (; k + 224 < en; k += 256) {
    uint32_t i[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) i[j] = ldg_u32_hint(a.in_idx + k + 32 * j, stream);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += ldg_f32_hint(a.contrib_old + i[j], keep);
  }
  for (; k < en; k += 32) s[0] += ldg_f32_hint(a.contrib_old + ldg_u32_hint(a.in_idx + k, stream), keep);
  return warp_sum(((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
}

// work distribution of the pull kernel: static round-robin over warps, or `batch` consecutive items
// at a time from a device counter (one atomic per batch, lane 0, broadcast)
struct PrCursor {
  uint32_t cur, end, stride;
};
__device__ __forceinline__ bool pr_next(PrCursor& c, uint32_t n_items, uint32_t* ctr, uint32_t batch, bool dynamic,
                                        int lane) {
  if (!dynamic) {
    if (c.cur >= n_items) return false;
    c.end = c.cur + 1;  // one item; the caller advances by stride
    return true;
  }
  uint32_t v = 0;
  if (lane == 0) v = atomicAdd(ctr, batch);
  v = __shfl_sync(0xffffffffu, v, 0);
  if (v >= n_items) return false;
  c.cur = v;
  c.end = min(v + batch, n_items);
  return true;
}

// The pull iteration: ONE persistent launch.
//   phase 1   hub chunks (<= 4096 in-edges of a row with in-degree > HUB_T): one warp per chunk writes
//             a partial sum; pr_hub_final_kernel adds a row's partials in chunk order afterwards
//   phase 2   medium rows (BLK_CAP < in-degree <= HUB_T): one warp per row
//   phase 3   CSR-stream at warp granularity: a warp owns a mini-block of <= 32 consecutive rows with
//             <= 256 in-edges in total; the in_idx slice is read coalesced (L2 evict-first), every lane
//             issues up to 8 independent gathers (L2 evict-last: the 4N-byte vector is asked to stay in
//             L2) before the first use and stages the values in the warp's 1 KB of shared memory; then
//             lane l sums row l's slice in stored (slot) order.
// Heavy items first, so the tail of the launch is made of the smallest work items.  No block-wide
// barrier before the final error reduction: independent warp pipelines.
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) pr_pull_kernel(const PrArgs a) {
  __shared__ float vals_all[WARPS][BLK_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* vals = vals_all[warp];
  const uint64_t keep = l2_policy_evict_last();
  const uint64_t stream = l2_policy_evict_first();
  const uint32_t wglobal = blockIdx.x * WARPS + warp, wtotal = gridDim.x * WARPS;
  const bool dyn = a.dynamic != 0;
  uint32_t* ctr = reinterpret_cast<uint32_t*>(a.err + 1);
  double e = 0.0;
  PrCursor c{wglobal, 0, wtotal};
  while (pr_next(c, a.n_chunks, ctr + 0, 1, dyn, lane)) {
    for (uint32_t i = c.cur; i < c.end; ++i) {
      const float s = pr_warp_row_sum(a, a.chunk_beg[i], a.chunk_end[i], lane, keep, stream);
      if (lane == 0) a.partial[i] = s;
    }
    c.cur += c.stride;
  }
  c = PrCursor{wglobal, 0, wtotal};
  while (pr_next(c, a.n_med, ctr + 1, 2, dyn, lane)) {
    for (uint32_t i = c.cur; i < c.end; ++i) {
      const uint32_t r = a.med_rows[i];
      const float s = pr_warp_row_sum(a, a.in_ptr[r], a.in_ptr[r + 1], lane, keep, stream);
      if (lane == 0) {
        const float nw = a.base + a.damping * s;
        e += (double)fabsf(nw - a.scores[r]);
        a.scores[r] = nw;
        const uint32_t od = a.od[r];
        a.contrib_new[r] = od ? nw / (float)od : 0.f;
      }
    }
    c.cur += c.stride;
  }
  c = PrCursor{wglobal, 0, wtotal};
  while (pr_next(c, a.n_blk, ctr + 2, 8, dyn, lane)) {
    for (uint32_t blk = c.cur; blk < c.end; ++blk) {
      const uint32_t r0 = a.blk_start[2 * blk], r1 = a.blk_start[2 * blk + 1];
      const uint32_t nrows = r1 - r0;
      uint32_t lo = 0, hi = 0;
      if ((uint32_t)lane < nrows) {
        lo = a.in_ptr[r0 + lane];
        hi = a.in_ptr[r0 + lane + 1];
      }
      const uint32_t e0 = __shfl_sync(0xffffffffu, lo, 0);
      const uint32_t e1 = __shfl_sync(0xffffffffu, hi, nrows - 1);
      uint32_t idx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t k = e0 + lane + 32 * j;
        idx[j] = k < e1 ? ldg_u32_hint(a.in_idx + k, stream) : NONE;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) vals[lane + 32 * j] = idx[j] != NONE ? ldg_f32_hint(a.contrib_old + idx[j], keep) : 0.f;
      __syncwarp();
      if ((uint32_t)lane < nrows) {
        const uint32_t r = r0 + lane;
        float s = 0.f;
        for (uint32_t j = lo - e0; j < hi - e0; ++j) s += vals[j];
        const float nw = a.base + a.damping * s;
        e += (double)fabsf(nw - a.scores[r]);
        a.scores[r] = nw;
        const uint32_t od = a.od[r];
        a.contrib_new[r] = od ? nw / (float)od : 0.f;
      }
      __syncwarp();
    }
    c.cur += c.stride;
  }
  block_add_err(e, a.err);
}

// one thread per hub row adds its chunk partials in chunk order (deterministic)
__global__ void __launch_bounds__(256) pr_hub_final_kernel(const uint32_t* __restrict__ hubs, uint32_t n_hubs,
                                                           const uint32_t* __restrict__ hub_chunk_ptr,
                                                           const float* __restrict__ partial,
                                                           const uint32_t* __restrict__ od, float base, float damping,
                                                           float* __restrict__ contrib_new,
                                                           float* __restrict__ scores, double* err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (i < n_hubs) {
    const uint32_t u = hubs[i];
    float s = 0.f;
    for (uint32_t c = hub_chunk_ptr[i]; c < hub_chunk_ptr[i + 1]; ++c) s += partial[c];
    const float nw = base + damping * s;
    e = (double)fabsf(nw - scores[u]);
    scores[u] = nw;
    const uint32_t d = od[u];
    contrib_new[u] = d ? nw / (float)d : 0.f;
  }
  block_add_err(e, err);
}

// ---- SSSP ---------------------------------------------------------------------
// One CTA per source.  Label-correcting relaxation to the fixed point
// dist[v] = min_u fl32(dist[u] + w(u,v)); with non-negative weights this is the
// value Dijkstra's `cost + path_weight` recursion produces
// (shortest_path_dijkstra.rs:304-309), bit for bit.  State per (source, node)
// is one u64 = (f32 bits of dist << 32) | predecessor, updated by CAS only on a
// strictly smaller distance (strict `<`, shortest_path_dijkstra.rs:305), so the
// predecessors always form a tree.
constexpr unsigned long long SSSP_INF = 0x7F800000FFFFFFFFull;  // (+inf, NONE)

// ForbiddenNode / ForbiddenEdge sets (shortest_path_dijkstra.rs:188-218) of source `si` are the
// slices [fn_ptr[si], fn_ptr[si+1]) of fn_nodes and [fe_ptr[si], fe_ptr[si+1]) of (fe_src, fe_dst);
// they are a handful of entries (KShortestPathYen: one root path + at most k edges).
struct ForbiddenSets {
  const uint32_t *fn_ptr, *fn_nodes, *fe_ptr, *fe_src, *fe_dst;
};

// SMEM = the per-source state (8 B/node) and the two frontier flag arrays (1 B/node each) live in
// shared memory (graphs up to ~22 k nodes); the final state is copied out for the read-out kernels.
// A warp takes one frontier node at a time and its lanes stride the node's out-edges.
template <bool FORB, bool SMEM>
__global__ void __launch_bounds__(256) sssp_kernel(const uint32_t* __restrict__ out_ptr,
                                                   const uint32_t* __restrict__ out_idx,
                                                   const float* __restrict__ out_w, uint32_t n,
                                                   const uint32_t* __restrict__ sources, uint32_t n_src,
                                                   unsigned long long* state, uint8_t* flags, ForbiddenSets fs) {
  extern __shared__ __align__(16) uint8_t sssp_smem[];
  const uint32_t si = blockIdx.x;
  if (si >= n_src) return;
  uint32_t fnb = 0, fne = 0, feb = 0, fee = 0;
  if (FORB) {
    fnb = fs.fn_ptr[si];
    fne = fs.fn_ptr[si + 1];
    feb = fs.fe_ptr[si];
    fee = fs.fe_ptr[si + 1];
  }
  unsigned long long* gst = state + (size_t)si * n;
  unsigned long long* st = SMEM ? reinterpret_cast<unsigned long long*>(sssp_smem) : gst;
  uint8_t* cur = SMEM ? sssp_smem + (size_t)n * 8 : flags + (size_t)si * 2 * n;
  uint8_t* nxt = cur + n;
  __shared__ int s_any;
  const int lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (uint32_t v = threadIdx.x; v < n; v += blockDim.x) {
    st[v] = SSSP_INF;
    cur[v] = 0;
    nxt[v] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = sources[si];
    st[s] = 0x00000000FFFFFFFFull;  // dist 0, no predecessor
    cur[s] = 1;
  }
  __syncthreads();
  for (;;) {
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    bool any_local = false;
    for (uint32_t u = warp; u < n; u += nwarps) {
      if (!cur[u]) continue;  // warp-uniform
      __syncwarp();
      if (lane == 0) cur[u] = 0;
      const float du = __uint_as_float((uint32_t)(st[u] >> 32));
      const uint32_t kb = out_ptr[u], ke = out_ptr[u + 1];
      for (uint32_t k = kb + lane; k < ke; k += 32) {
        const uint32_t v = out_idx[k];
        if (FORB) {  // shortest_path_dijkstra.rs:298-303
          bool skip = false;
          for (uint32_t f = fnb; f < fne; ++f) skip |= fs.fn_nodes[f] == v;
          for (uint32_t f = feb; f < fee; ++f) skip |= (fs.fe_src[f] == u) & (fs.fe_dst[f] == v);
          if (skip) continue;
        }
        const float nd = du + (out_w ? out_w[k] : 1.0f);
        unsigned long long old = st[v];
        while (nd < __uint_as_float((uint32_t)(old >> 32))) {
          unsigned long long want = ((unsigned long long)__float_as_uint(nd) << 32) | u;
          unsigned long long got = atomicCAS(&st[v], old, want);
          if (got == old) {
            nxt[v] = 1;
            any_local = true;
            break;
          }
          old = got;
        }
      }
    }
    if (any_local) s_any = 1;
    __syncthreads();
    const int any = s_any;
    __syncthreads();
    if (!any) break;
    uint8_t* t = cur;
    cur = nxt;
    nxt = t;
  }
  if (SMEM)
    for (uint32_t v = threadIdx.x; v < n; v += blockDim.x) gst[v] = st[v];
}

// launch helper: shared-memory variant when the state fits
template <bool FORB>
static cudaError_t launch_sssp(cozo_gpu_graph_t* g, const uint32_t* d_sources, uint32_t n_src,
                               unsigned long long* state, uint8_t* flags, ForbiddenSets fs) {
  const uint32_t n = g->n;
  const size_t need = (size_t)n * 10;
  if (need + 1024 <= device_info().smem_optin) {
    cudaError_t e = cudaFuncSetAttribute(sssp_kernel<FORB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
    if (e != cudaSuccess) return e;
    sssp_kernel<FORB, true><<<n_src, 256, need>>>(g->out_ptr, g->out_idx, g->out_w, n, d_sources, n_src, state, flags, fs);
  } else {
    sssp_kernel<FORB, false><<<n_src, 256>>>(g->out_ptr, g->out_idx, g->out_w, n, d_sources, n_src, state, flags, fs);
  }
  return cudaGetLastError();
}

// goal-directed read-out (shortest_path_dijkstra.rs:318-336): walk the predecessors from the goal
__global__ void sssp_path_kernel(const unsigned long long* state, uint32_t n, const uint32_t* sources,
                                 const uint32_t* goals, uint32_t n_src, uint32_t max_len, float* cost, uint32_t* len,
                                 uint32_t* paths) {
  const uint32_t si = blockIdx.x * blockDim.x + threadIdx.x;
  if (si >= n_src) return;
  const unsigned long long* st = state + (size_t)si * n;
  const uint32_t s = sources[si], t = goals[si];
  const float c = __uint_as_float((uint32_t)(st[t] >> 32));
  cost[si] = c;
  if (!isfinite(c)) {  // (target, inf, [])
    len[si] = 0;
    return;
  }
  uint32_t cnt = 1, cur = t;
  while (cur != s && cnt <= n) {
    cur = (uint32_t)(st[cur] & 0xFFFFFFFFull);
    ++cnt;
  }
  len[si] = cnt;
  if (cnt > max_len) return;  // caller sees len > max_len and retries with a larger buffer
  uint32_t* p = paths + (size_t)si * max_len;
  cur = t;
  for (uint32_t i = cnt; i-- > 0;) {
    p[i] = cur;
    cur = (uint32_t)(st[cur] & 0xFFFFFFFFull);
  }
}

__global__ void sssp_unpack_kernel(const unsigned long long* state, uint64_t total, float* dist, uint32_t* pred) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  unsigned long long s = state[i];
  if (dist) dist[i] = __uint_as_float((uint32_t)(s >> 32));
  if (pred) pred[i] = (uint32_t)(s & 0xFFFFFFFFull);
}

// closeness of one source: nc^2 / total / (n-1), nc counting the source itself
// (all_pairs_shortest_path.rs:118-120)
__global__ void __launch_bounds__(256) closeness_kernel(const unsigned long long* state, uint32_t n, uint32_t n_src,
                                                        uint32_t src_base, float* out) {
  const uint32_t si = blockIdx.x;
  if (si >= n_src) return;
  const unsigned long long* st = state + (size_t)si * n;
  double tot = 0.0;
  uint32_t cnt = 0;
  for (uint32_t v = threadIdx.x; v < n; v += blockDim.x) {
    float d = __uint_as_float((uint32_t)(st[v] >> 32));
    if (isfinite(d)) {
      tot += (double)d;
      cnt++;
    }
  }
  __shared__ double sh_t[8];
  __shared__ uint32_t sh_c[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    sh_t[warp] = tot;
    sh_c[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    uint32_t c = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      t += sh_t[i];
      c += sh_c[i];
    }
    float nc = (float)c;
    float total = (float)t;
    out[src_base + si] = nc * nc / total / (float)(n - 1);
  }
}

// Betweenness of one source (Brandes form of all_pairs_shortest_path.rs:54-68:
// every tied shortest path adds 1/l to each interior node, i.e. node v receives
// sigma_st(v)/sigma_st per target t).  sigma and delta are iterated to their
// fixed points over the shortest-path DAG { (u,v) : fl32(dist[u]+w) == dist[v] }.
__global__ void __launch_bounds__(256) betweenness_kernel(const uint32_t* __restrict__ out_ptr,
                                                          const uint32_t* __restrict__ out_idx,
                                                          const float* __restrict__ out_w, uint32_t n,
                                                          const uint32_t* __restrict__ sources, uint32_t n_src,
                                                          const unsigned long long* state, double* sigma_buf,
                                                          double* delta_buf, double* bc) {
  const uint32_t si = blockIdx.x;
  if (si >= n_src) return;
  const unsigned long long* st = state + (size_t)si * n;
  double* sigma = sigma_buf + (size_t)si * 2 * n;
  double* sigma2 = sigma + n;
  double* delta = delta_buf + (size_t)si * 2 * n;
  double* delta2 = delta + n;
  const uint32_t s = sources[si];
  __shared__ int s_any;
  for (uint32_t v = threadIdx.x; v < n; v += blockDim.x) {
    sigma[v] = v == s ? 1.0 : 0.0;
    sigma2[v] = 0.0;
    delta[v] = 0.0;
    delta2[v] = 0.0;
  }
  __syncthreads();
  // sigma: push along DAG edges until nothing changes (<= DAG depth rounds)
  for (uint32_t round = 0; round < n + 1; ++round) {
    if (threadIdx.x == 0) s_any = 0;
    for (uint32_t v = threadIdx.x; v < n; v += blockDim.x) sigma2[v] = v == s ? 1.0 : 0.0;
    __syncthreads();
    for (uint32_t u = threadIdx.x; u < n; u += blockDim.x) {
      const float du = __uint_as_float((uint32_t)(st[u] >> 32));
      if (!isfinite(du) || sigma[u] == 0.0) continue;
      for (uint32_t k = out_ptr[u]; k < out_ptr[u + 1]; ++k) {
        const uint32_t v = out_idx[k];
        if (v == s) continue;
        const float dv = __uint_as_float((uint32_t)(st[v] >> 32));
        if (du + (out_w ? out_w[k] : 1.0f) == dv) atomicAdd(&sigma2[v], sigma[u]);
      }
    }
    __syncthreads();
    for (uint32_t v = threadIdx.x; v < n; v += blockDim.x)
      if (sigma2[v] != sigma[v]) s_any = 1;
    __syncthreads();
    const int any = s_any;
    double* t = sigma;
    sigma = sigma2;
    sigma2 = t;
    __syncthreads();
    if (!any) break;
  }
  // delta: pull from DAG successors until nothing changes
  for (uint32_t round = 0; round < n + 1; ++round) {
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    for (uint32_t u = threadIdx.x; u < n; u += blockDim.x) {
      const float du = __uint_as_float((uint32_t)(st[u] >> 32));
      double acc = 0.0;
      if (isfinite(du) && sigma[u] != 0.0) {
        for (uint32_t k = out_ptr[u]; k < out_ptr[u + 1]; ++k) {
          const uint32_t v = out_idx[k];
          if (v == s) continue;
          const float dv = __uint_as_float((uint32_t)(st[v] >> 32));
          if (du + (out_w ? out_w[k] : 1.0f) == dv) acc += sigma[u] / sigma[v] * (1.0 + delta[v]);
        }
      }
      delta2[u] = acc;
      if (acc != delta[u]) s_any = 1;
    }
    __syncthreads();
    const int any = s_any;
    double* t = delta;
    delta = delta2;
    delta2 = t;
    __syncthreads();
    if (!any) break;
  }
  for (uint32_t v = threadIdx.x; v < n; v += blockDim.x)
    if (v != s && delta[v] != 0.0) atomicAdd(&bc[v], delta[v]);
}

// ClusteringCoefficients (fixed_rule/algos/triangles.rs:59-98): one warp per node u.  For every
// position i of u's (sorted, duplicate-keeping) out-neighbour list the lanes sweep the positions j
// whose value is smaller and test membership of that value in adj(edges[i]) by binary search.
// Integer work: counts are exact, cc is the same f64 expression as the reference.
__global__ void __launch_bounds__(256) clustering_kernel(const uint32_t* __restrict__ out_ptr,
                                                         const uint32_t* __restrict__ out_idx, uint32_t n,
                                                         double* __restrict__ cc,
                                                         unsigned long long* __restrict__ n_tri,
                                                         unsigned long long* __restrict__ degree) {
  const int lane = threadIdx.x & 31;
  const uint32_t u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= n) return;
  const uint32_t b = out_ptr[u], e = out_ptr[u + 1];
  const uint32_t deg = e - b;
  unsigned long long t = 0;
  if (deg >= 2) {
    uint32_t lo_end = b;  // first position whose value is >= a (the list is sorted, a is non-decreasing in i)
    for (uint32_t i = b; i < e; ++i) {
      const uint32_t a = out_idx[i];
      while (lo_end < e && out_idx[lo_end] < a) ++lo_end;
      const uint32_t ab = out_ptr[a], ae = out_ptr[a + 1];
      for (uint32_t j = b + lane; j < lo_end; j += 32) {
        const uint32_t v = out_idx[j];
        uint32_t lo = ab, hi = ae;
        while (lo < hi) {
          uint32_t mid = (lo + hi) >> 1;
          if (out_idx[mid] < v) lo = mid + 1;
          else hi = mid;
        }
        t += (lo < ae && out_idx[lo] == v) ? 1ull : 0ull;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  if (lane == 0) {
    degree[u] = deg;
    n_tri[u] = t;
    cc[u] = deg < 2 ? 0.0 : 2. * (double)t / ((double)deg * ((double)deg - 1.));
  }
}

__global__ void f64_to_f32_kernel(const double* in, uint32_t n, float* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

static bool poisoned(const volatile int* p) { return p && *p != 0; }

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

}  // namespace cozo

using namespace cozo;

extern "C" void cozo_gpu_graph_free(cozo_gpu_graph_t* g) {
  if (!g) return;
  void* ptrs[] = {g->out_ptr,  g->out_idx,       g->in_ptr,    g->in_idx,   g->out_w,
                  g->hubs,     g->blk_start,     g->hub_chunk_ptr, g->chunk_beg, g->chunk_end,
                  g->med_rows, g->slot,         g->pr_in_ptr, g->pr_in_idx, g->pr_od};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete g;
}

extern "C" int cozo_gpu_graph_stage(cozo_gpu_graph_t** out, uint32_t n, uint64_t m, const uint32_t* src,
                                    const uint32_t* dst, const float* w) {
  if (!out) return set_error(COZO_GPU_EINVAL, "null argument");
  *out = nullptr;
  int rc = ensure_init();
  if (rc) return rc;
  if (m && (!src || !dst)) return set_error(COZO_GPU_EINVAL, "null edge arrays");
  if (m >= 0xFFFFFFFFull) return set_error(COZO_GPU_EUNSUP, "more than 2^32-2 edges");
  auto* g = new cozo_gpu_graph();
  g->n = n;
  g->m = m;
  g->weighted = w != nullptr;
  auto fail = [&](int code) {
    cozo_gpu_graph_free(g);
    return code;
  };
#define G_CUDA(call)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(set_error(_e == cudaErrorMemoryAllocation ? COZO_GPU_ENOMEM : COZO_GPU_ECUDA,    \
                            "%s failed: %s (line %d)", #call, cudaGetErrorString(_e), __LINE__)); \
  } while (0)
  const size_t np1 = (size_t)n + 1;
  const size_t mm = std::max<uint64_t>(m, 1);
  G_CUDA(cudaMalloc(&g->out_ptr, np1 * 4));
  G_CUDA(cudaMalloc(&g->in_ptr, np1 * 4));
  G_CUDA(cudaMalloc(&g->out_idx, mm * 4));
  G_CUDA(cudaMalloc(&g->in_idx, mm * 4));
  if (w) G_CUDA(cudaMalloc(&g->out_w, mm * 4));
  G_CUDA(cudaMemset(g->out_ptr, 0, np1 * 4));
  G_CUDA(cudaMemset(g->in_ptr, 0, np1 * 4));
  if (m) {
    DevBuf dsrc, ddst, dw, keys, keys2, perm, perm2, tmp, bad;
    G_CUDA(cudaMalloc(&dsrc.p, m * 4));
    G_CUDA(cudaMalloc(&ddst.p, m * 4));
    G_CUDA(cudaMalloc(&keys.p, m * 8));
    G_CUDA(cudaMalloc(&keys2.p, m * 8));
    G_CUDA(cudaMalloc(&bad.p, 4));
    G_CUDA(cudaMemset(bad.p, 0, 4));
    G_CUDA(cudaMemcpy(dsrc.p, src, m * 4, cudaMemcpyHostToDevice));
    G_CUDA(cudaMemcpy(ddst.p, dst, m * 4, cudaMemcpyHostToDevice));
    if (w) {
      G_CUDA(cudaMalloc(&dw.p, m * 4));
      G_CUDA(cudaMemcpy(dw.p, w, m * 4, cudaMemcpyHostToDevice));
    }
    const uint32_t tb = 256;
    const uint32_t gb = (uint32_t)((m + tb - 1) / tb);
    edge_check_kernel<<<gb, tb>>>(dsrc.as<uint32_t>(), ddst.as<uint32_t>(), dw.as<float>(), m, n, bad.as<int>());
    int hbad = 0;
    G_CUDA(cudaMemcpy(&hbad, bad.p, 4, cudaMemcpyDeviceToHost));
    if (hbad == 1) return fail(set_error(COZO_GPU_EINVAL, "edge endpoint out of range (n=%u)", n));
    if (hbad == 2) return fail(set_error(COZO_GPU_EINVAL, "edge weight must be finite and non-negative"));
    int end_bit = 32;
    while (end_bit < 64 && (n >> (end_bit - 32)) != 0) ++end_bit;  // bits of the row id
    size_t tmp_bytes = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.as<unsigned long long>(), keys2.as<unsigned long long>(),
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int)m, 0, end_bit);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, g->out_ptr, g->out_ptr, (int)np1);
    tmp_bytes = std::max(tmp_bytes, tb2);
    G_CUDA(cudaMalloc(&tmp.p, tmp_bytes));
    // out-CSR: sort by (src, dst); equal pairs keep input order (stable LSD radix sort)
    make_keys_kernel<<<gb, tb>>>(dsrc.as<uint32_t>(), ddst.as<uint32_t>(), m, keys.as<unsigned long long>(),
                                 g->out_ptr);
    if (w) {
      G_CUDA(cudaMalloc(&perm.p, m * 4));
      G_CUDA(cudaMalloc(&perm2.p, m * 4));
      iota_kernel<<<gb, tb>>>(perm.as<uint32_t>(), m);
      size_t t3 = tmp_bytes;
      cub::DeviceRadixSort::SortPairs(tmp.p, t3, keys.as<unsigned long long>(), keys2.as<unsigned long long>(),
                                      perm.as<uint32_t>(), perm2.as<uint32_t>(), (int)m, 0, end_bit);
      gather_w_kernel<<<gb, tb>>>(dw.as<float>(), perm2.as<uint32_t>(), m, g->out_w);
    } else {
      size_t t3 = tmp_bytes;
      cub::DeviceRadixSort::SortKeys(tmp.p, t3, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), (int)m,
                                     0, end_bit);
    }
    split_keys_kernel<<<gb, tb>>>(keys2.as<unsigned long long>(), m, g->out_idx);
    {
      size_t t3 = tmp_bytes;
      cub::DeviceScan::ExclusiveSum(tmp.p, t3, g->out_ptr, g->out_ptr, (int)np1);
    }
    // in-CSR: sort by (dst, src)
    make_keys_kernel<<<gb, tb>>>(ddst.as<uint32_t>(), dsrc.as<uint32_t>(), m, keys.as<unsigned long long>(),
                                 g->in_ptr);
    {
      size_t t3 = tmp_bytes;
      cub::DeviceRadixSort::SortKeys(tmp.p, t3, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), (int)m,
                                     0, end_bit);
    }
    split_keys_kernel<<<gb, tb>>>(keys2.as<unsigned long long>(), m, g->in_idx);
    {
      size_t t3 = tmp_bytes;
      cub::DeviceScan::ExclusiveSum(tmp.p, t3, g->in_ptr, g->in_ptr, (int)np1);
    }
    G_CUDA(cudaGetLastError());
    G_CUDA(cudaDeviceSynchronize());
  }
  if (n) {
    // slot space for PageRank: slot = rank by out-degree (descending); in-CSR relabelled and re-sorted
    {
      DevBuf key, key2, val, perm, tmp2;
      G_CUDA(cudaMalloc(&key.p, (size_t)n * 4));
      G_CUDA(cudaMalloc(&key2.p, (size_t)n * 4));
      G_CUDA(cudaMalloc(&val.p, (size_t)n * 4));
      G_CUDA(cudaMalloc(&perm.p, (size_t)n * 4));
      G_CUDA(cudaMalloc(&g->slot, (size_t)n * 4));
      G_CUDA(cudaMalloc(&g->pr_od, (size_t)n * 4));
      G_CUDA(cudaMalloc(&g->pr_in_ptr, np1 * 4));
      G_CUDA(cudaMalloc(&g->pr_in_idx, mm * 4));
      pr_degkey_kernel<<<(n + 255) / 256, 256>>>(g->out_ptr, n, key.as<uint32_t>(), val.as<uint32_t>());
      size_t sb = 0, sb2 = 0, sb3 = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, sb, key.as<uint32_t>(), key2.as<uint32_t>(), val.as<uint32_t>(),
                                      perm.as<uint32_t>(), (int)n);
      cub::DeviceScan::ExclusiveSum(nullptr, sb2, g->pr_in_ptr, g->pr_in_ptr, (int)np1);
      int end_bit = 32;
      while (end_bit < 64 && (n >> (end_bit - 32)) != 0) ++end_bit;
      if (m)
        cub::DeviceRadixSort::SortKeys(nullptr, sb3, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)m, 0,
                                       end_bit);
      sb = std::max(sb, std::max(sb2, sb3));
      G_CUDA(cudaMalloc(&tmp2.p, sb));
      {
        size_t t = sb;
        cub::DeviceRadixSort::SortPairs(tmp2.p, t, key.as<uint32_t>(), key2.as<uint32_t>(), val.as<uint32_t>(),
                                        perm.as<uint32_t>(), (int)n);
      }
      pr_slot_kernel<<<(n + 255) / 256, 256>>>(perm.as<uint32_t>(), n, g->slot);
      G_CUDA(cudaMemset(g->pr_in_ptr, 0, np1 * 4));
      pr_degrees_kernel<<<(n + 255) / 256, 256>>>(g->out_ptr, g->in_ptr, g->slot, n, g->pr_od, g->pr_in_ptr);
      {
        size_t t = sb;
        cub::DeviceScan::ExclusiveSum(tmp2.p, t, g->pr_in_ptr, g->pr_in_ptr, (int)np1);
      }
      if (m) {
        DevBuf k1, k2;
        G_CUDA(cudaMalloc(&k1.p, m * 8));
        G_CUDA(cudaMalloc(&k2.p, m * 8));
        pr_edge_keys_kernel<<<(uint32_t)(((uint64_t)n * 32 + 255) / 256), 256>>>(g->in_ptr, g->in_idx, g->slot, n,
                                                                                 k1.as<unsigned long long>());
        size_t t = sb;
        cub::DeviceRadixSort::SortKeys(tmp2.p, t, k1.as<unsigned long long>(), k2.as<unsigned long long>(), (int)m, 0,
                                       end_bit);
        split_keys_kernel<<<(uint32_t)((m + 255) / 256), 256>>>(k2.as<unsigned long long>(), m, g->pr_in_idx);
      }
      G_CUDA(cudaGetLastError());
      G_CUDA(cudaDeviceSynchronize());
    }
    // row blocks for the pull kernel: consecutive rows, <= BLK_CAP in-edges and <= BLK_ROWS rows,
    // hub rows (in-degree > HUB_T) excluded and listed separately
    std::vector<uint32_t> hin(np1);
    G_CUDA(cudaMemcpy(hin.data(), g->pr_in_ptr, np1 * 4, cudaMemcpyDeviceToHost));  // slot-space rows
    std::vector<uint32_t> blocks, hubs, med;
    uint32_t r = 0;
    while (r < n) {
      const uint32_t d0 = hin[r + 1] - hin[r];
      if (d0 > HUB_T) {
        hubs.push_back(r++);
        continue;
      }
      if (d0 > BLK_CAP) {
        med.push_back(r++);
        continue;
      }
      const uint32_t r0 = r;
      while (r < n && r - r0 < BLK_ROWS && hin[r + 1] - hin[r] <= BLK_CAP && hin[r + 1] - hin[r0] <= BLK_CAP) ++r;
      blocks.push_back(r0);
      blocks.push_back(r);
    }
    g->n_med = (uint32_t)med.size();
    G_CUDA(cudaMalloc(&g->med_rows, std::max<size_t>(med.size(), 1) * 4));
    if (!med.empty()) G_CUDA(cudaMemcpy(g->med_rows, med.data(), med.size() * 4, cudaMemcpyHostToDevice));
    // blocks are [r0,r1) pairs; store starts and ends interleaved as consecutive pairs
    g->n_blk = (uint32_t)(blocks.size() / 2);
    g->n_hubs = (uint32_t)hubs.size();
    // pr_pull_kernel reads blk_start[2*blk], blk_start[2*blk+1];
    // gaps (hub rows) make blocks non-contiguous, so keep pairs in a 2*n_blk array
    G_CUDA(cudaMalloc(&g->blk_start, std::max<size_t>(blocks.size(), 2) * 4));
    if (!blocks.empty())
      G_CUDA(cudaMemcpy(g->blk_start, blocks.data(), blocks.size() * 4, cudaMemcpyHostToDevice));
    G_CUDA(cudaMalloc(&g->hubs, std::max<size_t>(hubs.size(), 1) * 4));
    if (!hubs.empty()) G_CUDA(cudaMemcpy(g->hubs, hubs.data(), hubs.size() * 4, cudaMemcpyHostToDevice));
    std::vector<uint32_t> cptr(1, 0), cbeg, cend;
    for (uint32_t hr : hubs) {
      for (uint32_t b = hin[hr]; b < hin[hr + 1]; b += HUB_CHUNK) {
        cbeg.push_back(b);
        cend.push_back(std::min(b + HUB_CHUNK, hin[hr + 1]));
      }
      cptr.push_back((uint32_t)cbeg.size());
    }
    g->n_chunks = (uint32_t)cbeg.size();
    G_CUDA(cudaMalloc(&g->hub_chunk_ptr, cptr.size() * 4));
    G_CUDA(cudaMemcpy(g->hub_chunk_ptr, cptr.data(), cptr.size() * 4, cudaMemcpyHostToDevice));
    G_CUDA(cudaMalloc(&g->chunk_beg, std::max<size_t>(cbeg.size(), 1) * 4));
    G_CUDA(cudaMalloc(&g->chunk_end, std::max<size_t>(cbeg.size(), 1) * 4));
    if (!cbeg.empty()) {
      G_CUDA(cudaMemcpy(g->chunk_beg, cbeg.data(), cbeg.size() * 4, cudaMemcpyHostToDevice));
      G_CUDA(cudaMemcpy(g->chunk_end, cend.data(), cend.size() * 4, cudaMemcpyHostToDevice));
    }
  }
#undef G_CUDA
  *out = g;
  return 0;
}

extern "C" int cozo_gpu_graph_export(cozo_gpu_graph_t* g, uint32_t* out_ptr, uint32_t* out_idx, float* out_w,
                                     uint32_t* in_ptr, uint32_t* in_idx) {
  if (!g) return set_error(COZO_GPU_EINVAL, "null graph handle");
  const size_t np1 = (size_t)g->n + 1;
  if (out_ptr) COZO_CUDA(cudaMemcpy(out_ptr, g->out_ptr, np1 * 4, cudaMemcpyDeviceToHost));
  if (in_ptr) COZO_CUDA(cudaMemcpy(in_ptr, g->in_ptr, np1 * 4, cudaMemcpyDeviceToHost));
  if (out_idx && g->m) COZO_CUDA(cudaMemcpy(out_idx, g->out_idx, g->m * 4, cudaMemcpyDeviceToHost));
  if (in_idx && g->m) COZO_CUDA(cudaMemcpy(in_idx, g->in_idx, g->m * 4, cudaMemcpyDeviceToHost));
  if (out_w && g->m && g->out_w) COZO_CUDA(cudaMemcpy(out_w, g->out_w, g->m * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int cozo_gpu_pagerank(cozo_gpu_graph_t* g, float damping, double tol, uint32_t max_iter,
                                 float* out_scores, uint32_t* out_iters, double* out_err, double* out_kernel_ms,
                                 const volatile int* poison) {
  if (!g || !out_scores) return set_error(COZO_GPU_EINVAL, "null argument");
  if (max_iter == 0) return set_error(COZO_GPU_EINVAL, "iterations must be positive");  // pos_integer_option
  int rc = ensure_init();
  if (rc) return rc;
  if (out_iters) *out_iters = 0;
  if (out_err) *out_err = 0;
  if (out_kernel_ms) *out_kernel_ms = 0;
  const uint32_t n = g->n;
  if (n == 0) return 0;  // pagerank.rs:43-45
  const DeviceInfo& di = device_info();
  DevBuf scores, c0, c1, err, partial, unperm;
  COZO_CUDA(cudaMalloc(&partial.p, (size_t)std::max(g->n_chunks, 1u) * 4));
  COZO_CUDA(cudaMalloc(&scores.p, (size_t)n * 4));
  COZO_CUDA(cudaMalloc(&c0.p, (size_t)n * 4));
  COZO_CUDA(cudaMalloc(&c1.p, (size_t)n * 4));
  COZO_CUDA(cudaMalloc(&unperm.p, (size_t)n * 4));
  COZO_CUDA(cudaMalloc(&err.p, 32));  // f64 error + three u32 work counters
  PrArgs a{};
  a.in_ptr = g->pr_in_ptr;
  a.in_idx = g->pr_in_idx;
  a.od = g->pr_od;
  a.blk_start = g->blk_start;
  a.med_rows = g->med_rows;
  a.chunk_beg = g->chunk_beg;
  a.chunk_end = g->chunk_end;
  a.n_blk = g->n_blk;
  a.n_med = g->n_med;
  a.n_chunks = g->n_chunks;
  a.scores = scores.as<float>();
  a.partial = partial.as<float>();
  a.err = err.as<double>();
  // launch shape: `pagerank.warps` warps per CTA (8 or 32), `pagerank.ctas_per_sm` resident CTAs per SM,
  // `pagerank.dynamic` = draw work from device counters instead of static round-robin
  const int64_t warps = get_option("pagerank.warps", 32);
  const int64_t cps = std::max<int64_t>(1, get_option("pagerank.ctas_per_sm", warps == 32 ? 2 : 8));
  a.dynamic = get_option("pagerank.dynamic", 1) ? 1u : 0u;
  const uint64_t items = (uint64_t)g->n_blk + g->n_med + g->n_chunks;
  const uint32_t wpc = warps == 32 ? 32u : 8u;
  const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((items + wpc - 1) / wpc, (uint64_t)di.sm_count * (uint64_t)cps));
  cudaEvent_t e0, e1;
  COZO_CUDA(cudaEventCreate(&e0));
  COZO_CUDA(cudaEventCreate(&e1));
  const float init = 1.0f / (float)n;
  a.base = (1.0f - damping) / (float)n;
  a.damping = damping;
  COZO_CUDA(cudaEventRecord(e0));
  pr_init_kernel<<<(n + 255) / 256, 256>>>(g->pr_od, n, init, scores.as<float>(), c0.as<float>());
  float* cold = c0.as<float>();
  float* cnew = c1.as<float>();
  uint32_t iter = 0;
  double herr = 0;
  int ret = 0;
  for (;;) {
    if (poisoned(poison)) {
      ret = set_error(COZO_GPU_EKILLED, "Running query is killed before completion");
      break;
    }
    cudaMemsetAsync(err.p, 0, 32);
    a.contrib_old = cold;
    a.contrib_new = cnew;
    if (wpc == 32) pr_pull_kernel<32, 2><<<grid, 1024>>>(a);
    else pr_pull_kernel<8, 8><<<grid, 256>>>(a);
    if (g->n_hubs)
      pr_hub_final_kernel<<<(g->n_hubs + 255) / 256, 256>>>(g->hubs, g->n_hubs, g->hub_chunk_ptr, partial.as<float>(),
                                                            g->pr_od, a.base, damping, cnew, scores.as<float>(),
                                                            err.as<double>());
    cudaError_t ce = cudaMemcpy(&herr, err.p, 8, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) {
      ret = set_error(COZO_GPU_ECUDA, "pagerank iteration failed: %s", cudaGetErrorString(ce));
      break;
    }
    std::swap(cold, cnew);
    ++iter;
    if (herr < tol || iter == max_iter) break;
  }
  if (!ret) pr_unpermute_kernel<<<(n + 255) / 256, 256>>>(scores.as<float>(), g->slot, n, unperm.as<float>());
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ret) return ret;
  COZO_CUDA(cudaMemcpy(out_scores, unperm.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (out_iters) *out_iters = iter;
  if (out_err) *out_err = herr;
  if (out_kernel_ms) *out_kernel_ms = ms;
  return 0;
}

namespace cozo {
// shared driver of the SSSP family: runs sources in chunks that fit `budget` bytes
template <class PerChunk>
static int sssp_chunks(cozo_gpu_graph_t* g, const uint32_t* sources_host, uint32_t n_src, size_t extra_per_src,
                       const volatile int* poison, double* out_ms, PerChunk per_chunk) {
  const uint32_t n = g->n;
  const size_t per_src = (size_t)n * (8 + 8) + extra_per_src;
  size_t freeb = 0, totalb = 0;
  COZO_CUDA(cudaMemGetInfo(&freeb, &totalb));
  size_t budget = std::min<size_t>(freeb / 2, (size_t)8 << 30);
  uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(n_src, budget / std::max<size_t>(per_src, 1)));
  DevBuf state, flags, dsrc;
  COZO_CUDA(cudaMalloc(&state.p, (size_t)chunk * n * 8));
  COZO_CUDA(cudaMalloc(&flags.p, (size_t)chunk * n * 8));
  COZO_CUDA(cudaMalloc(&dsrc.p, (size_t)chunk * 4));
  cudaEvent_t e0, e1;
  COZO_CUDA(cudaEventCreate(&e0));
  COZO_CUDA(cudaEventCreate(&e1));
  COZO_CUDA(cudaEventRecord(e0));
  int ret = 0;
  for (uint32_t s0 = 0; s0 < n_src && !ret; s0 += chunk) {
    if (poisoned(poison)) {
      ret = set_error(COZO_GPU_EKILLED, "Running query is killed before completion");
      break;
    }
    uint32_t c = std::min(chunk, n_src - s0);
    cudaMemcpy(dsrc.p, sources_host + s0, (size_t)c * 4, cudaMemcpyHostToDevice);
    launch_sssp<false>(g, dsrc.as<uint32_t>(), c, state.as<unsigned long long>(), flags.as<uint8_t>(), ForbiddenSets{});
    ret = per_chunk(s0, c, state.as<unsigned long long>(), dsrc.as<uint32_t>());
    cudaError_t ce = cudaDeviceSynchronize();
    if (!ret && ce != cudaSuccess) ret = set_error(COZO_GPU_ECUDA, "sssp failed: %s", cudaGetErrorString(ce));
  }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (out_ms) *out_ms = ms;
  return ret;
}
}  // namespace cozo

extern "C" int cozo_gpu_sssp_multi(cozo_gpu_graph_t* g, const uint32_t* sources, uint32_t n_src, float* out_dist,
                                   uint32_t* out_pred, double* out_kernel_ms, const volatile int* poison) {
  if (!g || (n_src && (!sources || !out_dist))) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (out_kernel_ms) *out_kernel_ms = 0;
  const uint32_t n = g->n;
  if (n_src == 0 || n == 0) return 0;
  for (uint32_t i = 0; i < n_src; ++i)
    if (sources[i] >= n) return set_error(COZO_GPU_EINVAL, "source %u out of range", sources[i]);
  DevBuf dd, dp;
  return sssp_chunks(g, sources, n_src, (size_t)n * 8, poison, out_kernel_ms,
                     [&](uint32_t s0, uint32_t c, unsigned long long* state, uint32_t*) -> int {
                       const uint64_t total = (uint64_t)c * n;
                       if (!dd.p) {
                         COZO_CUDA(cudaMalloc(&dd.p, total * 4));
                         COZO_CUDA(cudaMalloc(&dp.p, total * 4));
                       }
                       sssp_unpack_kernel<<<(uint32_t)((total + 255) / 256), 256>>>(state, total, dd.as<float>(),
                                                                                    dp.as<uint32_t>());
                       COZO_CUDA(cudaMemcpy(out_dist + (size_t)s0 * n, dd.p, total * 4, cudaMemcpyDeviceToHost));
                       if (out_pred)
                         COZO_CUDA(cudaMemcpy(out_pred + (size_t)s0 * n, dp.p, total * 4, cudaMemcpyDeviceToHost));
                       return 0;
                     });
}

extern "C" int cozo_gpu_closeness(cozo_gpu_graph_t* g, float* out, double* out_kernel_ms,
                                  const volatile int* poison) {
  if (!g || !out) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (out_kernel_ms) *out_kernel_ms = 0;
  const uint32_t n = g->n;
  if (n == 0) return 0;  // all_pairs_shortest_path.rs:111-113
  std::vector<uint32_t> sources(n);
  for (uint32_t i = 0; i < n; ++i) sources[i] = i;
  DevBuf dout;
  COZO_CUDA(cudaMalloc(&dout.p, (size_t)n * 4));
  rc = sssp_chunks(g, sources.data(), n, 0, poison, out_kernel_ms,
                   [&](uint32_t s0, uint32_t c, unsigned long long* state, uint32_t*) -> int {
                     closeness_kernel<<<c, 256>>>(state, n, c, s0, dout.as<float>());
                     return 0;
                   });
  if (rc) return rc;
  COZO_CUDA(cudaMemcpy(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int cozo_gpu_betweenness(cozo_gpu_graph_t* g, float* out, double* out_kernel_ms,
                                    const volatile int* poison) {
  if (!g || !out) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (out_kernel_ms) *out_kernel_ms = 0;
  const uint32_t n = g->n;
  if (n == 0) return 0;  // all_pairs_shortest_path.rs:43-46
  std::vector<uint32_t> sources(n);
  for (uint32_t i = 0; i < n; ++i) sources[i] = i;
  DevBuf bc, bcf, sig, del;
  COZO_CUDA(cudaMalloc(&bc.p, (size_t)n * 8));
  COZO_CUDA(cudaMemset(bc.p, 0, (size_t)n * 8));
  COZO_CUDA(cudaMalloc(&bcf.p, (size_t)n * 4));
  size_t sig_cap = 0;
  rc = sssp_chunks(g, sources.data(), n, (size_t)n * 32, poison, out_kernel_ms,
                   [&](uint32_t, uint32_t c, unsigned long long* state, uint32_t* dsrc) -> int {
                     size_t need = (size_t)c * 2 * n * 8;
                     if (sig_cap < need) {
                       if (sig.p) cudaFree(sig.p);
                       if (del.p) cudaFree(del.p);
                       sig.p = del.p = nullptr;
                       COZO_CUDA(cudaMalloc(&sig.p, need));
                       COZO_CUDA(cudaMalloc(&del.p, need));
                       sig_cap = need;
                     }
                     betweenness_kernel<<<c, 256>>>(g->out_ptr, g->out_idx, g->out_w, n, dsrc, c, state,
                                                    sig.as<double>(), del.as<double>(), bc.as<double>());
                     return 0;
                   });
  if (rc) return rc;
  f64_to_f32_kernel<<<(n + 255) / 256, 256>>>(bc.as<double>(), n, bcf.as<float>());
  COZO_CUDA(cudaMemcpy(out, bcf.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int cozo_gpu_clustering(cozo_gpu_graph_t* g, double* out_cc, uint64_t* out_triangles, uint64_t* out_degree,
                                   double* out_kernel_ms, const volatile int* poison) {
  if (!g || !out_cc || !out_triangles || !out_degree) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (out_kernel_ms) *out_kernel_ms = 0;
  const uint32_t n = g->n;
  if (n == 0) return 0;
  if (poisoned(poison)) return set_error(COZO_GPU_EKILLED, "Running query is killed before completion");
  DevBuf cc, nt, dg;
  COZO_CUDA(cudaMalloc(&cc.p, (size_t)n * 8));
  COZO_CUDA(cudaMalloc(&nt.p, (size_t)n * 8));
  COZO_CUDA(cudaMalloc(&dg.p, (size_t)n * 8));
  cudaEvent_t e0, e1;
  COZO_CUDA(cudaEventCreate(&e0));
  COZO_CUDA(cudaEventCreate(&e1));
  cudaEventRecord(e0);
  clustering_kernel<<<(n + 7) / 8, 256>>>(g->out_ptr, g->out_idx, n, cc.as<double>(), nt.as<unsigned long long>(),
                                          dg.as<unsigned long long>());
  cudaEventRecord(e1);
  cudaError_t ce = cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ce != cudaSuccess) return set_error(COZO_GPU_ECUDA, "clustering failed: %s", cudaGetErrorString(ce));
  COZO_CUDA(cudaMemcpy(out_cc, cc.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  COZO_CUDA(cudaMemcpy(out_triangles, nt.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  COZO_CUDA(cudaMemcpy(out_degree, dg.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (out_kernel_ms) *out_kernel_ms = ms;
  return 0;
}

extern "C" int cozo_gpu_sssp_paths(cozo_gpu_graph_t* g, const uint32_t* sources, const uint32_t* goals, uint32_t n_src,
                                   const uint32_t* forb_node_ptr, const uint32_t* forb_nodes,
                                   const uint32_t* forb_edge_ptr, const uint32_t* forb_edge_src,
                                   const uint32_t* forb_edge_dst, uint32_t max_len, float* out_cost, uint32_t* out_len,
                                   uint32_t* out_paths, double* out_kernel_ms, const volatile int* poison) {
  if (!g || (n_src && (!sources || !goals || !out_cost || !out_len || !out_paths)))
    return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (out_kernel_ms) *out_kernel_ms = 0;
  const uint32_t n = g->n;
  if (n_src == 0 || n == 0) return 0;
  if (max_len == 0) return set_error(COZO_GPU_EINVAL, "max_len must be positive");
  for (uint32_t i = 0; i < n_src; ++i)
    if (sources[i] >= n || goals[i] >= n) return set_error(COZO_GPU_EINVAL, "source/goal out of range");
  if (poisoned(poison)) return set_error(COZO_GPU_EKILLED, "Running query is killed before completion");
  const bool forb = forb_node_ptr && forb_edge_ptr;
  const uint32_t n_fn = forb ? forb_node_ptr[n_src] : 0, n_fe = forb ? forb_edge_ptr[n_src] : 0;
  DevBuf state, flags, dsrc, dgoal, fnp, fnn, fep, fes, fed, dc, dl, dp;
  COZO_CUDA(cudaMalloc(&state.p, (size_t)n_src * n * 8));
  COZO_CUDA(cudaMalloc(&flags.p, (size_t)n_src * n * 8));
  COZO_CUDA(cudaMalloc(&dsrc.p, (size_t)n_src * 4));
  COZO_CUDA(cudaMalloc(&dgoal.p, (size_t)n_src * 4));
  COZO_CUDA(cudaMalloc(&dc.p, (size_t)n_src * 4));
  COZO_CUDA(cudaMalloc(&dl.p, (size_t)n_src * 4));
  COZO_CUDA(cudaMalloc(&dp.p, (size_t)n_src * max_len * 4));
  COZO_CUDA(cudaMemcpy(dsrc.p, sources, (size_t)n_src * 4, cudaMemcpyHostToDevice));
  COZO_CUDA(cudaMemcpy(dgoal.p, goals, (size_t)n_src * 4, cudaMemcpyHostToDevice));
  ForbiddenSets fs{};
  if (forb) {
    COZO_CUDA(cudaMalloc(&fnp.p, ((size_t)n_src + 1) * 4));
    COZO_CUDA(cudaMalloc(&fep.p, ((size_t)n_src + 1) * 4));
    COZO_CUDA(cudaMalloc(&fnn.p, std::max<size_t>(n_fn, 1) * 4));
    COZO_CUDA(cudaMalloc(&fes.p, std::max<size_t>(n_fe, 1) * 4));
    COZO_CUDA(cudaMalloc(&fed.p, std::max<size_t>(n_fe, 1) * 4));
    COZO_CUDA(cudaMemcpy(fnp.p, forb_node_ptr, ((size_t)n_src + 1) * 4, cudaMemcpyHostToDevice));
    COZO_CUDA(cudaMemcpy(fep.p, forb_edge_ptr, ((size_t)n_src + 1) * 4, cudaMemcpyHostToDevice));
    if (n_fn) COZO_CUDA(cudaMemcpy(fnn.p, forb_nodes, (size_t)n_fn * 4, cudaMemcpyHostToDevice));
    if (n_fe) {
      COZO_CUDA(cudaMemcpy(fes.p, forb_edge_src, (size_t)n_fe * 4, cudaMemcpyHostToDevice));
      COZO_CUDA(cudaMemcpy(fed.p, forb_edge_dst, (size_t)n_fe * 4, cudaMemcpyHostToDevice));
    }
    fs = ForbiddenSets{fnp.as<uint32_t>(), fnn.as<uint32_t>(), fep.as<uint32_t>(), fes.as<uint32_t>(),
                       fed.as<uint32_t>()};
  }
  cudaEvent_t e0, e1;
  COZO_CUDA(cudaEventCreate(&e0));
  COZO_CUDA(cudaEventCreate(&e1));
  cudaEventRecord(e0);
  if (forb)
    launch_sssp<true>(g, dsrc.as<uint32_t>(), n_src, state.as<unsigned long long>(), flags.as<uint8_t>(), fs);
  else
    launch_sssp<false>(g, dsrc.as<uint32_t>(), n_src, state.as<unsigned long long>(), flags.as<uint8_t>(), fs);
  sssp_path_kernel<<<(n_src + 127) / 128, 128>>>(state.as<unsigned long long>(), n, dsrc.as<uint32_t>(),
                                                 dgoal.as<uint32_t>(), n_src, max_len, dc.as<float>(),
                                                 dl.as<uint32_t>(), dp.as<uint32_t>());
  cudaEventRecord(e1);
  cudaError_t ce = cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ce != cudaSuccess) return set_error(COZO_GPU_ECUDA, "sssp_paths failed: %s", cudaGetErrorString(ce));
  COZO_CUDA(cudaMemcpy(out_cost, dc.p, (size_t)n_src * 4, cudaMemcpyDeviceToHost));
  COZO_CUDA(cudaMemcpy(out_len, dl.p, (size_t)n_src * 4, cudaMemcpyDeviceToHost));
  COZO_CUDA(cudaMemcpy(out_paths, dp.p, (size_t)n_src * max_len * 4, cudaMemcpyDeviceToHost));
  if (out_kernel_ms) *out_kernel_ms = ms;
  return 0;
}
