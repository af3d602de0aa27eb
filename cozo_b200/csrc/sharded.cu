// sharded.cu — the row-sharded corpus behind the C ABI (SURVEY.md §8b "Multi-GPU", §8e):
// one process per GPU, every rank holds one shard (its own vectors + its own HNSW graph).
//   cozo_gpu_hnsw_search_sharded = query broadcast -> per-shard hnsw_knn -> exchange of the
//   per-shard top-k lists -> k-way merge to global ids, all inside libcozo_gpu.so.
// The reference answers one query at a time from one index relation (query/ra.rs:1102-1119,
// runtime/hnsw.rs:869-1012); the sharded result is the k-NN over the union of the shards as
// searched shard by shard.
//
// Two exchanges:
//   NCCL   ONE ncclAllGather per list (ids, distances), grouped, on the search stream;
//   FUSED  the search kernel's epilogue stores every query's top-k straight into all peers'
//          gather buffers (cudaMalloc memory opened on the peers through CUDA IPC handles, plain
//          st.global over NVLink), then one flag barrier kernel — no collective kernel at all.
// NCCL is resolved with dlopen at first use (the host process may already carry its own copy,
// e.g. the one bundled with PyTorch): there is no link-time dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "hnsw_host.hpp"

namespace cozo {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return 0;
  void* lib = nullptr;
  const char* env = getenv("COZO_GPU_NCCL_LIB");
  if (env && *env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  // a copy the process already mapped (PyTorch's bundled libnccl.so.2) wins over the system one
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return set_error(COZO_GPU_ENODEV, "libnccl.so.2 not found (%s); set COZO_GPU_NCCL_LIB", dlerror());
  NcclApi a;
  a.lib = lib;
#define SYM(field, name)                                                                      \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(lib, name));                            \
  if (!a.field) return set_error(COZO_GPU_ENODEV, "libnccl lacks %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(Broadcast, "ncclBroadcast");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GetVersion, "ncclGetVersion");
#undef SYM
  g_nccl = a;
  return 0;
}

#define COZO_NCCL(call)                                                                                  \
  do {                                                                                                   \
    ncclResult_t _r = (call);                                                                            \
    if (_r != ncclSuccess)                                                                               \
      return ::cozo::set_error(COZO_GPU_ECUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(_r), \
                               __FILE__, __LINE__);                                                      \
  } while (0)

}  // namespace cozo

struct cozo_gpu_shards {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  cozo_gpu_hnsw* shard = nullptr;
  std::vector<uint64_t> offsets;  // global id of every shard's row 0 (row-contiguous partition)
  uint64_t total_rows = 0;
  uint64_t* d_offsets = nullptr;
  cudaStream_t stream = nullptr;
  int exchange = 0;  // 0 = NCCL all-gather, 1 = fused peer stores
  std::string fused_why_not;
  // capacity of the exchange buffers: `tile` queries of `k` results
  uint32_t tile_cap = 0, k_cap = 0;
  // one IPC-shareable block per rank: [flags 256 B][set 0: ids S*tile*k, dist S*tile*k][set 1: ...]
  uint8_t* block = nullptr;
  size_t block_bytes = 0;
  uint8_t* peer_block[COZO_GPU_MAX_PEERS] = {};  // this rank's view of every rank's block (self = block)
  uint32_t* local_ids[2] = {nullptr, nullptr};   // NCCL form: the all-gather's send buffers
  float* local_dist[2] = {nullptr, nullptr};
  uint32_t epoch = 0;
  // running tile counter: consecutive tiles — of one call or of consecutive calls — alternate between the two
  // buffer sets.  A rank may only overwrite a peer's set once that peer has merged what the set held; passing the
  // barrier of tile i proves every peer has finished merging tile i-1 (stream order on the peer), i.e. the OTHER set.
  uint32_t tile_seq = 0;
  int* d_timeout = nullptr;
  // staging of the host-pointer call
  float* d_q = nullptr;
  size_t q_floats = 0;
  uint64_t* d_out_ids = nullptr;
  float* d_out_dist = nullptr;
  uint32_t* d_qstats = nullptr;
  size_t out_rows = 0, out_k = 0;
  std::mutex mu;  // one sharded search at a time per communicator (collectives must not interleave)
};

namespace cozo {

struct PeerFlagPtrs {
  uint32_t* p[COZO_GPU_MAX_PEERS];
};

// Cross-rank barrier over IPC-mapped flag words: thread t tells rank t "my stores of `epoch` have
// landed" and waits for rank t's word in its own block.  Release/acquire at system scope order the
// search kernel's peer stores (earlier on this stream) before the flag.
__global__ void shard_barrier_kernel(uint32_t* my_flags, PeerFlagPtrs peers, int rank, int world, uint32_t epoch,
                                     int* timeout) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.p[t] + rank), "r"(epoch) : "memory");
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + t) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    if (clock64() - t0 > 40000000000ll) {  // ~20 s: a peer died; never hang the GPU
      *timeout = 1;
      break;
    }
    __nanosleep(200);
  }
}

static size_t set_bytes(const cozo_gpu_shards* s, uint32_t tile, uint32_t k) {
  return (size_t)s->world * tile * k * 4;  // one list (ids or dist) of one set
}
static uint32_t* set_ids(uint8_t* blk, const cozo_gpu_shards* s, int b) {
  return reinterpret_cast<uint32_t*>(blk + 256 + (size_t)b * 2 * set_bytes(s, s->tile_cap, s->k_cap));
}
static float* set_dist(uint8_t* blk, const cozo_gpu_shards* s, int b) {
  return reinterpret_cast<float*>(blk + 256 + (size_t)b * 2 * set_bytes(s, s->tile_cap, s->k_cap) +
                                  set_bytes(s, s->tile_cap, s->k_cap));
}

static void close_peers(cozo_gpu_shards* s) {
  for (int r = 0; r < s->world && r < COZO_GPU_MAX_PEERS; ++r) {
    if (s->peer_block[r] && r != s->rank) cudaIpcCloseMemHandle(s->peer_block[r]);
    s->peer_block[r] = nullptr;
  }
}

// (Re)allocate the exchange buffers for `tile` x `k` and, for the fused form, open every peer's block.
// Collective: all ranks call it with the same arguments.
static int ensure_exchange(cozo_gpu_shards* s, uint32_t tile, uint32_t k) {
  if (tile <= s->tile_cap && k <= s->k_cap && s->block) return 0;
  COZO_CUDA(cudaStreamSynchronize(s->stream));
  tile = std::max(tile, s->tile_cap);
  k = std::max(k, s->k_cap);
  close_peers(s);
  if (s->block && s->world > 1) {
    // every rank has closed its mappings of the old blocks before any rank frees one (freeing exported memory
    // that an importer still has open is undefined): one tiny all-gather as the rendezvous
    uint32_t* d_tok = nullptr;
    COZO_CUDA(cudaMalloc(&d_tok, 4 * (size_t)(s->world + 1)));
    COZO_NCCL(g_nccl.AllGather(d_tok + s->world, d_tok, 1, ncclUint32, s->comm, s->stream));
    COZO_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(d_tok);
  }
  if (s->block) cudaFree(s->block);
  s->block = nullptr;
  for (int b = 0; b < 2; ++b) {
    if (s->local_ids[b]) cudaFree(s->local_ids[b]);
    if (s->local_dist[b]) cudaFree(s->local_dist[b]);
    s->local_ids[b] = nullptr;
    s->local_dist[b] = nullptr;
  }
  s->tile_cap = tile;
  s->k_cap = k;
  s->block_bytes = 256 + 4 * set_bytes(s, tile, k);
  COZO_CUDA(cudaMalloc(&s->block, s->block_bytes));
  COZO_CUDA(cudaMemsetAsync(s->block, 0, 256, s->stream));  // ordered before the handle exchange below
  s->epoch = 0;
  s->tile_seq = 0;
  for (int b = 0; b < 2; ++b) {
    COZO_CUDA(cudaMalloc(&s->local_ids[b], (size_t)tile * k * 4));
    COZO_CUDA(cudaMalloc(&s->local_dist[b], (size_t)tile * k * 4));
  }
  s->peer_block[s->rank] = s->block;
  if (s->world == 1) {
    s->exchange = get_option("shard.exchange", 1) ? 1 : 0;
    return 0;
  }
  // exchange the IPC handles with one all-gather, open the peers, agree on the outcome
  int want_fused = get_option("shard.exchange", 1) ? 1 : 0;
  if (s->world > COZO_GPU_MAX_PEERS) want_fused = 0;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  int ok = want_fused;
  if (ok && cudaIpcGetMemHandle(&mine, s->block) != cudaSuccess) {
    s->fused_why_not = cudaGetErrorString(cudaGetLastError());
    ok = 0;
  }
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  struct Rec {
    cudaIpcMemHandle_t h;
    int ok;
    int pad[15];
  };
  static_assert(sizeof(Rec) == 128, "record is one 128-byte line");
  (void)hb;
  Rec rec{};
  rec.h = mine;
  rec.ok = ok;
  Rec* d_rec = nullptr;
  COZO_CUDA(cudaMalloc(&d_rec, sizeof(Rec) * (size_t)(s->world + 1)));
  COZO_CUDA(cudaMemcpyAsync(d_rec + s->world, &rec, sizeof(Rec), cudaMemcpyHostToDevice, s->stream));
  COZO_NCCL(g_nccl.AllGather(d_rec + s->world, d_rec, sizeof(Rec), ncclUint8, s->comm, s->stream));
  std::vector<Rec> all(s->world);
  COZO_CUDA(cudaMemcpyAsync(all.data(), d_rec, sizeof(Rec) * (size_t)s->world, cudaMemcpyDeviceToHost, s->stream));
  COZO_CUDA(cudaStreamSynchronize(s->stream));
  for (int r = 0; r < s->world; ++r) ok &= all[r].ok;
  if (ok) {
    for (int r = 0; r < s->world; ++r) {
      if (r == s->rank) continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        s->fused_why_not = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(cudaGetLastError());
        ok = 0;
        break;
      }
      s->peer_block[r] = static_cast<uint8_t*>(p);
    }
  }
  // second round: did every rank manage to open every peer?
  rec.ok = ok;
  COZO_CUDA(cudaMemcpyAsync(d_rec + s->world, &rec, sizeof(Rec), cudaMemcpyHostToDevice, s->stream));
  COZO_NCCL(g_nccl.AllGather(d_rec + s->world, d_rec, sizeof(Rec), ncclUint8, s->comm, s->stream));
  COZO_CUDA(cudaMemcpyAsync(all.data(), d_rec, sizeof(Rec) * (size_t)s->world, cudaMemcpyDeviceToHost, s->stream));
  COZO_CUDA(cudaStreamSynchronize(s->stream));
  cudaFree(d_rec);
  for (int r = 0; r < s->world; ++r) ok &= all[r].ok;
  if (!ok) {
    close_peers(s);
    s->peer_block[s->rank] = s->block;
    if (want_fused && s->fused_why_not.empty()) s->fused_why_not = "a peer could not map the exchange buffers";
  }
  s->exchange = ok ? 1 : 0;
  return 0;
}

// the device form: everything already in HBM, asynchronous on the communicator's stream
static int sharded_search_dev(cozo_gpu_shards* s, const float* d_q, uint32_t B, uint32_t k, uint32_t ef, double radius,
                              uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_qstats) {
  cozo_gpu_hnsw* h = s->shard;
  const uint32_t dim = h->dev.dim;
  uint32_t tile = (uint32_t)std::max<int64_t>(1, get_option("shard.tile", 65536));
  tile = std::min(tile, std::max(B, 1u));
  int rc = ensure_exchange(s, tile, k);
  if (rc) return rc;
  cudaStream_t st = s->stream;
  for (uint32_t q0 = 0; q0 < B; q0 += tile) {
    const uint32_t nq = std::min(tile, B - q0);
    const int b = (int)(s->tile_seq++ & 1u);  // NOT the tile index of this call: back-to-back single-tile calls must alternate too
    const float* q = d_q + (size_t)q0 * dim;
    uint32_t* qs = d_qstats ? d_qstats + (size_t)q0 * 4 : nullptr;
    if (s->exchange == 1) {
      // gather layout of this tile: [world][nq][k] at the start of set b
      uint64_t dids[COZO_GPU_MAX_PEERS], ddist[COZO_GPU_MAX_PEERS];
      for (int r = 0; r < s->world; ++r) {
        dids[r] = reinterpret_cast<uint64_t>(set_ids(s->peer_block[r], s, b));
        ddist[r] = reinterpret_cast<uint64_t>(set_dist(s->peer_block[r], s, b));
      }
      rc = cozo_gpu_hnsw_search_scatter_dev(h, q, nq, k, ef, radius, (uint32_t)s->world, dids, ddist, (uint32_t)s->rank,
                                            qs, st);
      if (rc) return rc;
      if (s->world > 1) {
        PeerFlagPtrs pf{};
        for (int r = 0; r < s->world; ++r) pf.p[r] = reinterpret_cast<uint32_t*>(s->peer_block[r]);
        ++s->epoch;
        shard_barrier_kernel<<<1, 32, 0, st>>>(reinterpret_cast<uint32_t*>(s->block), pf, s->rank, s->world, s->epoch,
                                               s->d_timeout);
        COZO_CUDA(cudaGetLastError());
      }
    } else {
      rc = cozo_gpu_hnsw_search_dev(h, q, nq, k, ef, radius, s->local_ids[b], s->local_dist[b], nullptr, qs, st);
      if (rc) return rc;
      if (s->world > 1) {
        COZO_NCCL(g_nccl.GroupStart());
        COZO_NCCL(g_nccl.AllGather(s->local_ids[b], set_ids(s->block, s, b), (size_t)nq * k, ncclUint32, s->comm, st));
        COZO_NCCL(g_nccl.AllGather(s->local_dist[b], set_dist(s->block, s, b), (size_t)nq * k, ncclFloat32, s->comm, st));
        COZO_NCCL(g_nccl.GroupEnd());
      } else {
        COZO_CUDA(cudaMemcpyAsync(set_ids(s->block, s, b), s->local_ids[b], (size_t)nq * k * 4, cudaMemcpyDeviceToDevice, st));
        COZO_CUDA(cudaMemcpyAsync(set_dist(s->block, s, b), s->local_dist[b], (size_t)nq * k * 4, cudaMemcpyDeviceToDevice, st));
      }
    }
    rc = cozo_gpu_topk_merge_dev(set_dist(s->block, s, b), set_ids(s->block, s, b), (uint32_t)s->world, nq, k,
                                 s->d_offsets, d_out_ids + (size_t)q0 * k, d_out_dist + (size_t)q0 * k, st);
    if (rc) return rc;
  }
  return 0;
}

static int check_timeout(cozo_gpu_shards* s) {
  int t = 0;
  COZO_CUDA(cudaMemcpy(&t, s->d_timeout, 4, cudaMemcpyDeviceToHost));
  if (t) return set_error(COZO_GPU_ECUDA, "sharded search: a peer rank did not reach the exchange barrier");
  return 0;
}

}  // namespace cozo

using namespace cozo;

extern "C" int cozo_gpu_shards_unique_id(uint8_t* id) {
  if (!id) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId u;
  COZO_NCCL(g_nccl.GetUniqueId(&u));
  static_assert(sizeof(u) == COZO_GPU_UID_BYTES, "unique id size");
  memcpy(id, &u, sizeof(u));
  return 0;
}

extern "C" int cozo_gpu_shards_init(cozo_gpu_shards_t** out, const uint8_t* id, int rank, int world) {
  if (!out || !id) return set_error(COZO_GPU_EINVAL, "null argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return set_error(COZO_GPU_EINVAL, "rank %d of %d", rank, world);
  if (world > COZO_GPU_MAX_PEERS) return set_error(COZO_GPU_EUNSUP, "more than %d shards (one NVSwitch domain)", COZO_GPU_MAX_PEERS);
  int rc = ensure_init();
  if (rc) return rc;
  rc = nccl_load();
  if (rc) return rc;
  auto* s = new cozo_gpu_shards();
  s->rank = rank;
  s->world = world;
  auto fail = [&](int code) {
    cozo_gpu_shards_free(s);
    return code;
  };
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&s->d_timeout, 4) != cudaSuccess || cudaMemset(s->d_timeout, 0, 4) != cudaSuccess ||
      cudaMalloc(&s->d_offsets, (size_t)world * 8) != cudaSuccess)
    return fail(set_error(COZO_GPU_ECUDA, "shard group creation failed: %s", cudaGetErrorString(cudaGetLastError())));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclResult_t r = g_nccl.CommInitRank(&s->comm, world, u, rank);
  if (r != ncclSuccess) return fail(set_error(COZO_GPU_ECUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)));
  *out = s;
  return 0;
}

extern "C" void cozo_gpu_shards_free(cozo_gpu_shards_t* s) {
  if (!s) return;
  if (s->stream) cudaStreamSynchronize(s->stream);
  close_peers(s);
  void* ptrs[] = {s->block,    s->local_ids[0], s->local_ids[1], s->local_dist[0], s->local_dist[1], s->d_offsets,
                  s->d_timeout, s->d_q,         s->d_out_ids,    s->d_out_dist,    s->d_qstats};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

// Collective.  Attaches this rank's shard (staged or built here) and all-gathers the shard sizes:
// shard r owns the global ids [offset_r, offset_r + rows_r) of the row-contiguous partition.
extern "C" int cozo_gpu_hnsw_stage_sharded(cozo_gpu_shards_t* s, cozo_gpu_hnsw_t* local_shard,
                                           uint64_t* out_global_offset, uint64_t* out_total_rows) {
  if (!s || !local_shard) return set_error(COZO_GPU_EINVAL, "null argument");
  int rc = ensure_init();
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(s->mu);
  uint64_t rows = local_shard->dev.n;
  uint64_t* d_tmp = nullptr;
  COZO_CUDA(cudaMalloc(&d_tmp, 8));
  COZO_CUDA(cudaMemcpyAsync(d_tmp, &rows, 8, cudaMemcpyHostToDevice, s->stream));
  if (s->world > 1)
    COZO_NCCL(g_nccl.AllGather(d_tmp, s->d_offsets, 1, ncclUint64, s->comm, s->stream));
  else
    COZO_CUDA(cudaMemcpyAsync(s->d_offsets, d_tmp, 8, cudaMemcpyDeviceToDevice, s->stream));
  std::vector<uint64_t> all(s->world);
  COZO_CUDA(cudaMemcpyAsync(all.data(), s->d_offsets, (size_t)s->world * 8, cudaMemcpyDeviceToHost, s->stream));
  COZO_CUDA(cudaStreamSynchronize(s->stream));
  cudaFree(d_tmp);
  s->offsets.assign(s->world, 0);
  uint64_t acc = 0;
  for (int r = 0; r < s->world; ++r) {
    s->offsets[r] = acc;
    acc += all[r];
  }
  s->total_rows = acc;
  COZO_CUDA(cudaMemcpy(s->d_offsets, s->offsets.data(), (size_t)s->world * 8, cudaMemcpyHostToDevice));
  s->shard = local_shard;
  if (out_global_offset) *out_global_offset = s->offsets[s->rank];
  if (out_total_rows) *out_total_rows = acc;
  return 0;
}

extern "C" int cozo_gpu_shards_info(cozo_gpu_shards_t* s, int* rank, int* world, int* exchange, uint64_t* total_rows) {
  if (!s) return set_error(COZO_GPU_EINVAL, "null argument");
  if (rank) *rank = s->rank;
  if (world) *world = s->world;
  if (exchange) *exchange = s->exchange;
  if (total_rows) *total_rows = s->total_rows;
  return 0;
}

extern "C" int cozo_gpu_hnsw_search_sharded_dev(cozo_gpu_shards_t* s, const float* queries_dev, uint32_t B, uint32_t k,
                                                uint32_t ef, double radius, uint64_t* out_ids_dev, float* out_dist_dev,
                                                uint32_t* per_query_stats_dev, void* stream) {
  if (!s || !s->shard) return set_error(COZO_GPU_EINVAL, "no shard attached (cozo_gpu_hnsw_stage_sharded)");
  if (k == 0) return set_error(COZO_GPU_EINVAL, "k must be positive");
  if (ef == 0) return set_error(COZO_GPU_EINVAL, "ef must be positive");
  if (B && (!queries_dev || !out_ids_dev || !out_dist_dev)) return set_error(COZO_GPU_EINVAL, "null buffer");
  int rc = ensure_init();
  if (rc) return rc;
  if (B == 0) return 0;
  std::lock_guard<std::mutex> lk(s->mu);
  // order the communicator's stream after the caller's stream and back
  cudaEvent_t e;
  COZO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  COZO_CUDA(cudaEventRecord(e, (cudaStream_t)stream));
  COZO_CUDA(cudaStreamWaitEvent(s->stream, e, 0));
  rc = sharded_search_dev(s, queries_dev, B, k, ef, radius, out_ids_dev, out_dist_dev, per_query_stats_dev);
  if (!rc) {
    COZO_CUDA(cudaEventRecord(e, s->stream));
    COZO_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, e, 0));
  }
  cudaEventDestroy(e);
  return rc;
}

extern "C" int cozo_gpu_hnsw_search_sharded(cozo_gpu_shards_t* s, const float* queries, uint32_t B, uint32_t k,
                                            uint32_t ef, double radius, int root, uint64_t* out_ids, float* out_dist,
                                            uint32_t* out_count, CozoGpuSearchStats* stats) {
  if (!s || !s->shard) return set_error(COZO_GPU_EINVAL, "no shard attached (cozo_gpu_hnsw_stage_sharded)");
  if (k == 0) return set_error(COZO_GPU_EINVAL, "k must be positive");
  if (ef == 0) return set_error(COZO_GPU_EINVAL, "ef must be positive");
  if (root >= s->world) return set_error(COZO_GPU_EINVAL, "root %d out of range", root);
  const bool have_q = root < 0 || root == s->rank;
  if (B && have_q && !queries) return set_error(COZO_GPU_EINVAL, "null query buffer");
  if (B && ((out_ids == nullptr) != (out_dist == nullptr))) return set_error(COZO_GPU_EINVAL, "ids/dist must come together");
  int rc = ensure_init();
  if (rc) return rc;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (B == 0) return 0;
  std::lock_guard<std::mutex> lk(s->mu);
  const uint32_t dim = s->shard->dev.dim;
  const size_t qf = (size_t)B * dim;
  if (s->q_floats < qf) {
    if (s->d_q) cudaFree(s->d_q);
    s->d_q = nullptr;
    s->q_floats = 0;
    COZO_CUDA(cudaMalloc(&s->d_q, qf * 4));
    s->q_floats = qf;
  }
  if (s->out_rows < B || s->out_k < k) {
    for (void* p : {(void*)s->d_out_ids, (void*)s->d_out_dist, (void*)s->d_qstats})
      if (p) cudaFree(p);
    s->d_out_ids = nullptr;
    s->d_out_dist = nullptr;
    s->d_qstats = nullptr;
    const size_t rows = std::max<size_t>(B, s->out_rows), kk = std::max<size_t>(k, s->out_k);
    s->out_rows = s->out_k = 0;
    COZO_CUDA(cudaMalloc(&s->d_out_ids, rows * kk * 8));
    COZO_CUDA(cudaMalloc(&s->d_out_dist, rows * kk * 4));
    COZO_CUDA(cudaMalloc(&s->d_qstats, rows * 16));
    s->out_rows = rows;
    s->out_k = kk;
  }
  cudaStream_t st = s->stream;
  cudaEvent_t e0, e1;
  COZO_CUDA(cudaEventCreate(&e0));
  COZO_CUDA(cudaEventCreate(&e1));
  if (have_q) COZO_CUDA(cudaMemcpyAsync(s->d_q, queries, qf * 4, cudaMemcpyHostToDevice, st));
  if (root >= 0 && s->world > 1) COZO_NCCL(g_nccl.Broadcast(s->d_q, s->d_q, qf, ncclFloat32, root, s->comm, st));
  COZO_CUDA(cudaEventRecord(e0, st));
  rc = sharded_search_dev(s, s->d_q, B, k, ef, radius, s->d_out_ids, s->d_out_dist, s->d_qstats);
  if (rc) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
  }
  COZO_CUDA(cudaEventRecord(e1, st));
  if (out_ids) {
    COZO_CUDA(cudaMemcpyAsync(out_ids, s->d_out_ids, (size_t)B * k * 8, cudaMemcpyDeviceToHost, st));
    COZO_CUDA(cudaMemcpyAsync(out_dist, s->d_out_dist, (size_t)B * k * 4, cudaMemcpyDeviceToHost, st));
  }
  std::vector<uint32_t> qs;
  if (stats) {
    qs.resize((size_t)B * 4);
    COZO_CUDA(cudaMemcpyAsync(qs.data(), s->d_qstats, (size_t)B * 16, cudaMemcpyDeviceToHost, st));
  }
  COZO_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  rc = check_timeout(s);
  if (rc) return rc;
  if (out_count && out_ids)
    for (uint32_t i = 0; i < B; ++i) {
      uint32_t c = 0;
      while (c < k && out_ids[(size_t)i * k + c] != ~0ull) ++c;
      out_count[i] = c;
    }
  if (stats) {
    stats->n_queries = B;
    stats->kernel_ms = ms;
    for (uint32_t i = 0; i < B; ++i) {
      stats->dist_evals += qs[(size_t)i * 4 + 0];
      stats->nodes_expanded += qs[(size_t)i * 4 + 1];
      stats->nbr_reads += qs[(size_t)i * 4 + 2];
    }
  }
  return 0;
}
