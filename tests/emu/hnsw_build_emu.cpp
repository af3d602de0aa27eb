// hnsw_build_emu.cpp — runs the index builder's DEVICE code (cozo_b200/csrc/hnsw_build_kernels.cuh + hnsw_device.cuh:
// K1 batch search with the TMA ring, K2 heuristic selection with optional candidate extension, K4 link / shrink) under
// the CPU SIMT emulator in FIDELITY mode (batches of one node, as hnsw_insert_range does for max_batch = 1 and for
// extend_candidates), then writes the graph.  tests/test_emu_cpu.py builds the same index with the oracle's faithful
// builder (same levels) and compares the two graphs edge for edge.
// With max_batch > 1 it runs the DEFAULT (throughput) mode instead: batches of min(max_batch, linked/16) nodes that do
// not see each other, K1 with 4 warps per CTA, in-edges sorted by (layer, target) and linked one warp per target.
// Usage: hnsw_build_emu in.bin out.bin n dim m ef_c keep_pruned extend [max_batch = 1]
//   in.bin  = f32 X[n*dim] then u8 level[n];  out.bin = per level L: u32 n_rows, then per row: u32 id, u32 deg, u32 ids[deg]
#include "cuda_emu.hpp"

#include <algorithm>
#include <numeric>

#include "../../cozo_b200/csrc/hnsw_build_kernels.cuh"

using namespace cozo;
constexpr int NV = 1;  // dim <= 128

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  const char *fin = argv[1], *fout = argv[2];
  const uint32_t n = (uint32_t)atoi(argv[3]), dim = (uint32_t)atoi(argv[4]), m = (uint32_t)atoi(argv[5]);
  const uint32_t ef_c = (uint32_t)atoi(argv[6]);
  const int keep_pruned = atoi(argv[7]), extend = atoi(argv[8]);
  uint32_t max_batch = argc > 9 ? (uint32_t)atoi(argv[9]) : 1u;
  if (extend || max_batch == 0) max_batch = 1;  // as hnsw_insert_range
  max_batch = std::min(max_batch, n);
  if (dim > 128 || dim % 4) return 2;
  std::vector<float> X((size_t)n * dim);
  std::vector<uint8_t> level(n);
  {
    FILE* f = fopen(fin, "rb");
    if (!f || fread(X.data(), 4, X.size(), f) != X.size() || fread(level.data(), 1, n, f) != n) return 2;
    fclose(f);
  }
  // ---- handle state as cozo_gpu_hnsw_build lays it out ------------------------------------------------------------
  HnswDev g{};
  g.n = n;
  g.dim = dim;
  g.ld = dim;
  g.metric = COZO_GPU_L2;
  g.s0 = round_up(2 * m, 32);
  g.su = round_up(m, 32);
  g.entry = NONE;
  g.top_level = 0;
  std::vector<uint32_t> upper_off(n, NONE);
  uint64_t up_rows = 0;
  for (uint32_t i = 0; i < n; ++i)
    if (level[i]) {
      upper_off[i] = (uint32_t)up_rows;
      up_rows += level[i];
    }
  std::vector<uint32_t> adj0((size_t)n * g.s0, NONE), adj_up((size_t)std::max<uint64_t>(up_rows, 1) * g.su, NONE);
  std::vector<float> adj0_d((size_t)n * g.s0, 0.f), adj_up_d((size_t)std::max<uint64_t>(up_rows, 1) * g.su, 0.f);
  std::vector<uint32_t> deg0(n, 0), deg_up(std::max<uint64_t>(up_rows, 1), 0);
  std::vector<float> vec_al((size_t)n * dim + 8);
  float* vec = (float*)(((uintptr_t)vec_al.data() + 15) & ~(uintptr_t)15);   // rows 16-byte aligned (cp.async.bulk)
  std::memcpy(vec, X.data(), X.size() * 4);
  g.vec = vec;
  g.adj0 = adj0.data();
  g.upper_off = upper_off.data();
  g.adj_up = adj_up.data();
  BuildDev b{};
  b.adj0 = adj0.data();
  b.adj0_d = adj0_d.data();
  b.deg0 = deg0.data();
  b.adj_up = adj_up.data();
  b.adj_up_d = adj_up_d.data();
  b.deg_up = deg_up.data();
  b.node_level = level.data();
  b.m_max0 = 2 * m;
  b.m_max = m;
  b.keep_pruned = keep_pruned;
  b.extend = extend;
  b.upper_off = upper_off.data();
  const uint32_t mcap = std::max(b.m_max0, b.m_max);
  uint32_t max_lvl = 0;
  for (auto l : level) max_lvl = std::max<uint32_t>(max_lvl, l);
  const uint32_t Tcap = max_batch * (max_lvl + 1);
  std::vector<unsigned long long> ext_keys;
  std::vector<float> ext_d;
  std::vector<uint32_t> ext_id;
  if (extend) {
    uint32_t cap = 32;
    const uint32_t need = std::max(ef_c, mcap + 32) * (1 + std::max(g.s0, g.su));
    while (cap < need) cap <<= 1;
    b.ext_cap = cap;
    const size_t slices = std::max<size_t>(Tcap, 4);
    ext_keys.resize(slices * cap);
    ext_d.resize(slices * cap);
    ext_id.resize(slices * cap);
    b.ext_keys = ext_keys.data();
    b.ext_d = ext_d.data();
    b.ext_id = ext_id.data();
  }
  // ---- scratch of hnsw_insert_range ---------------------------------------------------------------------------------
  const uint32_t ns = 4;
  const SmemLayout lay = make_layout(ef_c, ns, g.ld);
  const uint32_t wpc = 4;  // warps per CTA of K1
  const uint32_t nwords = round_up((n + 31) / 32, 4), logcap = std::max<uint32_t>(4096u, 64u * ef_c);
  const uint32_t max_grid1 = (max_batch + wpc - 1) / wpc;
  std::vector<uint32_t> vis((size_t)max_grid1 * wpc * nwords, 0), vlog((size_t)max_grid1 * wpc * logcap), counters(16, 0);
  const uint64_t req_cap = (uint64_t)Tcap * mcap;
  std::vector<uint32_t> coff(max_batch + 1), lnode(Tcap), llevel(Tcap), cand_id((size_t)Tcap * ef_c), cand_cnt(Tcap), req_src(req_cap), perm(req_cap),
      perm2(req_cap), heads(req_cap);
  std::vector<float> cand_d((size_t)Tcap * ef_c), req_d(req_cap);
  std::vector<unsigned long long> req_key(req_cap), req_key2(req_cap);
  std::iota(perm.begin(), perm.end(), 0);
  std::vector<uint8_t> smem_buf((size_t)lay.warp_bytes * 4 + (size_t)4 * (4 * (mcap + 32) + 2 * mcap) * 4 + 512);
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_buf.data() + 127) & ~(uintptr_t)127);
  uint32_t inserted = 0;
  g.entry = 0;  // first vector: fresh self-loops only (hnsw.rs:360-373)
  g.top_level = level[0];
  inserted = 1;
  uint32_t n_live = 1;
  while (inserted < n) {
    const uint32_t bs = std::min<uint32_t>(std::min<uint32_t>(max_batch, std::max<uint32_t>(1u, n_live / 16)), n - inserted);
    const uint32_t top = g.top_level;
    uint32_t nl = 0;
    for (uint32_t i = 0; i < bs; ++i) {
      coff[i] = nl;
      const uint32_t cnt = std::min<uint32_t>(level[inserted + i], top) + 1;
      for (uint32_t L = 0; L < cnt; ++L) {
        lnode[nl] = inserted + i;
        llevel[nl] = L;
        ++nl;
      }
    }
    coff[bs] = nl;
    std::fill(counters.begin(), counters.end(), 0);
    BatchParams p{};
    p.begin = inserted;
    p.count = bs;
    p.top = top;
    p.ef_c = ef_c;
    p.coff = coff.data();
    p.list_node = lnode.data();
    p.list_level = llevel.data();
    p.T = nl;
    p.cand_d = cand_d.data();
    p.cand_id = cand_id.data();
    p.cand_cnt = cand_cnt.data();
    p.req_key = req_key.data();
    p.req_src = req_src.data();
    p.req_d = req_d.data();
    p.req_count = counters.data() + 1;
    p.counter = counters.data() + 0;
    p.vis = vis.data();
    p.nwords = nwords;
    p.vlog = vlog.data();
    p.logcap = logcap;
    p.ns = ns;
    p.lay = lay;
    if (max_batch == 1)
      emu::launch(dim3(1), 32, [&] { build_search_body<NV, COZO_GPU_L2>(g, b, p, smem); }, 120, "K1 build_search");
    else
      emu::launch(dim3((bs + wpc - 1) / wpc), wpc * 32, [&] { build_search_body<NV, COZO_GPU_L2>(g, b, p, smem); }, 240, "K1 build_search");
    emu::launch(dim3((nl + 3) / 4), 128, [&] { build_select_body<NV, COZO_GPU_L2>(g, b, p, smem); }, 120, "K2 build_select");
    const uint32_t nreq = counters[1];
    if (nreq > req_cap) { std::fprintf(stderr, "in-edge queue overflow\n"); return 1; }
    if (nreq && extend) {
      for (uint32_t r = 0; r < nreq; ++r)
        emu::launch(dim3(1), 32, [&] { build_link_body<NV, COZO_GPU_L2>(g, b, req_key.data(), perm.data(), req_src.data(), req_d.data(), nreq, perm.data() + r, 1, smem); }, 120, "K4 build_link (sequential)");
    } else if (nreq) {
      std::iota(perm2.begin(), perm2.begin() + nreq, 0);
      std::stable_sort(perm2.begin(), perm2.begin() + nreq, [&](uint32_t x, uint32_t y) { return req_key[x] < req_key[y]; });
      for (uint32_t i = 0; i < nreq; ++i) req_key2[i] = req_key[perm2[i]];
      uint32_t nheads = 0;
      for (uint32_t r = 0; r < nreq; ++r)
        if (r == 0 || req_key2[r] != req_key2[r - 1]) heads[nheads++] = r;
      emu::launch(dim3((nheads + 3) / 4), 128, [&] { build_link_body<NV, COZO_GPU_L2>(g, b, req_key2.data(), perm2.data(), req_src.data(), req_d.data(), nreq, heads.data(), nheads, smem); }, 120, "K4 build_link");
    }
    for (uint32_t i = 0; i < bs; ++i) {
      const uint32_t id = inserted + i;
      if (level[id] > g.top_level) {
        g.top_level = level[id];
        g.entry = id;
      }
    }
    inserted += bs;
    n_live += bs;
  }
  for (auto v : vis)
    if (v) { std::fprintf(stderr, "visited bitmap not clean after the build\n"); return 1; }
  // ---- dump -----------------------------------------------------------------------------------------------------------
  FILE* f = fopen(fout, "wb");
  const uint32_t n_levels = g.top_level + 1;
  fwrite(&n_levels, 4, 1, f);
  fwrite(&g.entry, 4, 1, f);
  for (uint32_t L = 0; L < n_levels; ++L) {
    uint32_t rows = 0;
    for (uint32_t i = 0; i < n; ++i) rows += level[i] >= L;
    fwrite(&rows, 4, 1, f);
    for (uint32_t i = 0; i < n; ++i) {
      if (level[i] < L) continue;
      const uint32_t* row = L == 0 ? adj0.data() + (size_t)i * g.s0 : adj_up.data() + (size_t)(upper_off[i] + L - 1) * g.su;
      const uint32_t stride = L == 0 ? g.s0 : g.su;
      uint32_t deg = 0;
      while (deg < stride && row[deg] != NONE) ++deg;
      fwrite(&i, 4, 1, f);
      fwrite(&deg, 4, 1, f);
      fwrite(row, 4, deg, f);
    }
  }
  fclose(f);
  std::printf("built n=%u levels=%u entry=%u\nEMU_OK\n", n, n_levels, g.entry);
  return 0;
}
