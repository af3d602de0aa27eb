// pagerank_pb_emu.cpp — runs the propagation-blocking PageRank engine's DEVICE code (cozo_b200/csrc/pagerank_pb.cuh:
// staging kernels + the four passes) under the CPU SIMT emulator and checks the scores against an f64 Jacobi iteration.
// Host orchestration (allocation, CUB sorts/scans, launch geometry) is restated here with std:: algorithms; every line
// of kernel code is the product's own.  Usage: pagerank_pb_emu n m NH GS WIN CHUNK iters seed [star]
#include "cuda_emu.hpp"

#include <algorithm>
#include <numeric>
#include <random>

#include "../../cozo_b200/csrc/pagerank_pb.cuh"

using namespace cozo;

template <class F>
static void launch1d(size_t n_threads_total, unsigned block, F f, const char* name) {
  const unsigned grid = (unsigned)((n_threads_total + block - 1) / block);
  emu::launch(dim3(grid ? grid : 1), block, f, 120.0, name);
}

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  const uint32_t n = (uint32_t)atoi(argv[1]);
  const uint64_t m = (uint64_t)atoll(argv[2]);
  uint32_t NH = (uint32_t)atoi(argv[3]) & ~3u;
  const uint32_t GS = (uint32_t)atoi(argv[4]) & ~3u, WIN = (uint32_t)atoi(argv[5]);
  const uint32_t CH = (uint32_t)atoi(argv[6]) & ~7u;
  const int iters = atoi(argv[7]);
  std::mt19937_64 rng((uint64_t)atoll(argv[8]));
  const bool star = argc > 9;
  // ---- graph: skewed random edges (+ optionally a star: one row longer than many windows) -------------------------
  std::vector<uint32_t> src, dst;
  for (uint64_t e = 0; e < m; ++e) {
    auto pick = [&]() { double u = (double)(rng() >> 11) / 9007199254740992.0; return (uint32_t)(std::pow(u, 3.0) * n); };
    src.push_back(pick() % n);
    dst.push_back(pick() % n);
  }
  if (star)
    for (uint32_t v = 1; v < n; ++v) {
      src.push_back(v);
      dst.push_back(0);
    }
  const uint64_t E = src.size();
  // ---- slot space (pr_stage_slots): slot = rank by out-degree, descending; in-CSR sorted by (dst slot, src slot) ----
  std::vector<uint32_t> od(n, 0);
  for (auto s : src) od[s]++;
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return od[a] > od[b]; });
  std::vector<uint32_t> slot(n), od_slot(n);
  for (uint32_t r = 0; r < n; ++r) {
    slot[order[r]] = r;
    od_slot[r] = od[order[r]];
  }
  std::vector<std::pair<uint32_t, uint32_t>> ed(E);
  for (uint64_t e = 0; e < E; ++e) ed[e] = {slot[dst[e]], slot[src[e]]};
  std::sort(ed.begin(), ed.end());
  std::vector<uint32_t> in_ptr(n + 1, 0), in_idx(std::max<uint64_t>(E, 1));
  for (uint64_t e = 0; e < E; ++e) {
    in_ptr[ed[e].first + 1]++;
    in_idx[e] = ed[e].second;
  }
  for (uint32_t r = 0; r < n; ++r) in_ptr[r + 1] += in_ptr[r];
  // ---- pr_stage_pb, with the product's staging kernels ------------------------------------------------------------
  NH = std::min(NH, n);
  const size_t np1 = (size_t)n + 1;
  std::vector<uint32_t> hptr(np1), mptr(np1);
  launch1d(np1, 256, [&] { pb_row_split_kernel(in_ptr.data(), in_idx.data(), n, NH, hptr.data(), mptr.data()); }, "pb_row_split");
  auto exscan = [](std::vector<uint32_t>& v) {
    uint32_t acc = 0;
    for (auto& x : v) {
      uint32_t t = x;
      x = acc;
      acc += t;
    }
  };
  exscan(hptr);
  exscan(mptr);
  const uint64_t Htot = hptr[n], Mtot = mptr[n];
  if (Htot + Mtot != E) { std::fprintf(stderr, "split lost edges\n"); return 1; }
  uint32_t n_src = 0;
  while (n_src < n && od_slot[n_src] > 0) ++n_src;
  const uint32_t G = n_src > NH ? (uint32_t)(((uint64_t)(n_src - NH) + GS - 1) / GS) : 0;
  const uint32_t NB = (uint32_t)((Mtot + WIN - 1) / WIN);
  std::vector<uint16_t> hub_idx(std::max<uint64_t>(Htot, 8));
  std::vector<uint32_t> key(std::max<uint64_t>(Mtot, 1));
  std::vector<unsigned long long> val(std::max<uint64_t>(Mtot, 1));
  launch1d((size_t)n * 32, 256, [&] { pb_emit_kernel(in_ptr.data(), in_idx.data(), hptr.data(), mptr.data(), n, NH, GS, hub_idx.data(), key.data(), val.data()); }, "pb_emit");
  std::vector<uint16_t> a_src, b_pos;
  std::vector<uint32_t> ctab, rowstart, gbase(std::max(G, 1u)), gcnt(std::max(G, 1u)), gend(std::max(G, 1u));
  std::vector<uint4> items;
  uint64_t pad = 0;
  if (Mtot) {
    std::vector<uint32_t> perm(Mtot);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });  // CUB radix sort is stable
    std::vector<uint32_t> skey(Mtot);
    std::vector<unsigned long long> sval(Mtot);
    for (uint64_t i = 0; i < Mtot; ++i) {
      skey[i] = key[perm[i]];
      sval[i] = val[perm[i]];
    }
    std::vector<uint32_t> gfirst(G, NONE);
    launch1d(Mtot, 256, [&] { pb_group_bounds_kernel(skey.data(), Mtot, gfirst.data()); }, "pb_group_bounds");
    uint32_t next = (uint32_t)Mtot;
    for (uint32_t g = G; g-- > 0;) {
      if (gfirst[g] == NONE) gfirst[g] = next;
      gcnt[g] = next - gfirst[g];
      next = gfirst[g];
    }
    for (uint32_t g = 0; g < G; ++g) {
      gbase[g] = (uint32_t)pad;
      gend[g] = gbase[g] + gcnt[g];
      const uint32_t padded = (gcnt[g] + 7u) & ~7u;
      for (uint32_t c = 0; c < padded; c += CH) items.push_back(uint4{g, gbase[g] + c, gbase[g] + std::min(padded, c + CH), 0});
      pad += padded;
    }
    a_src.assign(pad + 8, 0);
    b_pos.assign(pad + 8, 0);
    ctab.assign((size_t)(NB + 1) * G, 0);
    launch1d((size_t)(NB + 1) * G, 256, [&] { pb_ctab_init_kernel(ctab.data(), NB + 1, G, gend.data()); }, "pb_ctab_init");
    launch1d(Mtot, 256, [&] { pb_place_kernel(skey.data(), sval.data(), Mtot, gfirst.data(), gbase.data(), WIN, G, a_src.data(), b_pos.data(), ctab.data()); }, "pb_place");
    rowstart.assign(NB + 1, 0);
    launch1d(NB + 1, 256, [&] { pb_rowstart_kernel(mptr.data(), n, WIN, NB + 1, rowstart.data()); }, "pb_rowstart");
  }
  // ---- iterations with the product's passes ------------------------------------------------------------------------
  const float damping = 0.85f, init = 1.0f / (float)n, base = (1.0f - damping) / (float)n;
  const size_t slack = 65536 + 8;
  std::vector<float> scores(n, init), c0(n + slack, 0.f), c1(n + slack, 0.f), vals(std::max<uint64_t>(pad, 8) + 8, 0.f), msum(n, 0.f);
  std::vector<float> part_a(std::max(NB, 1u), 0.f), part_z(std::max(NB, 1u), 0.f);
  for (uint32_t r = 0; r < n; ++r) c0[r] = od_slot[r] ? init / (float)od_slot[r] : 0.f;
  alignas(16) unsigned long long err[4];
  PbArgs a{};
  a.n = n; a.NH = NH; a.GS = GS; a.WIN = WIN; a.G = G; a.NB = NB;
  a.hptr = hptr.data(); a.mptr = mptr.data(); a.od = od_slot.data();
  a.hub_idx = hub_idx.data(); a.a_src = a_src.data(); a.b_pos = b_pos.data();
  a.ctab = ctab.data(); a.rowstart = rowstart.data(); a.items = items.data(); a.n_items = (uint32_t)items.size();
  a.scores = scores.data(); a.val = vals.data(); a.msum = msum.data(); a.part_a = part_a.data(); a.part_z = part_z.data();
  a.base = base; a.damping = damping; a.err = err;
  const size_t smemA = (size_t)GS * 4 + 16, smemB = (size_t)WIN * 4 + ((size_t)2 * G + 2) * 4;
  const size_t smemF = (size_t)((NH + 3u) & ~3u) * 4 + 16 + (size_t)(KF_THREADS / 32) * KF_STAGE * 2;
  std::vector<uint8_t> smem(std::max({smemA, smemB, smemF}) + 256);
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
  float* cold = c0.data();
  float* cnew = c1.data();
  double last_err = 0;
  for (int it = 0; it < iters; ++it) {
    std::memset(err, 0, sizeof err);
    a.contrib_old = cold;
    a.contrib_new = cnew;
    if (a.n_items) {
      emu::launch(dim3(std::min<uint32_t>(a.n_items, 3)), 1024, [&] { pb_gather_body(a, sm); }, 300.0, "pb_gather");
      emu::launch(dim3(std::min<uint32_t>(NB, 3)), 512, [&] { pb_accumulate_body<512>(a, sm); }, 300.0, "pb_accumulate");
      launch1d(NB, 256, [&] { pb_straddle_kernel(a); }, "pb_straddle");
    }
    emu::launch(dim3(2), KF_THREADS, [&] { pb_final_body(a, sm); }, 300.0, "pb_final");
    last_err = (double)err[0] / ERR_SCALE;
    std::swap(cold, cnew);
  }
  // ---- f64 Jacobi ground truth in slot space ------------------------------------------------------------------------
  std::vector<double> x(n, (double)init), nx(n);
  double ref_err = 0;
  for (int it = 0; it < iters; ++it) {
    ref_err = 0;
    for (uint32_t r = 0; r < n; ++r) {
      double s = 0;
      for (uint32_t e = in_ptr[r]; e < in_ptr[r + 1]; ++e) s += x[in_idx[e]] / (double)od_slot[in_idx[e]];
      nx[r] = (double)base + (double)damping * s;
      ref_err += std::fabs(nx[r] - x[r]);
    }
    x.swap(nx);
  }
  double worst = 0;
  for (uint32_t r = 0; r < n; ++r) worst = std::max(worst, std::fabs((double)scores[r] - x[r]) / x[r]);
  std::printf("n=%u E=%llu NH=%u GS=%u WIN=%u G=%u NB=%u items=%zu H=%llu M=%llu max_rel_err=%.3e err=%.6e ref_err=%.6e\n", n,
              (unsigned long long)E, NH, GS, WIN, G, NB, items.size(), (unsigned long long)Htot, (unsigned long long)Mtot, worst,
              last_err, ref_err);
  if (!(worst <= 1e-5) || std::fabs(last_err - ref_err) > 0.02 * ref_err + 1e-6) {
    std::fprintf(stderr, "MISMATCH\n");
    return 1;
  }
  std::printf("EMU_OK\n");
  return 0;
}
