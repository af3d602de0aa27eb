// graph_emu.cpp — runs the DEVICE code of the FixedRule graph algorithms (cozo_b200/csrc/graph_kernels.cuh) under the
// CPU SIMT emulator: the SSSP kernel in all its frontier forms (flag scan in shared memory / in global memory, compacted
// queues, the wide multi-CTA form) with and without forbidden sets, closeness, betweenness + its ordered reduction (run
// twice: bit-identical), the zero-weight-cycle refusal, clustering.  Checked against a host Dijkstra with f32 path sums
// (bit-exact distances, valid predecessor trees), a host Brandes in f64 and brute-force triangle counts.
// Usage: graph_emu n m seed
#include "cuda_emu.hpp"

#include <algorithm>
#include <map>
#include <queue>
#include <random>
#include <set>

#include "../../cozo_b200/csrc/graph_kernels.cuh"

using namespace cozo;

struct Csr {
  uint32_t n;
  std::vector<uint32_t> ptr, idx;
  std::vector<float> w;
};

static Csr make_csr(uint32_t n, std::vector<std::tuple<uint32_t, uint32_t, float>> e) {
  std::stable_sort(e.begin(), e.end(), [](auto& a, auto& b) { return std::make_pair(std::get<0>(a), std::get<1>(a)) < std::make_pair(std::get<0>(b), std::get<1>(b)); });
  Csr g;
  g.n = n;
  g.ptr.assign(n + 1, 0);
  for (auto& x : e) g.ptr[std::get<0>(x) + 1]++;
  for (uint32_t i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
  for (auto& x : e) {
    g.idx.push_back(std::get<1>(x));
    g.w.push_back(std::get<2>(x));
  }
  return g;
}

// dijkstra with the reference's f32 recursion (shortest_path_dijkstra.rs:304), forbidden sets as :298-303
static std::vector<float> dijkstra(const Csr& g, uint32_t s, const std::set<uint32_t>& fn = {}, const std::set<std::pair<uint32_t, uint32_t>>& fe = {}) {
  std::vector<float> d(g.n, INFINITY);
  using Q = std::pair<float, uint32_t>;
  std::priority_queue<Q, std::vector<Q>, std::greater<Q>> pq;
  d[s] = 0;
  pq.push({0.f, s});
  while (!pq.empty()) {
    auto [du, u] = pq.top();
    pq.pop();
    if (du > d[u]) continue;
    for (uint32_t k = g.ptr[u]; k < g.ptr[u + 1]; ++k) {
      const uint32_t v = g.idx[k];
      if (fn.count(v) || fe.count({u, v})) continue;
      const float nd = du + g.w[k];
      if (nd < d[v]) {
        d[v] = nd;
        pq.push({nd, v});
      }
    }
  }
  return d;
}

static int fails = 0;
#define CHECK(cond, ...)                                   \
  do {                                                     \
    if (!(cond)) {                                         \
      std::fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
      std::fprintf(stderr, __VA_ARGS__);                   \
      std::fprintf(stderr, "\n");                          \
      ++fails;                                             \
    }                                                      \
  } while (0)

static void check_state(const char* what, const Csr& g, const std::vector<unsigned long long>& st, const std::vector<uint32_t>& sources,
                        const std::vector<std::vector<float>>& ref) {
  for (size_t si = 0; si < sources.size(); ++si)
    for (uint32_t v = 0; v < g.n; ++v) {
      const unsigned long long x = st[si * g.n + v];
      const float d = __uint_as_float((uint32_t)(x >> 32));
      const uint32_t p = (uint32_t)x;
      const bool same = (std::isinf(d) && std::isinf(ref[si][v])) || d == ref[si][v];
      CHECK(same, "%s: source %u node %u dist %g expected %g", what, sources[si], v, d, ref[si][v]);
      if (same && std::isfinite(d) && v != sources[si]) {   // predecessor edge realises the distance
        bool ok = false;
        if (p < g.n)
          for (uint32_t k = g.ptr[p]; k < g.ptr[p + 1]; ++k)
            if (g.idx[k] == v && __uint_as_float((uint32_t)(st[si * g.n + p] >> 32)) + g.w[k] == d) ok = true;
        CHECK(ok, "%s: source %u node %u bad predecessor %u", what, sources[si], v, p);
      }
    }
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const uint32_t n = (uint32_t)atoi(argv[1]);
  const uint32_t m = (uint32_t)atoi(argv[2]);
  std::mt19937_64 rng((uint64_t)atoll(argv[3]));
  std::vector<std::tuple<uint32_t, uint32_t, float>> edges;
  for (uint32_t e = 0; e < m; ++e) {
    uint32_t a = rng() % n, b = rng() % n;
    edges.push_back({a, b, (float)(1 + rng() % 32) / 8.0f});   // dyadic weights: many exact ties
  }
  const Csr g = make_csr(n, edges);
  std::vector<uint32_t> sources;
  for (uint32_t s = 0; s < n; s += std::max(1u, n / 5)) sources.push_back(s);
  const uint32_t ns = (uint32_t)sources.size();
  std::vector<std::vector<float>> ref;
  for (auto s : sources) ref.push_back(dijkstra(g, s));
  const size_t fstride = ((size_t)n * 9 + 15) & ~(size_t)15;
  std::vector<unsigned long long> st((size_t)ns * n);
  std::vector<uint8_t> flags(ns * fstride + 64);
  uint8_t* fl = (uint8_t*)(((uintptr_t)flags.data() + 15) & ~(uintptr_t)15);
  std::vector<uint8_t> smem((size_t)n * 10 + 64);
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
  ForbiddenSets none{};
  // 1. default kernel, shared-memory and global-memory flag scan
  emu::launch(dim3(ns), 256, [&] { sssp_body<false, true>(g.ptr.data(), g.idx.data(), g.w.data(), n, sources.data(), ns, st.data(), fl, fstride, none, sm); }, 120, "sssp<smem>");
  check_state("scan/smem", g, st, sources, ref);
  std::fill(st.begin(), st.end(), 0);
  emu::launch(dim3(ns), 256, [&] { sssp_body<false, false>(g.ptr.data(), g.idx.data(), g.w.data(), n, sources.data(), ns, st.data(), fl, fstride, none, sm); }, 120, "sssp<global>");
  check_state("scan/global", g, st, sources, ref);
  // 2. compacted frontier queues
  std::fill(st.begin(), st.end(), 0);
  emu::launch(dim3(ns), 256, [&] { sssp_queue_kernel<false>(g.ptr.data(), g.idx.data(), g.w.data(), n, sources.data(), ns, st.data(), fl, fstride, none); }, 120, "sssp_queue");
  check_state("queue", g, st, sources, ref);
  // 3. wide form: host loop over rounds, as launch_sssp does
  auto run_wide = [&](bool forb, ForbiddenSets fs) {
    std::vector<uint32_t> counts(2 * ns, 0);
    const uint32_t ctas = 3;
    emu::launch(dim3(std::min<uint32_t>(ctas, (n + 255) / 256), ns), 256, [&] { sssp_wide_init_kernel(n, sources.data(), ns, st.data(), fl, fstride, counts.data()); }, 120, "wide_init");
    uint32_t parity = 0, rounds = 0;
    for (; rounds < n + 2; ++rounds) {
      if (forb) emu::launch(dim3(ctas, ns), 256, [&] { sssp_wide_round_kernel<true>(g.ptr.data(), g.idx.data(), g.w.data(), n, ns, st.data(), fl, fstride, counts.data(), parity, fs); }, 120, "wide_round");
      else emu::launch(dim3(ctas, ns), 256, [&] { sssp_wide_round_kernel<false>(g.ptr.data(), g.idx.data(), g.w.data(), n, ns, st.data(), fl, fstride, counts.data(), parity, fs); }, 120, "wide_round");
      emu::launch(dim3((ns + 255) / 256), 256, [&] { sssp_wide_reset_kernel(counts.data(), ns, parity); }, 120, "wide_reset");
      parity ^= 1u;
      bool any = false;
      for (uint32_t i = 0; i < ns; ++i) any |= counts[parity * ns + i] != 0;
      if (!any) break;
    }
    CHECK(rounds < n + 2, "wide form did not settle in n+2 rounds");
    return rounds;
  };
  std::fill(st.begin(), st.end(), 0);
  const uint32_t wr = run_wide(false, none);
  check_state("wide", g, st, sources, ref);
  // 4. forbidden sets in every form (KShortestPathYen's searches)
  std::vector<uint32_t> fnp(ns + 1, 0), fep(ns + 1, 0), fnn, fes, fed;
  std::vector<std::vector<float>> fref;
  for (uint32_t i = 0; i < ns; ++i) {
    std::set<uint32_t> fn;
    std::set<std::pair<uint32_t, uint32_t>> fe;
    for (int t = 0; t < 3; ++t) {
      uint32_t v = rng() % n;
      if (v != sources[i]) fn.insert(v);
      uint32_t u = rng() % n;
      if (g.ptr[u + 1] > g.ptr[u]) fe.insert({u, g.idx[g.ptr[u] + rng() % (g.ptr[u + 1] - g.ptr[u])]});
    }
    for (auto v : fn) fnn.push_back(v);
    for (auto& e : fe) {
      fes.push_back(e.first);
      fed.push_back(e.second);
    }
    fnp[i + 1] = (uint32_t)fnn.size();
    fep[i + 1] = (uint32_t)fes.size();
    fref.push_back(dijkstra(g, sources[i], fn, fe));
  }
  fnn.push_back(0); fes.push_back(0); fed.push_back(0);
  ForbiddenSets fs{fnp.data(), fnn.data(), fep.data(), fes.data(), fed.data()};
  std::fill(st.begin(), st.end(), 0);
  emu::launch(dim3(ns), 256, [&] { sssp_body<true, true>(g.ptr.data(), g.idx.data(), g.w.data(), n, sources.data(), ns, st.data(), fl, fstride, fs, sm); }, 120, "sssp<forb,smem>");
  check_state("forb/scan", g, st, sources, fref);
  std::fill(st.begin(), st.end(), 0);
  emu::launch(dim3(ns), 256, [&] { sssp_queue_kernel<true>(g.ptr.data(), g.idx.data(), g.w.data(), n, sources.data(), ns, st.data(), fl, fstride, fs); }, 120, "sssp_queue<forb>");
  check_state("forb/queue", g, st, sources, fref);
  std::fill(st.begin(), st.end(), 0);
  run_wide(true, fs);
  check_state("forb/wide", g, st, sources, fref);
  // 5. closeness + betweenness over ALL sources (state from the default kernel)
  std::vector<uint32_t> all(n);
  for (uint32_t i = 0; i < n; ++i) all[i] = i;
  std::vector<unsigned long long> sta((size_t)n * n);
  std::vector<uint8_t> flags2((size_t)n * fstride + 64);
  uint8_t* fl2 = (uint8_t*)(((uintptr_t)flags2.data() + 15) & ~(uintptr_t)15);
  emu::launch(dim3(n), 256, [&] { sssp_body<false, true>(g.ptr.data(), g.idx.data(), g.w.data(), n, all.data(), n, sta.data(), fl2, fstride, none, sm); }, 300, "sssp all");
  std::vector<float> clo(n, -1.f);
  emu::launch(dim3(n), 256, [&] { closeness_kernel(sta.data(), n, n, 0, clo.data()); }, 300, "closeness");
  for (uint32_t s = 0; s < n; ++s) {
    auto d = dijkstra(g, s);
    double tot = 0;
    uint32_t cnt = 0;
    for (auto x : d)
      if (std::isfinite(x)) {
        tot += x;
        ++cnt;
      }
    const float exp = (float)cnt * (float)cnt / (float)tot / (float)(n - 1);
    CHECK((std::isnan(exp) && std::isnan(clo[s])) || (std::isinf(exp) && std::isinf(clo[s])) || std::fabs(clo[s] - exp) <= 1e-6f * std::fabs(exp), "closeness %u: %g vs %g", s, clo[s], exp);
  }
  std::vector<double> sig((size_t)n * 2 * n), del((size_t)n * 2 * n), bc(n, 0.0), bc2(n, 0.0);
  int cyclic = 0;
  auto run_bc = [&](std::vector<double>& out) {
    emu::launch(dim3(n), 256, [&] { betweenness_kernel(g.ptr.data(), g.idx.data(), g.w.data(), n, all.data(), n, sta.data(), sig.data(), del.data(), &cyclic); }, 600, "betweenness");
    emu::launch(dim3((n + 255) / 256), 256, [&] { betweenness_reduce_kernel(del.data(), n, n, out.data()); }, 120, "betweenness_reduce");
  };
  run_bc(bc);
  run_bc(bc2);
  CHECK(cyclic == 0, "cyclic flag on an acyclic tie graph");
  CHECK(std::memcmp(bc.data(), bc2.data(), n * 8) == 0, "betweenness differs between two runs");
  {  // host Brandes over the tie DAG (f64), the form betweenness_kernel implements
    std::vector<double> ref_bc(n, 0.0);
    for (uint32_t s = 0; s < n; ++s) {
      auto d = dijkstra(g, s);
      std::vector<uint32_t> ord;
      for (uint32_t v = 0; v < n; ++v)
        if (std::isfinite(d[v])) ord.push_back(v);
      std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return d[a] < d[b]; });
      std::vector<double> sigma(n, 0.0), delta(n, 0.0);
      sigma[s] = 1.0;
      // zero weights do not occur here, so the distance order is a topological order of the tie DAG
      for (uint32_t u : ord)
        for (uint32_t k = g.ptr[u]; k < g.ptr[u + 1]; ++k) {
          const uint32_t v = g.idx[k];
          if (v != s && d[u] + g.w[k] == d[v]) sigma[v] += sigma[u];
        }
      for (size_t i = ord.size(); i-- > 0;) {
        const uint32_t u = ord[i];
        for (uint32_t k = g.ptr[u]; k < g.ptr[u + 1]; ++k) {
          const uint32_t v = g.idx[k];
          if (v != s && d[u] + g.w[k] == d[v]) delta[u] += sigma[u] / sigma[v] * (1.0 + delta[v]);
        }
      }
      for (uint32_t v = 0; v < n; ++v)
        if (v != s) ref_bc[v] += delta[v];
    }
    for (uint32_t v = 0; v < n; ++v) CHECK(std::fabs(bc[v] - ref_bc[v]) <= 1e-9 * std::max(1.0, ref_bc[v]), "betweenness %u: %.12g vs %.12g", v, bc[v], ref_bc[v]);
  }
  {  // zero-weight cycle: sigma never settles -> the cyclic flag, and the kernel returns
    // 0 -> 1 (1.0), 1 <-> 2 at weight 0, 2 -> 3 (1.0): seen from source 0 the tie graph has the cycle 1 <-> 2.
    // (A zero-weight cycle THROUGH the source is harmless: edges into the source are never tie edges, and the
    // reference's path enumeration stops at the start node, shortest_path_dijkstra.rs:410-421.)
    const Csr z = make_csr(4, {{0, 1, 1.f}, {1, 2, 0.f}, {2, 1, 0.f}, {2, 3, 1.f}});
    std::vector<uint32_t> zs{0, 1, 2, 3};
    std::vector<unsigned long long> zst(16);
    emu::launch(dim3(4), 256, [&] { sssp_body<false, true>(z.ptr.data(), z.idx.data(), z.w.data(), 4, zs.data(), 4, zst.data(), fl2, fstride, none, sm); }, 60, "sssp zero-cycle");
    std::vector<double> zs1(4 * 8), zd1(4 * 8);
    int zc = 0;
    emu::launch(dim3(4), 256, [&] { betweenness_kernel(z.ptr.data(), z.idx.data(), z.w.data(), 4, zs.data(), 4, zst.data(), zs1.data(), zd1.data(), &zc); }, 60, "betweenness zero-cycle");
    CHECK(zc == 1, "zero-weight cycle not flagged");
    CHECK(__uint_as_float((uint32_t)(zst[0 * 4 + 3] >> 32)) == 2.f, "distances over a zero-weight cycle");
    // the cycle through the source: must NOT be flagged
    const Csr y = make_csr(4, {{0, 1, 0.f}, {1, 0, 0.f}, {1, 2, 1.f}, {2, 3, 1.f}});
    emu::launch(dim3(4), 256, [&] { sssp_body<false, true>(y.ptr.data(), y.idx.data(), y.w.data(), 4, zs.data(), 4, zst.data(), fl2, fstride, none, sm); }, 60, "sssp cycle via source");
    int yc = 0;
    emu::launch(dim3(4), 256, [&] { betweenness_kernel(y.ptr.data(), y.idx.data(), y.w.data(), 4, zs.data(), 4, zst.data(), zs1.data(), zd1.data(), &yc); }, 60, "betweenness cycle via source");
    CHECK(yc == 0, "a zero-weight cycle through the source was flagged");
  }
  // 6. clustering on the mirrored graph
  {
    std::vector<std::tuple<uint32_t, uint32_t, float>> me;
    for (auto& e : edges) {
      me.push_back({std::get<0>(e), std::get<1>(e), 1.f});
      me.push_back({std::get<1>(e), std::get<0>(e), 1.f});
    }
    const Csr u = make_csr(n, me);
    std::vector<double> cc(n);
    std::vector<unsigned long long> tri(n), deg(n);
    emu::launch(dim3((n + 7) / 8), 256, [&] { clustering_kernel(u.ptr.data(), u.idx.data(), n, cc.data(), tri.data(), deg.data()); }, 300, "clustering");
    for (uint32_t x = 0; x < n; ++x) {   // triangles.rs:59-98: positions i, j of the sorted duplicate-keeping list with value[j] < value[i]
      unsigned long long t = 0;
      const uint32_t b = u.ptr[x], e = u.ptr[x + 1];
      for (uint32_t i = b; i < e; ++i)
        for (uint32_t j = b; j < e; ++j)
          if (u.idx[j] < u.idx[i]) {
            const uint32_t a = u.idx[i], v = u.idx[j];
            if (std::binary_search(u.idx.begin() + u.ptr[a], u.idx.begin() + u.ptr[a + 1], v)) ++t;
          }
      CHECK(tri[x] == t && deg[x] == e - b, "clustering node %u: %llu vs %llu", x, tri[x], t);
    }
  }
  std::printf("n=%u m=%u sources=%u wide_rounds=%u fails=%d\n", n, m, ns, wr, fails);
  if (fails) return 1;
  std::printf("EMU_OK\n");
  return 0;
}
