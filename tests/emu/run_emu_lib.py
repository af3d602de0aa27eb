"""Scenarios that drive libcozo_gpu_emu.so (the library's own .cu files compiled for the CPU, tests/emu/build_emu_lib.py)
through the same ctypes binding and against the same oracle as the `-m gpu` tests, at sizes an emulator can finish.
Usage: run_emu_lib.py <path to libcozo_gpu_emu.so> <scenario>.  Prints EMU_OK at the end.  TEST INFRASTRUCTURE ONLY.

It runs in its own process because it points cozo_b200.capi at the emulated library before the first load: the product's
loader has no such switch (no CPU fallback), this script sets the module attribute by hand.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cozo_b200 import capi  # noqa: E402

capi.LIB_PATH = os.path.abspath(sys.argv[1])
assert capi._lib is None
from oracle import oracle as O  # noqa: E402
from tests.test_graph_gpu import _check_tree, _random_graph  # noqa: E402

T0 = time.time()
LONG = True                                          # the full matrix: seconds since the emulator schedules lanes as fibers


def step(msg):
    print(f"[{time.time() - T0:7.1f}s] {msg}", flush=True)


def opts(**kv):
    class _Ctx:
        def __enter__(self):
            for k, v in kv.items():
                capi.set_option(k.replace("__", "."), v)

        def __exit__(self, *a):
            for k in kv:
                capi.set_option(k.replace("__", "."), -1 if k.startswith("hnsw") else 0)
    return _Ctx()


# ---------------------------------------------------------------------------------------------------------------------
def scenario_graph():
    """graph.cu end to end: staging, SSSP in the three frontier forms, closeness, betweenness, clustering, constrained paths"""
    n = 60
    src, dst, w = _random_graph(n, 420, 1)
    g = capi.Graph(n, src, dst, w)
    o = O.OracleGraph(n, src, dst, w)
    for a, b in zip(g.export(), o.export()):
        assert np.array_equal(a, b)
    step("CSR staging == oracle")
    for bad in (([0, 5], [1, 2], None), ([0, 1], [1, 2], [1.0, -2.0]), ([0, 1], [1, 2], [1.0, np.inf])):
        try:
            capi.Graph(3, bad[0], bad[1], bad[2])
            raise AssertionError("bad edges accepted")
        except capi.CozoGpuError:
            pass
    sources = np.arange(0, n, 7, dtype=np.uint32)
    od, _ = o.sssp(sources, n_threads=2)
    oc = o.closeness(n_threads=2)
    fin = np.isfinite(oc)
    for name, kv in (("flag scan (default)", {}), ("compacted frontier", {"sssp__frontier": 1}), ("wide", {"sssp__wide": 1})):
        with opts(**kv):
            gd, gp, _ = g.sssp(sources)
            assert np.array_equal(gd, od), name
            _check_tree(src, dst, w, sources, gd, gp, n)
            if name != "wide":           # closeness = n sources: the wide form (one launch per round) is for few sources
                gc, _ = g.closeness()
                assert np.array_equal(np.isfinite(gc), fin) and np.allclose(gc[fin], oc[fin], rtol=1e-5), name
        step(f"sssp + closeness, {name}: distances bit-identical, predecessor trees valid")
    # unweighted, with unreachable nodes and zero weights
    s2 = np.array([0, 1, 2, 5, 5], np.uint32)
    d2 = np.array([1, 2, 0, 6, 5], np.uint32)
    w2 = np.array([0, 0.5, 0, 2, 0], np.float32)
    g2, o2 = capi.Graph(8, s2, d2, w2), O.OracleGraph(8, s2, d2, w2)
    assert np.array_equal(g2.sssp(np.arange(8, dtype=np.uint32))[0], o2.sssp(np.arange(8, dtype=np.uint32))[0])
    gu, ou = capi.Graph(n, src, dst), O.OracleGraph(n, src, dst, np.ones(src.size, np.float32))   # no weights == all ones
    assert np.array_equal(gu.sssp(sources)[0], ou.sssp(sources)[0])
    step("zero weights / unreachable / unweighted")
    a, _ = g.betweenness()
    b, _ = g.betweenness()
    ob = o.betweenness(n_threads=2)
    assert np.array_equal(a, b) and np.allclose(a, ob, rtol=1e-4, atol=1e-6)
    step("betweenness == oracle, rerun bit-identical (ordered reduction)")
    zs, zd, zw = np.array([0, 1, 0, 2], np.uint32), np.array([1, 3, 2, 3], np.uint32), np.array([0, 1, 0, 1], np.float32)
    assert np.allclose(capi.Graph(4, zs, zd, zw).betweenness()[0], O.OracleGraph(4, zs, zd, zw).betweenness())
    cs, cd, cw = np.array([0, 1, 2, 2], np.uint32), np.array([1, 2, 1, 3], np.uint32), np.array([1, 0, 0, 1], np.float32)
    try:
        capi.Graph(4, cs, cd, cw).betweenness()
        raise AssertionError("zero-weight cycle accepted")
    except capi.CozoGpuError as e:
        assert e.code == capi.E_UNSUP
    assert capi.Graph(4, cs, cd, cw).sssp([0])[0][0].tolist() == [0, 1, 1, 2]
    step("zero-weight tie DAG accepted, zero-weight cycle refused (EUNSUP)")
    ms, md = np.concatenate([src, dst]), np.concatenate([dst, src])
    gc_, gt, gdg, _ = capi.Graph(n, ms, md).clustering()
    oc_, ot, odg = O.OracleGraph(n, ms, md).clustering(n_threads=2)
    assert np.array_equal(gt, ot) and np.array_equal(gdg, odg) and np.array_equal(gc_, oc_) and ot.max() > 0
    step("clustering bit-identical")
    rng = np.random.default_rng(8)
    pairs = sorted({(int(a), int(b)) for a, b in zip(rng.integers(0, n, 500), rng.integers(0, n, 500)) if a != b})
    ps = np.array([p[0] for p in pairs], np.uint32)
    pd = np.array([p[1] for p in pairs], np.uint32)
    pw = (rng.random(ps.size) * 10 + 0.5).astype(np.float32)
    gp_ = capi.Graph(n, ps, pd, pw)
    S = 12
    srcs, goals = rng.integers(0, n, S).astype(np.uint32), rng.integers(0, n, S).astype(np.uint32)
    fn = [[int(x) for x in rng.integers(0, n, rng.integers(0, 4)) if x != srcs[i]] for i in range(S)]
    fe = [[pairs[int(j)] for j in rng.integers(0, len(pairs), rng.integers(0, 5))] for i in range(S)]
    res, _ = gp_.sssp_paths(srcs, goals, fn, fe, max_len=4)
    for i in range(S):
        keep = np.array([(a, b) not in set(fe[i]) and b not in set(fn[i]) for a, b in pairs])
        oo = O.OracleGraph(n, ps[keep], pd[keep], pw[keep])
        dd, bb = oo.sssp([int(srcs[i])])
        cost, path = res[i]
        exp = float(dd[0, goals[i]])
        if np.isinf(exp):
            assert np.isinf(cost) and path == []
            continue
        assert cost == exp
        p, cur = [], int(goals[i])
        while cur != int(srcs[i]):
            p.append(cur)
            cur = int(bb[0, cur])
        p.append(int(srcs[i]))
        assert path == p[::-1]
    for kv in ({"sssp__frontier": 1}, {"sssp__wide": 1}):
        with opts(**kv):
            assert gp_.sssp_paths(srcs, goals, fn, fe, max_len=16)[0] == res
    step("constrained paths (forbidden nodes / edges) == oracle on the pruned graph, in all three frontier forms")
    # poison set before the call
    flag = np.ones(1, np.int32)
    try:
        g.closeness(poison=flag)
        raise AssertionError("poisoned call ran")
    except capi.CozoGpuError as e:
        assert e.code == capi.E_KILLED
    step("poison -> EKILLED")


# ---------------------------------------------------------------------------------------------------------------------
PR_GEOMETRIES = [
    {"pagerank__mode": 0},
    {"pagerank__mode": 1, "pagerank__hub_slots": 16384, "pagerank__group_slots": 32768, "pagerank__window": 24576, "pagerank__chunk": 262144},
    {"pagerank__mode": 1, "pagerank__hub_slots": 64, "pagerank__group_slots": 256, "pagerank__window": 512, "pagerank__chunk": 1024},
    {"pagerank__mode": 1, "pagerank__hub_slots": 0, "pagerank__group_slots": 64, "pagerank__window": 64, "pagerank__chunk": 1024},
    {"pagerank__mode": 1, "pagerank__hub_slots": 8, "pagerank__group_slots": 128, "pagerank__window": 301, "pagerank__chunk": 1024},
    # back to "every source in the hub table" on the SAME staged graphs: the blocking tables of the previous geometry
    # are freed and nothing of them may be used (a stale item count once launched the gather pass on freed tables)
    {"pagerank__mode": 1, "pagerank__hub_slots": 16384, "pagerank__group_slots": 32768, "pagerank__window": 24576, "pagerank__chunk": 262144},
]


def scenario_pagerank():
    """pagerank.cu end to end: slot staging, both engines, blocking (re)staging when the geometry options change"""
    from tests.util import rmat_edges
    n, src, dst = rmat_edges(9, 12, 0x5EED0004 + 9)
    o = O.OracleGraph(n, src, dst)
    os_, oit, oerr = o.pagerank(0.85, 0.0, 4, variant="jacobi", n_threads=2)
    ot, otit, _ = o.pagerank(0.85, 1e-3, 30, variant="jacobi", n_threads=2)
    star_n = 700
    ss = np.concatenate([np.arange(1, star_n), np.zeros(star_n - 1)]).astype(np.uint32)
    sd = np.concatenate([np.zeros(star_n - 1), np.arange(1, star_n)]).astype(np.uint32)
    s64 = np.full(star_n, 1.0 / star_n)
    outdeg = np.bincount(ss, minlength=star_n).astype(np.float64)
    for _ in range(4):
        nxt = np.full(star_n, 0.15 / star_n)
        np.add.at(nxt, sd, 0.85 * (s64 / outdeg)[ss])
        s64 = nxt
    ps = np.arange(0, 90, dtype=np.uint32)
    o_path, _, _ = O.OracleGraph(300, ps, ps + 1).pagerank(0.85, 0.0, 5)
    g = capi.Graph(n, src, dst)                     # ONE staged graph: the blocking is rebuilt when the options change
    gstar = capi.Graph(star_n, ss, sd)
    gpath = capi.Graph(300, ps, ps + 1)
    for geo in PR_GEOMETRIES:
        with opts(**geo):
            gs, git, gerr, _ = g.pagerank(0.85, 0.0, 4)
            rel = float(np.max(np.abs(gs - os_) / os_))
            assert git == oit and rel <= 1e-5, (geo, rel)
            assert abs(gerr - oerr) <= 0.02 * oerr + 1e-6
            assert np.array_equal(gs, g.pagerank(0.85, 0.0, 4)[0])          # bit-identical rerun
            gt, gtit, _, _ = g.pagerank(0.85, 1e-3, 30)
            assert gtit == otit and np.max(np.abs(gt - ot) / ot) <= 1e-5    # same stopping iteration
            gst = gstar.pagerank(0.85, 0.0, 4)[0]
            assert np.max(np.abs(gst - s64) / s64) <= 4e-6                  # a row many windows long
            assert np.allclose(gpath.pagerank(0.85, 0.0, 5)[0], o_path, rtol=1e-6)   # ids without edges
        step(f"{geo}: rmat-9 max rel err {rel:.2e}, tol stop at {gtit}, star + path graphs ok, launches "
             f"{capi.get_option('pagerank.last_launches')}")
    g0 = capi.Graph(0, np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    s, it, _, _ = g0.pagerank()
    assert s.size == 0 and it == 0
    for geo in (PR_GEOMETRIES[0], PR_GEOMETRIES[2]):
        with opts(**geo):
            e0 = capi.Graph(5, np.zeros(0, np.uint32), np.zeros(0, np.uint32)).pagerank(0.85, 0.0, 2)[0]   # no edges
            assert np.allclose(e0, 0.15 / 5)
            s3, d3 = np.array([0, 1], np.uint32), np.array([1, 2], np.uint32)
            assert np.allclose(capi.Graph(3, s3, d3).pagerank(0.85, 0.0, 3)[0], O.OracleGraph(3, s3, d3).pagerank(0.85, 0.0, 3)[0], rtol=1e-6)
    step("empty graph, no edges, dangling node")


# ---------------------------------------------------------------------------------------------------------------------
def _graph_sets(levels):
    ni, rp, ci, ep = levels
    out = []
    for L in range(len(rp)):
        nodes = np.arange(len(rp[0]) - 1) if L == 0 else ni[L]
        out.append({int(v): set(ci[L][int(rp[L][r]):int(rp[L][r + 1])].tolist()) for r, v in enumerate(nodes)})
    return out, ep


def scenario_hnsw():
    """hnsw_build.cu + hnsw.cu + hnsw_f64.cu + merge.cu end to end through the ctypes binding"""
    rng = np.random.default_rng(1)
    n, dim = (140 if LONG else 100), 32
    X = rng.random((n, dim), dtype=np.float32)
    Q = rng.random((9, dim), dtype=np.float32)
    for metric in ((capi.L2, capi.COSINE) if LONG else (capi.L2,)):
        g = capi.HnswIndex.build(X - (0.5 if metric else 0), metric=metric, m=6, ef_construction=24, keep_pruned_connections=bool(metric))
        Xm = X - (0.5 if metric else 0)
        ni, rp, ci, ep = g.export_levels()
        ix = O.OracleHnsw.from_levels(Xm, O.HnswLevels(ni, rp, ci, ep), metric=metric)
        oi, od, oc, ost = ix.search(Q - (0.5 if metric else 0), 5, 20, n_threads=2)
        for mode in (1, 0, 2):
            with opts(hnsw__mode=mode):
                gi, gd, gc, st = g.search(Q - (0.5 if metric else 0), 5, 20)
            assert np.array_equal(gi, oi) and np.allclose(gd, od, rtol=1e-5, atol=1e-6) and np.array_equal(gc, oc)
            assert st.dist_evals == int(ost[:, 0].sum()) and st.nodes_expanded == int(ost[:, 1].sum())
        step(f"metric {metric}: build ({n} x {dim}), search modes TMA ring / ld.global.nc / cooperative: ids, distances, "
             f"traversal counters == oracle")
        if metric:
            continue
        # radius, k > ef, filter mask (trim after the filter), in all modes
        full_i, full_d, full_c, _ = ix.search(Q, 20, 20, n_threads=2)
        r = float(np.median(od[:, 3])) * (1 + 1e-4)     # not ON a candidate distance: device and oracle differ in the last ulp
        assert np.min(np.abs(full_d / r - 1)) > 1e-5
        keep = rng.random(n) < 0.4
        for mode in (1, 0, 2):
            with opts(hnsw__mode=mode):
                ri, rd, rc, _ = g.search(Q, 5, 20, radius=r)
                fi, fd, fc, _ = g.search(Q, 5, 20, radius=r, row_pass=keep)
                ki, kd, kc, _ = g.search(Q, 30, 20)
            for q in range(len(Q)):
                cand = [(int(i), float(d)) for i, d in zip(full_i[q, :full_c[q]], full_d[q, :full_c[q]])]
                assert list(ri[q, :rc[q]]) == [i for i, d in cand if d <= r][:5]
                assert list(fi[q, :fc[q]]) == [i for i, d in cand if d <= r and keep[i]][:5]
                assert kc[q] == full_c[q] and list(ki[q, :kc[q]]) == [i for i, _ in cand]
        step("radius, k > ef, filter mask applied before the trim (all modes)")
        # stage the exported graph again (the Rust glue's path) and an F64 copy
        g2 = capi.HnswIndex.stage(X, ni, rp, ci, ep, m_max0=12, m_max=6)
        assert np.array_equal(g2.search(Q, 5, 20)[0], oi)
        X64 = rng.random((n, dim)) - 0.5
        Q64 = rng.random((9, dim)) - 0.5
        for m64 in (capi.L2, capi.COSINE, capi.IP):
            g64 = capi.HnswIndex.stage(X64, ni, rp, ci, ep, metric=m64, m_max0=12, m_max=6)
            ix64 = O.OracleHnsw.from_levels(X64.astype(np.float32), O.HnswLevels(ni, rp, ci, ep), metric=m64)
            o64i, o64d, o64c, _ = ix64.search_f64(X64, Q64, 5, 20, n_threads=2)
            g64i, g64d, g64c, _ = g64.search_f64(Q64, 5, 20, row_pass=None)
            assert np.array_equal(g64i, o64i) and np.allclose(g64d, o64d, rtol=1e-12, atol=1e-14) and np.array_equal(g64c, o64c)
            f64i, _, f64c, _ = g64.search_f64(Q64, 5, 20, row_pass=keep)
            fo_i, _, fo_c, _ = ix64.search_f64(X64, Q64, 20, 20, n_threads=2)
            for q in range(len(Q64)):
                assert list(f64i[q, :f64c[q]]) == [int(i) for i in fo_i[q, :fo_c[q]] if keep[i]][:5]
            try:
                g64.search(Q, 5, 20)
                raise AssertionError("f32 search on an F64 handle")
            except capi.CozoGpuError as e:
                assert e.code == capi.E_INVAL
        step("re-staged graph; F64 index L2 / Cosine / IP == oracle f64 path, filtered form too")
        # device-pointer forms: search_dev into caller buffers, top-k merge of two "shards"
        B, k = len(Q), 5
        ids = np.zeros((2, B, k), np.uint32)
        dist = np.zeros((2, B, k), np.float32)
        cnt = np.zeros(B, np.uint32)
        Qc = np.ascontiguousarray(Q)
        g.search_dev(Qc.ctypes.data, B, k, 20, ids[0].ctypes.data, dist[0].ctypes.data, cnt.ctypes.data)
        g2.search_dev(Qc.ctypes.data, B, k, 20, ids[1].ctypes.data, dist[1].ctypes.data, None)
        assert np.array_equal(ids[0], oi) and np.array_equal(ids[1], oi)
        offs = np.array([0, 1000], np.uint64)
        mi, md = np.zeros((B, k), np.uint64), np.zeros((B, k), np.float32)
        capi.topk_merge_dev(dist.ctypes.data, ids.ctypes.data, 2, B, k, offs.ctypes.data, mi.ctypes.data, md.ctypes.data)
        for q in range(B):
            allp = sorted([(float(dist[s, q, j]), s, int(ids[s, q, j]) + int(offs[s])) for s in range(2) for j in range(k)])
            assert [p[2] for p in allp[:k]] == mi[q].tolist()
        step("search_dev + topk_merge_dev")


def scenario_hnsw_maintenance():
    """insert / update / remove / exports on built and staged handles, insert into an empty staged handle"""
    rng = np.random.default_rng(3)
    n, dim = 90, 16
    X = rng.random((n + 30, dim), dtype=np.float32)
    Q = rng.random((8, dim), dtype=np.float32)
    g = capi.HnswIndex.build(X[:n], m=5, ef_construction=20)
    first = g.insert(X[n:n + 20])
    assert first == n
    g.update(np.array([3, 4, 50, n + 19], np.uint32), X[n + 20:n + 24])
    g.remove(np.arange(10, 25, dtype=np.uint32))
    live = g.export_live()
    assert live.sum() == n + 20 - 15
    Xc = X[:n + 20].copy()
    Xc[[3, 4, 50, n + 19]] = X[n + 20:n + 24]
    ni, rp, ci, ep = g.export_levels()
    sets, _ = _graph_sets((ni, rp, ci, ep))
    for L, rows in enumerate(sets):
        for v, nb in rows.items():
            assert not (nb & set(range(10, 25))), "edge to a removed node"
            assert len(nb) <= (10 if L == 0 else 5)
    ix = O.OracleHnsw.from_levels(Xc, O.HnswLevels(ni, rp, ci, ep))
    gi, gd, gc, _ = g.search(Q, 5, 20)
    oi, od, oc, _ = ix.search(Q, 5, 20, n_threads=2)
    assert np.array_equal(gi, oi) and np.allclose(gd, od, rtol=1e-5)
    bi, _ = O.bruteforce_knn(np.where(live[:, None] > 0, Xc, 1e3), Q, 5, n_threads=2)
    rec = np.mean([len(set(a) & set(b)) / 5 for a, b in zip(gi, bi)])
    assert rec >= 0.9, rec
    step(f"insert / update / remove: graph consistent, search == oracle on the exported graph, recall vs brute force {rec:.2f}")
    dists = g.export_dists()
    for L in range(len(ci)):
        nodes = np.arange(len(rp[0]) - 1) if L == 0 else ni[L]
        for r, v in enumerate(nodes):
            for e in range(int(rp[L][r]), int(rp[L][r + 1])):
                d = float(np.sum((Xc[v] - Xc[ci[L][e]]) ** 2, dtype=np.float64))
                assert abs(dists[L][e] - d) <= 1e-5 * max(d, 1e-6)
    step("stored edge distances == recomputed")
    # a handle staged from another handle's export grows by insert (build state derived on demand)
    gs = capi.HnswIndex.stage(Xc, ni, rp, ci, ep, m_max0=10, m_max=5)
    f2 = gs.insert(X[n + 24:n + 30], ef_construction=20)
    assert f2 == n + 20 and (gs.search(Q, 5, 20)[2] == 5).all()
    # the canary-only index: staged with zero rows, then filled
    e = capi.HnswIndex.stage(np.zeros((0, dim), np.float32), [np.zeros(0, np.uint32)], [np.zeros(1, np.uint32)], [np.zeros(0, np.uint32)],
                             0xFFFFFFFF, m_max0=10, m_max=5)
    assert e.insert(X[:40], ef_construction=20) == 0
    ei, _, ec, _ = e.search(X[:6], 1, 20)
    assert (ec == 1).all() and ei[:, 0].tolist() == list(range(6))
    step("insert into a staged handle and into an empty staged handle")


def scenario_builder_fidelity():
    """hnsw_insert_range in fidelity mode (max_batch = 1) and with extend_candidates == the oracle's sequential builder"""
    dim, m, efc = 16, 4, 12
    cases = ((70, False, False), (60, True, False), (48, False, True), (44, True, True))
    for n, keep, extend in (cases if LONG else (cases[0], cases[2])):
        X = np.random.default_rng(n).random((n, dim), dtype=np.float32)
        g = capi.HnswIndex.build(X, m=m, ef_construction=efc, keep_pruned_connections=keep, level_seed=77, max_batch=1, extend_candidates=extend)
        ni, rp, ci, ep = g.export_levels()
        level = np.zeros(n, np.int64)
        for L in range(1, len(rp)):
            level[ni[L]] = L
        ix = O.OracleHnsw.new(n, dim, m=m, ef_construction=efc, keep_pruned_connections=keep, extend_candidates=extend)
        for i in range(n):
            ix.insert(i, X[i], forced_level=-int(level[i]))
        lv = ix.levels()
        assert lv.entry == ep and lv.n_levels == len(rp)
        dsets, _ = _graph_sets((ni, rp, ci, ep))
        osets, _ = _graph_sets((lv.node_ids, lv.row_ptr, lv.col_idx, lv.entry))
        same = sum(dsets[L][v] == osets[L][v] for L in range(len(dsets)) for v in dsets[L])
        total = sum(len(d) for d in dsets)
        assert all(sorted(a) == sorted(b) for a, b in zip(dsets, osets))
        assert same >= total - max(1, total // 100), (same, total)
        step(f"n={n} keep_pruned={keep} extend_candidates={extend}: {same} / {total} rows identical to the oracle builder's")


def scenario_sharded():
    """sharded.cu end to end with the ranks of a communicator as THREADS of this process (tests/emu/fake_nccl.cpp, CUDA IPC
    handles = pointers): NCCL-form and fused peer-store exchange, tiling with double-buffered sets, host and device forms,
    broadcast from a root, growth of the exchange buffers (close peers -> rendezvous -> reallocate), radius"""
    import threading

    from cozo_b200.sharded import merge_lists
    os.environ["COZO_GPU_NCCL_LIB"] = os.path.join(os.path.dirname(capi.LIB_PATH), "libfake_nccl.so")
    dim, m, k, ef, B = 24, 6, 5, 20, 37
    Q = np.random.default_rng(7).random((B, dim), dtype=np.float32)
    Qbig = np.random.default_rng(8).random((3 * B, dim), dtype=np.float32)
    for world in ((1, 2, 3, 8) if (os.cpu_count() or 1) >= 8 else (1, 2, 3)):     # 8 = the driver's scaling run
        shards, rows = [], []
        for r in range(world):
            n = 110 + 25 * r                                           # ragged shards
            X = np.random.default_rng(100 + r).random((n, dim), dtype=np.float32)
            g = capi.HnswIndex.build(X, m=m, ef_construction=24, level_seed=5 + r)
            ix = O.OracleHnsw.from_levels(X, O.HnswLevels(*g.export_levels()))
            shards.append((g, ix))
            rows.append(n)
        offsets = np.cumsum([0] + rows[:-1])

        def expected(Qx, kk, radius=None):
            li = [g.search(Qx, kk, ef, radius=radius) for g, _ in shards]
            oi = [ix.search(Qx, kk, ef, n_threads=2) for _, ix in shards]
            e = merge_lists(np.stack([x[0] for x in li]), np.stack([x[1] for x in li]), offsets, kk)
            o = merge_lists(np.stack([x[0] for x in oi]), np.stack([x[1] for x in oi]), offsets, kk)
            return e, o, sum(int(x[3].dist_evals) for x in li)
        (exp_i, exp_d), (orc_i, orc_d), _ = expected(Q, k)
        assert np.array_equal(exp_i, orc_i) and np.allclose(exp_d, orc_d, rtol=1e-5)      # same sharding, same answer
        (big_i, big_d), _, _ = expected(Qbig, k)
        rad = float(np.median(exp_d[:, 3])) * (1 + 1e-4)
        (rad_i, rad_d), _, _ = expected(Q, k, radius=rad)
        seen = set()
        for exchange in (1, 0):
            for tile in (65536, 10):
                capi.set_option("shard.exchange", exchange)
                capi.set_option("shard.tile", tile)
                uid = capi.ShardGroup.unique_id()
                errors, infos = [], [None] * world

                def rank_main(r):
                    try:
                        grp = capi.ShardGroup(uid, r, world)
                        off, total = grp.attach(shards[r][0])
                        assert off == int(offsets[r]) and total == sum(rows)
                        for root in (-1, 0, world - 1):
                            ids, dd, cnt, st = grp.search(Q if root in (-1, r) else None, k, ef, root=root, B=B)
                            assert np.array_equal(ids, exp_i) and np.array_equal(dd, exp_d), (exchange, tile, root, r)
                            assert np.all(cnt == k)
                        infos[r] = grp.info()
                        # device form, back to back (double-buffered sets), then a BIGGER batch: the exchange buffers grow
                        oi_d, od_d = np.zeros((B, k), np.uint64), np.zeros((B, k), np.float32)
                        for _ in range(3):
                            grp.search_dev(Q.ctypes.data, B, k, ef, oi_d.ctypes.data, od_d.ctypes.data, None, None)
                        assert np.array_equal(oi_d, exp_i) and np.array_equal(od_d, exp_d)
                        ids, dd, _, _ = grp.search(Qbig, k, ef)
                        assert np.array_equal(ids, big_i) and np.array_equal(dd, big_d)
                        ids, dd, cnt, _ = grp.search(Q, k, ef, radius=rad)
                        assert np.array_equal(ids, rad_i) and np.array_equal(cnt, (rad_d <= np.float32(rad)).sum(1))
                        grp.close()
                    except BaseException as e:           # noqa: BLE001
                        import traceback
                        errors.append(f"rank {r}: {e!r}\n{traceback.format_exc()}")
                th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
                for t in th:
                    t.start()
                for t in th:
                    t.join(600)
                assert not any(t.is_alive() for t in th), "a rank hangs"
                assert not errors, "\n".join(errors)
                assert len({i["exchange"] for i in infos}) == 1
                seen.add(infos[0]["exchange"])
        assert seen == {"fused", "nccl"}, seen
        step(f"world {world}: host / device forms, roots, tiles of 65536 and 10, growth, radius == numpy merge of the per-shard "
             f"lists == oracle under the same sharding; exchanges {sorted(seen)}")
    capi.set_option("shard.exchange", 1)
    capi.set_option("shard.tile", 65536)


FUZZ_N = int(os.environ.get("COZO_EMU_FUZZ", 40))


def scenario_fuzz_hnsw():
    """differential fuzzing of build + search against the oracle on the exported graph: random sizes (1..129 rows), dims
    (1..130, padded and unpadded), metrics, m, ef_construction, k, ef, batch, search mode, duplicates (distance ties),
    radius, filter mask; the expected list is the oracle's ef candidates put through the reference's post-processing"""
    N, seed0 = FUZZ_N, 1000
    fails = 0
    for case in range(N):
        rng = np.random.default_rng(seed0 + case)
        n = int(rng.integers(1, 130)); dim = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 16, 17, 31, 32, 33, 48, 100, 130]))
        metric = int(rng.integers(0, 3)); m = int(rng.integers(2, 9)); efc = int(rng.integers(2, 40))
        k = int(rng.integers(1, 16)); ef = int(rng.integers(1, 40)); B = int(rng.integers(1, 12))
        keep = bool(rng.integers(0, 2)); mode = int(rng.choice([-1, 0, 1, 2]))
        X = (rng.random((n, dim), dtype=np.float32) - 0.5)
        if rng.random() < 0.2 and n > 4:       # duplicates -> distance ties
            X[n // 2:] = X[: n - n // 2]
        if metric == 1: X += 0.01            # no zero vectors for cosine
        Q = (rng.random((B, dim), dtype=np.float32) - 0.5)
        desc = dict(n=n, dim=dim, metric=metric, m=m, efc=efc, k=k, ef=ef, B=B, keep=keep, mode=mode, seed=seed0 + case)
        try:
            g = capi.HnswIndex.build(X, metric=metric, m=m, ef_construction=efc, keep_pruned_connections=keep, level_seed=case)
            lv = g.export_levels()
            ix = O.OracleHnsw.from_levels(X, O.HnswLevels(*lv), metric=metric)
            use_r = rng.random() < 0.3
            oi, od, oc, ost = ix.search(Q, k, ef, n_threads=1)
            radius = None
            if use_r and oc.min() > 0:
                radius = float(np.median(od[:, 0])) * 1.5 + 1e-3 if np.median(od[:, 0]) > 0 else None
            use_f = rng.random() < 0.3
            row_pass = (rng.random(n) < 0.5) if use_f else None
            capi.set_option("hnsw.mode", mode)
            gi, gd, gc, st = g.search(Q, k, ef, radius=radius, row_pass=row_pass)
            capi.set_option("hnsw.mode", -1)
            fi, fd, fc, _ = ix.search(Q, ef, ef, n_threads=1)     # all candidates, then the reference's post-processing
            for q in range(B):
                cand = [(int(i), float(d)) for i, d in zip(fi[q, :fc[q]], fd[q, :fc[q]])]
                if radius is not None: cand = [c for c in cand if c[1] <= radius]
                if row_pass is not None: cand = [c for c in cand if row_pass[c[0]]]
                cand = cand[:k]
                got = [(int(i), float(d)) for i, d in zip(gi[q, :gc[q]], gd[q, :gc[q]])]
                ok = len(got) == len(cand) and all(abs(a[1] - b[1]) <= 1e-5 * max(1, abs(b[1])) + 1e-6 for a, b in zip(got, cand))
                # ids may swap inside an exact distance tie; compare as multisets of (rounded dist) and sets when no ties
                if ok and [a[0] for a in got] != [b[0] for b in cand]:
                    ds = [round(b[1], 6) for b in cand]
                    ok = len(set(ds)) < len(ds) or (radius is not None and any(abs(b[1] / radius - 1) < 1e-5 for b in cand))
                if not ok:
                    fails += 1
                    print('MISMATCH', desc, 'q', q, 'got', got[:6], 'exp', cand[:6], flush=True)
                    break
        except Exception as e:
            fails += 1
            print('ERROR', desc, repr(e), flush=True)
    
    assert fails == 0, fails
    step(f"{N} random index / query configurations: 0 mismatches")


def scenario_fuzz_graph():
    """differential fuzzing of the graph rules: random multigraphs (self loops, duplicate edges, isolated nodes, zero
    weights), SSSP in the three frontier forms, closeness, betweenness, PageRank in both engines with random blocking
    geometries, clustering — all against the oracle"""
    N, seed0, fails = FUZZ_N, 5000, 0
    for case in range(N):
        rng = np.random.default_rng(seed0 + case)
        n = int(rng.integers(1, 90)); m = int(rng.integers(0, 6 * n + 1))
        src = rng.integers(0, n, m).astype(np.uint32); dst = rng.integers(0, n, m).astype(np.uint32)
        kind = int(rng.integers(0, 3))
        w = [(rng.integers(0, 8, m) / 4).astype(np.float32), (rng.random(m) * 5).astype(np.float32), (rng.integers(1, 64, m) / 8).astype(np.float32)][kind]
        desc = dict(n=n, m=m, kind=kind, seed=seed0 + case)
        try:
            g = capi.Graph(n, src, dst, w); o = O.OracleGraph(n, src, dst, w)
            srcs = rng.integers(0, n, int(rng.integers(1, 6))).astype(np.uint32)
            od, _ = o.sssp(srcs)
            for opt in (None, "sssp.frontier", "sssp.wide"):
                if opt: capi.set_option(opt, 1)
                gd, gp, _ = g.sssp(srcs)
                if opt: capi.set_option(opt, 0)
                if not np.array_equal(gd, od):
                    fails += 1; print('SSSP MISMATCH', desc, opt, flush=True)
            oc = o.closeness(); gc, _ = g.closeness()
            fin = np.isfinite(oc)
            if not (np.array_equal(np.isfinite(gc), fin) and np.allclose(gc[fin], oc[fin], rtol=1e-5)):
                fails += 1; print('CLOSENESS MISMATCH', desc, flush=True)
            if kind == 2:     # strictly positive weights: betweenness well defined
                ob = o.betweenness(); gb, _ = g.betweenness()
                if not np.allclose(gb, ob, rtol=1e-4, atol=1e-6):
                    fails += 1; print('BETWEENNESS MISMATCH', desc, np.max(np.abs(gb - ob)), flush=True)
            gu = capi.Graph(n, src, dst); ou = O.OracleGraph(n, src, dst)
            it = int(rng.integers(1, 8))
            os_, oit, _ = ou.pagerank(0.85, 0.0, it)
            for mode in (0, 1):
                capi.set_option("pagerank.mode", mode)
                if mode:
                    capi.set_option("pagerank.hub_slots", int(rng.choice([0, 4, 8, 64])))
                    capi.set_option("pagerank.group_slots", int(rng.choice([64, 128, 256])))
                    capi.set_option("pagerank.window", int(rng.choice([64, 97, 512])))
                    capi.set_option("pagerank.chunk", 1024)
                gs, git, _, _ = gu.pagerank(0.85, 0.0, it)
                if git != oit or (n and np.max(np.abs(gs - os_) / os_) > 1e-5):
                    fails += 1; print('PAGERANK MISMATCH', desc, mode, flush=True)
            capi.set_option("pagerank.mode", 0)
            ms, md = np.concatenate([src, dst]), np.concatenate([dst, src])
            a = capi.Graph(n, ms, md).clustering(); b = O.OracleGraph(n, ms, md).clustering()
            if not (np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])):
                fails += 1; print('CLUSTERING MISMATCH', desc, flush=True)
        except Exception as e:
            fails += 1; print('ERROR', desc, repr(e), flush=True)
    
    assert fails == 0, fails
    step(f"{N} random graphs: 0 mismatches")


def scenario_fuzz_maintenance():
    """random sequences of insert / remove / update on a built index: live flags, row invariants (degree caps, no duplicate
    edge, no edge to a removed node), a live entry point, and search == oracle on the exported graph after every step"""
    N, fails = max(1, FUZZ_N // 2), 0
    for case in range(N):
        rng = np.random.default_rng(9000 + case)
        dim = int(rng.choice([3, 8, 17, 32])); m = int(rng.integers(2, 7)); efc = int(rng.integers(4, 30))
        n0 = int(rng.integers(1, 60)); metric = int(rng.integers(0, 3))
        pool = rng.random((400, dim), dtype=np.float32) - 0.5 + (0.01 if metric == 1 else 0)
        X = pool[:n0].copy(); nxt = n0
        desc = dict(case=case, dim=dim, m=m, efc=efc, n0=n0, metric=metric)
        try:
            g = capi.HnswIndex.build(X, metric=metric, m=m, ef_construction=efc, level_seed=case)
            live = np.ones(n0, bool)
            for op in range(int(rng.integers(2, 7))):
                kind = rng.integers(0, 3); n = len(X)
                if kind == 0:
                    c = int(rng.integers(1, 25)); first = g.insert(pool[nxt:nxt + c]); assert first == n
                    X = np.concatenate([X, pool[nxt:nxt + c]]); live = np.concatenate([live, np.ones(c, bool)]); nxt += c
                elif kind == 1 and live.sum() > 2:
                    ids = rng.choice(np.flatnonzero(live), size=int(rng.integers(1, max(2, live.sum() // 3))), replace=False).astype(np.uint32)
                    g.remove(ids); live[ids] = False
                else:
                    ids = rng.choice(n, size=int(rng.integers(1, min(n, 6) + 1)), replace=False).astype(np.uint32)
                    g.update(ids, pool[nxt:nxt + len(ids)]); X[ids] = pool[nxt:nxt + len(ids)]; live[ids] = True; nxt += len(ids)
                assert np.array_equal(g.export_live().astype(bool), live), 'live flags'
                ni, rp, ci, ep = g.export_levels()
                for L in range(len(rp)):
                    nodes = np.arange(len(X)) if L == 0 else ni[L]
                    for r, v in enumerate(nodes):
                        nb = ci[L][int(rp[L][r]):int(rp[L][r + 1])]
                        assert len(nb) <= (2 * m if L == 0 else m), 'degree'
                        assert len(set(nb.tolist())) == len(nb), 'duplicate edge'
                        assert live[nb].all() if len(nb) else True, 'edge to a removed node'
                        if not live[v]: assert len(nb) == 0 or True
                if live.any():
                    assert ep != 0xFFFFFFFF and live[ep], 'entry point'
                    Q = rng.random((5, dim), dtype=np.float32) - 0.5
                    ix = O.OracleHnsw.from_levels(X, O.HnswLevels(ni, rp, ci, ep), metric=metric)
                    gi, gd, gc, _ = g.search(Q, 4, 16); oi, od, oc, _ = ix.search(Q, 4, 16, n_threads=1)
                    assert np.array_equal(gc, oc) and np.allclose(gd, od, rtol=1e-5, atol=1e-6), 'search vs oracle on the exported graph'
                    assert all(live[i] for q in range(5) for i in gi[q, :gc[q]]), 'removed id returned'
        except Exception as e:
            fails += 1; print('FAIL', desc, repr(e), flush=True)
    
    assert fails == 0, fails
    step(f"{N} random maintenance sequences: 0 failures")


def scenario_sanitize_workload():
    """tools/sanitize.py — the tiny end-to-end exercise of every kernel written for compute-sanitizer — on the emulated library
    (with COZO_EMU_SANITIZE=1 that is AddressSanitizer + UBSan over host code and kernels)"""
    import runpy
    os.environ["COZO_GPU_NCCL_LIB"] = os.path.join(os.path.dirname(capi.LIB_PATH), "libfake_nccl.so")
    runpy.run_path(os.path.join(ROOT, "tools", "sanitize.py"), run_name="__main__")
    step("tools/sanitize.py ran to the end")


def scenario_fuzz_sequences():
    """random API sequences on one handle — built, or staged from an export like the glue does — insert / remove / update with
    a search of random shape (k, ef, batch, mode, filter) after every step: results and traversal counters == oracle on the
    exported graph, no removed id ever returned (exercises workspace regrowth and the lazily derived build state)"""
    N, fails = max(1, FUZZ_N // 2), 0
    for case in range(N):
        rng = np.random.default_rng(31000 + case)
        dim = int(rng.choice([4, 9, 32, 130])); m = int(rng.integers(2, 9)); efc = int(rng.integers(4, 40)); metric = int(rng.integers(0, 3))
        pool = rng.random((600, dim), dtype=np.float32) - 0.5 + (0.01 if metric == 1 else 0)
        n0 = int(rng.integers(1, 80)); X = pool[:n0].copy(); nxt = n0; live = np.ones(n0, bool)
        desc = dict(case=case, dim=dim, m=m, efc=efc, metric=metric, n0=n0)
        try:
            staged = rng.random() < 0.4
            g = capi.HnswIndex.build(X, metric=metric, m=m, ef_construction=efc, level_seed=case)
            if staged:      # continue on a handle staged from the export (the glue's path)
                ni, rp, ci, ep = g.export_levels()
                g = capi.HnswIndex.stage(X, ni, rp, ci, ep, metric=metric, m_max0=2 * m, m_max=m)
            for op in range(int(rng.integers(3, 10))):
                kind = int(rng.integers(0, 5)); n = len(X)
                if kind == 0:
                    c = int(rng.integers(1, 40)); first = g.insert(pool[nxt:nxt + c], ef_construction=efc); assert first == n
                    X = np.concatenate([X, pool[nxt:nxt + c]]); live = np.concatenate([live, np.ones(c, bool)]); nxt += c
                elif kind == 1 and live.sum() > 2:
                    ids = rng.choice(np.flatnonzero(live), size=int(rng.integers(1, max(2, live.sum() // 2))), replace=False).astype(np.uint32)
                    g.remove(ids); live[ids] = False
                elif kind == 2:
                    ids = rng.choice(n, size=int(rng.integers(1, min(n, 8) + 1)), replace=False).astype(np.uint32)
                    g.update(ids, pool[nxt:nxt + len(ids)], ef_construction=efc); X[ids] = pool[nxt:nxt + len(ids)]; live[ids] = True; nxt += len(ids)
                if not live.any():
                    continue
                # a search with random shape after every step (workspace regrowth, all modes, filter, radius)
                k = int(rng.integers(1, 20)); ef = int(rng.integers(1, 60)); B = int(rng.integers(1, 40)); mode = int(rng.choice([-1, 0, 1, 2]))
                Q = rng.random((B, dim), dtype=np.float32) - 0.5
                ni, rp, ci, ep = g.export_levels()
                ix = O.OracleHnsw.from_levels(X, O.HnswLevels(ni, rp, ci, ep), metric=metric)
                row_pass = (rng.random(len(X)) < 0.6) if rng.random() < 0.4 else None
                capi.set_option("hnsw.mode", mode)
                gi, gd, gc, st = g.search(Q, k, ef, row_pass=row_pass)
                capi.set_option("hnsw.mode", -1)
                fi, fd, fc, ost = ix.search(Q, ef, ef, n_threads=1)
                assert st.dist_evals == int(ost[:, 0].sum()), ('counters', st.dist_evals, int(ost[:, 0].sum()))
                for q in range(B):
                    cand = [(int(i), float(d)) for i, d in zip(fi[q, :fc[q]], fd[q, :fc[q]])]
                    if row_pass is not None: cand = [c for c in cand if row_pass[c[0]]]
                    cand = cand[:k]
                    got = [(int(i), float(d)) for i, d in zip(gi[q, :gc[q]], gd[q, :gc[q]])]
                    assert len(got) == len(cand) and all(abs(a[1] - b[1]) <= 1e-5 * max(1, abs(b[1])) + 1e-6 for a, b in zip(got, cand)), ('search', op, q, got[:4], cand[:4])
                    assert all(live[i] for i, _ in got)
        except Exception as e:
            fails += 1; print('FAIL', desc, repr(e)[:400], flush=True)
    
    assert fails == 0, fails
    step(f"{N} random API sequences: 0 failures")


def scenario_fuzz_sharded():
    """random communicators (1-4 rank threads, shards of 1-89 rows, metrics), exchanges, tiles (1, 3, 7, 65536) and call
    sequences with random batch / k / ef / root: the operator's result == numpy merge of the per-shard lists on every rank"""
    import threading

    from cozo_b200.sharded import merge_lists
    os.environ["COZO_GPU_NCCL_LIB"] = os.path.join(os.path.dirname(capi.LIB_PATH), "libfake_nccl.so")
    N, fails = max(1, FUZZ_N // 4), 0
    for case in range(N):
        rng = np.random.default_rng(4400 + case)
        world = int(rng.integers(1, 5)); dim = int(rng.choice([3, 16, 33])); metric = int(rng.integers(0, 3))
        shards, rows = [], []
        for r in range(world):
            n = int(rng.integers(1, 90))
            X = rng.random((n, dim), dtype=np.float32) - 0.5 + (0.01 if metric == 1 else 0)
            shards.append(capi.HnswIndex.build(X, metric=metric, m=4, ef_construction=16, level_seed=case * 10 + r)); rows.append(n)
        offsets = np.cumsum([0] + rows[:-1])
        exchange = int(rng.integers(0, 2)); tile = int(rng.choice([1, 3, 7, 65536]))
        capi.set_option("shard.exchange", exchange); capi.set_option("shard.tile", tile)
        calls = []
        for _ in range(int(rng.integers(1, 5))):
            B = int(rng.integers(1, 30)); k = int(rng.integers(1, 12)); ef = int(rng.integers(1, 30))
            Q = rng.random((B, dim), dtype=np.float32) - 0.5
            li = [g.search(Q, k, ef) for g in shards]
            calls.append((Q, k, ef, merge_lists(np.stack([x[0] for x in li]), np.stack([x[1] for x in li]), offsets, k), int(rng.integers(-1, world))))
        uid = capi.ShardGroup.unique_id(); errs = []
        def rank_main(r):
            try:
                grp = capi.ShardGroup(uid, r, world); grp.attach(shards[r])
                for Q, k, ef, (ei, ed), root in calls:
                    ids, dd, cnt, _ = grp.search(Q if root in (-1, r) else None, k, ef, root=root, B=len(Q))
                    assert np.array_equal(ids, ei) and np.array_equal(dd, ed), (r, root)
                    assert np.array_equal(cnt, (ei != np.uint64(0xFFFFFFFFFFFFFFFF)).sum(1))
                grp.close()
            except BaseException as e:
                errs.append(repr(e)[:300])
        th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        [t.start() for t in th]; [t.join(300) for t in th]
        if errs or any(t.is_alive() for t in th):
            fails += 1; print('FAIL', dict(case=case, world=world, dim=dim, metric=metric, exchange=exchange, tile=tile, rows=rows), errs[:2], flush=True)
    
    capi.set_option("shard.exchange", 1)
    capi.set_option("shard.tile", 65536)
    assert fails == 0, fails
    step(f"{N} random sharded configurations: 0 failures")


def scenario_misuse():
    """edge cases and misuse of every handle type: each call either works or returns a cozo_gpu error code — never a crash
    (run it with COZO_EMU_SANITIZE=1 for the memory-safety half of that statement)"""
    os.environ["COZO_GPU_NCCL_LIB"] = os.path.join(os.path.dirname(capi.LIB_PATH), "libfake_nccl.so")
    rng = np.random.default_rng(0)
    def expect_err(name, fn, codes=None):
        try:
            fn()
        except capi.CozoGpuError as e:
            assert codes is None or e.code in codes, (name, e.code, str(e))
            print(f"  {name}: error {e.code} ok")
            return
        except (AssertionError, ValueError, TypeError, AttributeError) as e:
            print(f"  {name}: python-level {type(e).__name__} ok"); return
        raise AssertionError(f"{name}: accepted")
    def ok(name, fn):
        r = fn(); print(f"  {name}: ok"); return r
    X = rng.random((50, 8), dtype=np.float32)
    g = capi.HnswIndex.build(X, m=4, ef_construction=10)
    Q = rng.random((3, 8), dtype=np.float32)
    expect_err("k=0", lambda: g.search(Q, 0, 10))
    expect_err("ef=0", lambda: g.search(Q, 3, 0))
    ok("B=0", lambda: g.search(np.zeros((0, 8), np.float32), 3, 10))
    ok("k > n", lambda: g.search(Q, 500, 600))
    expect_err("huge ef (beyond shared memory)", lambda: g.search(Q, 3, 100000), codes=(capi.E_UNSUP,))
    ok("large ef", lambda: g.search(Q, 3, 5000))
    expect_err("remove out of range", lambda: g.remove(np.array([50], np.uint32)))
    ok("remove nothing", lambda: g.remove(np.zeros(0, np.uint32)))
    ok("remove twice", lambda: (g.remove(np.array([3], np.uint32)), g.remove(np.array([3], np.uint32))))
    expect_err("update duplicate ids", lambda: g.update(np.array([1, 1], np.uint32), X[:2]))
    expect_err("update out of range", lambda: g.update(np.array([99], np.uint32), X[:1]))
    ok("insert nothing", lambda: g.insert(np.zeros((0, 8), np.float32)))
    ok("remove everything then search", lambda: (g.remove(np.flatnonzero(g.export_live()).astype(np.uint32)), g.search(Q, 3, 10)))
    ok("insert after removing everything", lambda: (g.insert(X[:5]), g.search(Q, 3, 10)))
    expect_err("m = 1", lambda: capi.HnswIndex.build(X, m=1, ef_construction=10))
    expect_err("m = 65", lambda: capi.HnswIndex.build(X, m=65, ef_construction=10))
    expect_err("efc = 0", lambda: capi.HnswIndex.build(X, m=4, ef_construction=0))
    expect_err("metric 7", lambda: capi.HnswIndex.build(X, metric=7, m=4, ef_construction=10))
    ok("one vector", lambda: capi.HnswIndex.build(X[:1], m=4, ef_construction=10).search(Q, 3, 10))
    ni, rp, ci, ep = capi.HnswIndex.build(X, m=4, ef_construction=10).export_levels()
    bad = [c.copy() for c in ci]; bad[0][0] = 4000
    expect_err("stage: neighbour out of range", lambda: capi.HnswIndex.stage(X, ni, rp, bad, ep, m_max0=8, m_max=4))
    badrp = [r.copy() for r in rp]; badrp[0][3] = badrp[0][2] - 1 if badrp[0][2] > 0 else 7
    expect_err("stage: row_ptr not monotone", lambda: capi.HnswIndex.stage(X, ni, badrp, ci, ep, m_max0=8, m_max=4))
    expect_err("stage: entry out of range", lambda: capi.HnswIndex.stage(X, ni, rp, ci, 999, m_max0=8, m_max=4))
    g64 = capi.HnswIndex.stage(X.astype(np.float64), ni, rp, ci, ep, m_max0=8, m_max=4)
    expect_err("f32 search on F64", lambda: g64.search(Q, 3, 10))
    expect_err("insert on F64", lambda: g64.insert(X[:2]))
    gf = capi.HnswIndex.stage(X, ni, rp, ci, ep, m_max0=8, m_max=4)
    expect_err("f64 search on F32", lambda: gf.search_f64(Q.astype(np.float64), 3, 10))
    # graphs
    e = np.zeros(0, np.uint32)
    g0 = capi.Graph(0, e, e)
    ok("n=0 pagerank", g0.pagerank); ok("n=0 closeness", g0.closeness); ok("n=0 betweenness", g0.betweenness); ok("n=0 clustering", g0.clustering)
    ok("n=0 sssp no sources", lambda: g0.sssp(e))
    g1 = capi.Graph(5, np.array([0, 1], np.uint32), np.array([1, 2], np.uint32), np.array([1, 2], np.float32))
    expect_err("sssp source out of range", lambda: g1.sssp(np.array([5], np.uint32)))
    expect_err("pagerank iterations 0", lambda: g1.pagerank(0.85, 1e-4, 0))
    expect_err("paths max_len 0", lambda: g1.sssp_paths(np.array([0], np.uint32), np.array([2], np.uint32), max_len=0))
    expect_err("paths goal out of range", lambda: g1.sssp_paths(np.array([0], np.uint32), np.array([9], np.uint32)))
    ok("paths unreachable", lambda: g1.sssp_paths(np.array([2], np.uint32), np.array([0], np.uint32)))
    ok("self loops only", lambda: capi.Graph(3, np.array([0, 1, 2], np.uint32), np.array([0, 1, 2], np.uint32)).betweenness())
    for mode in (0, 1):
        capi.set_option("pagerank.mode", mode)
        ok(f"pagerank mode {mode} self loops", lambda: capi.Graph(3, np.array([0, 1, 2], np.uint32), np.array([0, 1, 2], np.uint32)).pagerank(0.85, 0.0, 3))
        ok(f"pagerank mode {mode} one node", lambda: capi.Graph(1, e, e).pagerank(0.85, 0.0, 3))
    capi.set_option("pagerank.mode", 0)
    # shards
    expect_err("shards rank out of range", lambda: capi.ShardGroup(capi.ShardGroup.unique_id(), 3, 2))
    expect_err("shards world 17", lambda: capi.ShardGroup(capi.ShardGroup.unique_id(), 0, 17))
    grp = capi.ShardGroup(capi.ShardGroup.unique_id(), 0, 1)
    expect_err("sharded search before attach", lambda: grp.search(Q, 3, 10))
    grp.attach(gf)
    expect_err("sharded k=0", lambda: grp.search(Q, 0, 10))
    ok("sharded B=0", lambda: grp.search(np.zeros((0, 8), np.float32), 3, 10))
    ok("sharded k > rows", lambda: grp.search(Q, 200, 300))
    grp.close()
    step("every misuse answered with an error code, every edge case served")


def scenario_fuzz_csr():
    """staging of ARBITRARY layered CSRs (not built by this library: degrees above m_max0, empty rows, sparse upper layers,
    unsorted rows): search results and traversal counters == oracle on the same structure, export returns the same rows in
    key order"""
    N, fails = FUZZ_N, 0
    for case in range(N):
        rng = np.random.default_rng(88000 + case)
        n = int(rng.integers(1, 100)); dim = int(rng.choice([2, 7, 16, 40])); metric = int(rng.integers(0, 3))
        X = rng.random((n, dim), dtype=np.float32) - 0.5 + (0.01 if metric == 1 else 0)
        nl = int(rng.integers(1, 5))
        node_ids, row_ptr, col_idx = [], [], []
        members = np.arange(n)
        for L in range(nl):
            if L > 0:
                keep = rng.random(len(members)) < 0.4
                if not keep.any(): keep[rng.integers(0, len(members))] = True
                members = members[keep]
            rp, ci = [0], []
            for v in members:
                others = members[members != v]
                deg = int(rng.integers(0, min(len(others), 45) + 1)) if rng.random() < 0.9 else 0
                nb = rng.choice(others, size=deg, replace=False) if deg else np.zeros(0, np.int64)
                ci += [int(x) for x in nb]; rp.append(len(ci))
            node_ids.append(members.astype(np.uint32) if L > 0 else np.arange(n, dtype=np.uint32))
            row_ptr.append(np.array(rp, np.uint32)); col_idx.append(np.array(ci, np.uint32))
        entry = int(members.min())
        desc = dict(case=case, n=n, dim=dim, metric=metric, nl=nl)
        try:
            g = capi.HnswIndex.stage(X, node_ids, row_ptr, col_idx, entry, metric=metric, m_max0=8, m_max=4)
            ix = O.OracleHnsw.from_levels(X, O.HnswLevels(node_ids, row_ptr, col_idx, entry), metric=metric)
            B = int(rng.integers(1, 10)); k = int(rng.integers(1, 12)); ef = int(rng.integers(1, 40)); mode = int(rng.choice([-1, 0, 1, 2]))
            Q = rng.random((B, dim), dtype=np.float32) - 0.5
            capi.set_option("hnsw.mode", mode)
            gi, gd, gc, st = g.search(Q, k, ef)
            capi.set_option("hnsw.mode", -1)
            oi, od, oc, ost = ix.search(Q, k, ef, n_threads=1)
            assert np.array_equal(gc, oc), ('count', gc, oc)
            assert np.allclose(gd, od, rtol=1e-5, atol=1e-6, equal_nan=True), 'dist'
            assert st.dist_evals == int(ost[:, 0].sum()), ('evals', st.dist_evals, int(ost[:, 0].sum()))
            ni2, rp2, ci2, ep2 = g.export_levels()
            for L in range(nl):      # rows come back in key order (ascending neighbour id), as the index relation stores them
                assert np.array_equal(rp2[L], row_ptr[L]), 'export row_ptr'
                for r in range(len(row_ptr[L]) - 1):
                    a, b = int(row_ptr[L][r]), int(row_ptr[L][r + 1])
                    assert np.array_equal(ci2[L][a:b], np.sort(col_idx[L][a:b])), 'export row'
            assert ep2 == entry
        except Exception as e:
            fails += 1; print('FAIL', desc, repr(e)[:300], flush=True)
    
    assert fails == 0, fails
    step(f"{N} random layered CSRs: 0 failures")


SCENARIOS = {"graph": scenario_graph, "pagerank": scenario_pagerank, "hnsw": scenario_hnsw,
             "hnsw_maintenance": scenario_hnsw_maintenance, "builder_fidelity": scenario_builder_fidelity,
             "sharded": scenario_sharded, "fuzz_hnsw": scenario_fuzz_hnsw, "fuzz_graph": scenario_fuzz_graph,
             "fuzz_maintenance": scenario_fuzz_maintenance, "sanitize_workload": scenario_sanitize_workload,
             "fuzz_sequences": scenario_fuzz_sequences, "fuzz_sharded": scenario_fuzz_sharded, "misuse": scenario_misuse, "fuzz_csr": scenario_fuzz_csr}

if __name__ == "__main__":
    capi.init(0)
    SCENARIOS[sys.argv[2]]()
    print("EMU_OK")
