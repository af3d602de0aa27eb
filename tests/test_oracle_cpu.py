"""CPU tests: the oracle against independent implementations and known answers.
PARITY UNPINNED: the reference has no golden vectors for this path;
these tests pin the oracle to closed-form values and to scipy / networkx / brute force."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import uniform_vectors, recall, rmat_edges, SEED_DATA, SEED_QUERY, SEED_LEVEL


def test_distance_known_answers():
    # sanity identities the reference asserts informally (runtime/tests.rs:693-694)
    v = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9], np.float32)
    assert O.vec_dist(O.L2, v, v) == 0.0
    assert abs(O.vec_dist(O.COSINE, v, v)) < 1e-7
    n = v / np.sqrt(np.float32(np.dot(v, v)))
    assert abs(O.vec_dist(O.IP, n, n)) < 1e-6
    a = np.array([1, 0, 0], np.float32)
    b = np.array([0, 1, 0], np.float32)
    assert O.vec_dist(O.L2, a, b) == 2.0          # squared, no sqrt (hnsw.rs:70-71)
    assert O.vec_dist(O.COSINE, a, b) == 1.0
    assert O.vec_dist(O.IP, a, b) == 1.0
    assert O.vec_dist(O.IP, a, a) == 0.0


def test_unrolled_dot_order():
    # ndarray 0.15.6 unrolled_dot: 8 accumulators, (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then tail
    rng = np.random.default_rng(3)
    x = rng.standard_normal(27).astype(np.float32)
    y = rng.standard_normal(27).astype(np.float32)
    p = np.zeros(8, np.float32)
    for j in range(3):
        for i in range(8):
            p[i] = np.float32(p[i] + np.float32(x[8 * j + i] * y[8 * j + i]))
    s = np.float32(0)
    for i in range(4):
        s = np.float32(s + np.float32(p[i] + p[i + 4]))
    for i in range(24, 27):
        s = np.float32(s + np.float32(x[i] * y[i]))
    assert O.vec_dist(O.IP, x, y) == 1.0 - float(s)


@pytest.fixture(scope="module")
def small_index():
    X = uniform_vectors(3000, 48, SEED_DATA)
    ix = O.OracleHnsw.new(3000, 48, m=12, ef_construction=80, level_seed=SEED_LEVEL)
    ix.insert_all(X)
    return X, ix


def test_hnsw_structure(small_index):
    X, ix = small_index
    lv = ix.levels()
    assert lv.n_levels >= 2
    deg0 = np.diff(lv.row_ptr[0])
    assert deg0.max() <= 24 and deg0.min() >= 1           # m_max0 = 2m (relation.rs:1146)
    for L in range(1, lv.n_levels):
        assert np.diff(lv.row_ptr[L]).max() <= 12          # m_max = m
        assert np.all(np.diff(lv.node_ids[L].astype(np.int64)) > 0)
        # every node of layer -L is on layer -(L-1)
        lower = set(range(3000)) if L == 1 else set(lv.node_ids[L - 1].tolist())
        assert set(lv.node_ids[L].tolist()) <= lower
    # entry = smallest id on the top layer (first row in key order, hnsw.rs:891-899)
    assert lv.entry == int(lv.node_ids[-1].min())
    # no self loops, targets sorted (key order)
    rp, ci = lv.row_ptr[0], lv.col_idx[0]
    for i in range(0, 3000, 97):
        row = ci[rp[i]:rp[i + 1]]
        assert i not in row and np.all(np.diff(row.astype(np.int64)) > 0)


def test_hnsw_search_vs_bruteforce(small_index):
    X, ix = small_index
    Q = uniform_vectors(100, 48, SEED_QUERY)
    ids, dist, cnt, st = ix.search(Q, 10, 100)
    bi, bd = O.bruteforce_knn(X, Q, 10)
    assert recall(ids, bi) > 0.9
    assert np.all(cnt == 10)
    assert np.all(np.diff(dist, axis=1) >= 0)              # nearest first (hnsw.rs:1005)
    for i in range(0, 100, 7):                              # reported distances are the true ones
        for j in range(10):
            assert dist[i, j] == O.vec_dist(O.L2, Q[i], X[ids[i, j]])
    assert np.all(st[:, 0] >= st[:, 1]) and np.all(st[:, 1] >= 1)


def test_hnsw_radius_and_k(small_index):
    X, ix = small_index
    Q = uniform_vectors(8, 48, 99)
    ids, dist, cnt, _ = ix.search(Q, 10, 64)
    r = float(dist[:, 4].max())
    ids2, dist2, cnt2, _ = ix.search(Q, 10, 64, radius=r)
    for i in range(8):
        keep = dist[i] <= r
        assert cnt2[i] == keep.sum()
        assert np.array_equal(ids2[i, :cnt2[i]], ids[i][keep])
        assert np.all(ids2[i, cnt2[i]:] == 0xFFFFFFFF)
    ids3, _, cnt3, _ = ix.search(Q, 3, 64)
    assert np.array_equal(ids3, ids[:, :3]) and np.all(cnt3 == 3)


def test_hnsw_view_roundtrip_and_empty(small_index):
    X, ix = small_index
    Q = uniform_vectors(20, 48, 5)
    a = ix.search(Q, 5, 40)
    ix2 = O.OracleHnsw.from_levels(X, ix.levels())
    b = ix2.search(Q, 5, 40)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
    empty = O.OracleHnsw.new(10, 48)
    ids, dist, cnt, _ = empty.search(Q, 5, 40)
    assert np.all(cnt == 0) and np.all(ids == 0xFFFFFFFF)  # canary only (hnsw.rs:903-909)


def test_hnsw_remove_and_reinsert():
    X = uniform_vectors(400, 16, 11)
    ix = O.OracleHnsw.new(400, 16, m=6, ef_construction=40)
    ix.insert_all(X)
    for i in range(0, 400, 3):
        ix.remove(i)
    lv = ix.levels()
    gone = set(range(0, 400, 3))
    assert not (set(lv.col_idx[0].tolist()) & gone)
    Q = uniform_vectors(10, 16, 12)
    ids, _, cnt, _ = ix.search(Q, 5, 50)
    assert not (set(ids.ravel().tolist()) & gone)
    ix.insert(0, X[0])
    ids, dist, _, _ = ix.search(X[0:1], 1, 50)
    assert ids[0, 0] == 0 and dist[0, 0] == 0.0


def test_cosine_and_ip_indexes():
    X = uniform_vectors(1500, 24, 21) - 0.5
    Q = uniform_vectors(40, 24, 22) - 0.5
    for metric in (O.COSINE, O.IP):
        ix = O.OracleHnsw.new(1500, 24, metric=metric, m=10, ef_construction=60)
        ix.insert_all(X)
        ids, _, _, _ = ix.search(Q, 10, 120)
        bi, _ = O.bruteforce_knn(X, Q, 10, metric=metric)
        assert recall(ids, bi) > 0.8


# ---------------------------------------------------------------- graphs ---
def _random_graph(n, m, seed, weighted=True):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, m).astype(np.uint32)
    dst = rng.integers(0, n, m).astype(np.uint32)
    w = (rng.integers(1, 64, m) / 8.0).astype(np.float32) if weighted else None   # exact in f32
    return src, dst, w


def test_csr_layout():
    src = np.array([2, 0, 0, 2, 1, 0], np.uint32)
    dst = np.array([1, 2, 1, 0, 2, 1], np.uint32)
    w = np.array([5, 1, 2, 3, 4, 9], np.float32)
    g = O.OracleGraph(3, src, dst, w)
    op, oi, ow, ip, ii = g.export()
    assert op.tolist() == [0, 3, 4, 6] and oi.tolist() == [1, 1, 2, 2, 0, 1]   # sorted targets, parallel edge kept
    assert ow.tolist() == [2, 9, 1, 4, 3, 5]
    assert ip.tolist() == [0, 1, 4, 6] and ii.tolist() == [2, 0, 0, 2, 0, 1]


def test_pagerank_hand_graph():
    # 0->1, 0->2, 1->2, 2->0, 3->2 ; node 3 has no in-edges
    src = np.array([0, 0, 1, 2, 3], np.uint32)
    dst = np.array([1, 2, 2, 0, 2], np.uint32)
    g = O.OracleGraph(4, src, dst)
    s, it, err = g.pagerank(damping=0.85, tol=0.0, max_iter=1)
    init, base = 0.25, 0.15 / 4
    exp = [base + 0.85 * init, base + 0.85 * init / 2, base + 0.85 * (init / 2 + init + init), base]
    assert it == 1 and np.allclose(s, exp, rtol=1e-6)
    # no dangling redistribution / renormalisation: mass is not conserved in general
    s, it, err = g.pagerank(tol=1e-12, max_iter=200)
    s2, it2, _ = g.pagerank(tol=1e-12, max_iter=200, variant="gs")
    assert np.allclose(s, s2, rtol=2e-6)


def test_pagerank_dangling_node():
    # node 2 is dangling: its contribution is inf/unused, never read
    src = np.array([0, 1], np.uint32)
    dst = np.array([1, 2], np.uint32)
    g = O.OracleGraph(3, src, dst)
    s, it, _ = g.pagerank(tol=0.0, max_iter=3)
    assert np.all(np.isfinite(s)) and it == 3
    assert abs(s.sum() - 1.0) > 1e-3     # documents the "no redistribution" behaviour


def test_pagerank_stops_on_epsilon():
    src, dst, _ = _random_graph(200, 2000, 5, weighted=False)
    g = O.OracleGraph(200, src, dst)
    s, it, err = g.pagerank(tol=1e-4, max_iter=100)
    assert it < 100 and err < float(np.float32(1e-4))
    s, it, err = g.pagerank(tol=1e-4, max_iter=3)
    assert it == 3


def test_dijkstra_vs_scipy():
    import scipy.sparse as sp
    import scipy.sparse.csgraph as cg
    n = 300
    src, dst, w = _random_graph(n, 1500, 7)
    g = O.OracleGraph(n, src, dst, w)
    sources = np.arange(0, n, 13, dtype=np.uint32)
    dist, back = g.sssp(sources, n_threads=4)
    A = {}
    for s, t, ww in zip(src, dst, w):
        A[(int(s), int(t))] = min(A.get((int(s), int(t)), 1e30), float(ww))
    S = sp.csr_matrix((list(A.values()), ([k[0] for k in A], [k[1] for k in A])), shape=(n, n))
    ref = cg.dijkstra(S, indices=sources)
    assert np.array_equal(np.isfinite(ref), np.isfinite(dist))
    fin = np.isfinite(ref)
    assert np.array_equal(dist[fin], ref[fin].astype(np.float32))   # weights are dyadic: sums exact
    # back pointers give a path of that cost
    for si, s in enumerate(sources):
        for t in range(0, n, 17):
            if not np.isfinite(dist[si, t]) or t == s:
                continue
            cost, cur, hops = 0.0, t, 0
            while cur != s:
                p = int(back[si, cur])
                cost += A[(p, cur)]
                cur = p
                hops += 1
                assert hops <= n
            assert cost == dist[si, t]


def test_dijkstra_goals_stop_early():
    n = 200
    src, dst, w = _random_graph(n, 1200, 8)
    g = O.OracleGraph(n, src, dst, w)
    full, _ = g.sssp([0])
    goals = np.array([5, 9], np.uint32)
    part, _ = g.sssp([0], goals=goals)
    assert np.array_equal(part[0, goals], full[0, goals])


def test_keep_ties_predecessors():
    # diamond: 0->1->3, 0->2->3 with equal cost, plus a longer 0->3
    src = np.array([0, 0, 1, 2, 0], np.uint32)
    dst = np.array([1, 2, 3, 3, 3], np.uint32)
    w = np.array([1, 1, 1, 1, 5], np.float32)
    g = O.OracleGraph(4, src, dst, w)
    dist, bp, bi = g.sssp_keep_ties(0)
    assert dist.tolist() == [0, 1, 1, 2]
    assert sorted(bi[bp[3]:bp[4]].tolist()) == [1, 2]


def test_closeness_and_betweenness_vs_networkx():
    import networkx as nx
    n = 120
    src, dst, w = _random_graph(n, 700, 9)
    g = O.OracleGraph(n, src, dst, w)
    A = {}
    for s, t, ww in zip(src, dst, w):
        A[(int(s), int(t))] = min(A.get((int(s), int(t)), 1e30), float(ww))
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    for (a, b), ww in A.items():
        G.add_edge(a, b, weight=ww)
    bt = g.betweenness(n_threads=4)
    nb = nx.betweenness_centrality(G, normalized=False, weight="weight")
    assert np.allclose(bt, [nb[i] for i in range(n)], rtol=1e-5, atol=1e-5)
    cl = g.closeness(n_threads=4)
    lengths = dict(nx.all_pairs_dijkstra_path_length(G, weight="weight"))
    for u in range(0, n, 5):
        d = lengths[u]
        nc, tot = len(d), sum(d.values())     # nc counts u itself (d(u,u)=0), all_pairs_shortest_path.rs:118-120
        exp = np.inf if tot == 0 else nc * nc / tot / (n - 1)
        assert np.isclose(cl[u], exp, rtol=1e-5) or (np.isinf(exp) and np.isinf(cl[u]))


def test_rmat_generator_shape():
    n, s, d = rmat_edges(10, 16, 0x5EED0004)
    assert n == 1024 and s.size == 16384 and s.max() < n and d.max() < n


def test_frozen_oracle_outputs():
    """tests/golden/oracle_small.npz (made by tests/golden/make_oracle_golden.py): the oracle must keep
    producing what it produced when the fixtures were frozen."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_small.npz"))
    ix = O.OracleHnsw.new(600, 24, m=6, ef_construction=40, level_seed=SEED_LEVEL)
    ix.insert_all(z["X"])
    lv = ix.levels()
    assert lv.entry == int(z["entry"]) and lv.n_levels == int(z["n_levels"])
    assert np.array_equal(np.diff(lv.row_ptr[0]), z["deg0"]) and np.array_equal(lv.col_idx[0], z["col0"])
    ids, dist, cnt, st = ix.search(z["Q"], 8, 32)
    assert np.array_equal(ids, z["ids"]) and np.array_equal(dist, z["dist"]) and np.array_equal(st, z["stats"])
    g = O.OracleGraph(200, z["src"], z["dst"], z["w"])
    pr, it, _ = O.OracleGraph(200, z["src"], z["dst"]).pagerank(0.85, 1e-4, 10)
    assert it == int(z["pr_iters"]) and np.array_equal(pr, z["pr"])
    sd, _ = g.sssp(np.arange(0, 200, 25, dtype=np.uint32))
    assert np.array_equal(sd, z["sssp"])
    assert np.array_equal(g.closeness(), z["closeness"]) and np.array_equal(g.betweenness(), z["betweenness"])


def test_clustering_vs_networkx():
    """ClusteringCoefficients (triangles.rs:25-98) on a simple undirected graph == networkx"""
    import networkx as nx
    rng = np.random.default_rng(17)
    n = 150
    pairs = {(int(a), int(b)) for a, b in zip(rng.integers(0, n, 900), rng.integers(0, n, 900)) if a < b}
    src = np.array([p[0] for p in pairs] + [p[1] for p in pairs], np.uint32)     # mirrored (triangles.rs:35)
    dst = np.array([p[1] for p in pairs] + [p[0] for p in pairs], np.uint32)
    g = O.OracleGraph(n, src, dst)
    cc, nt, deg = g.clustering(n_threads=4)
    G = nx.Graph()
    G.add_nodes_from(range(n))
    G.add_edges_from(pairs)
    tri, clu = nx.triangles(G), nx.clustering(G)
    assert [int(x) for x in nt] == [tri[i] for i in range(n)]
    assert [int(x) for x in deg] == [G.degree[i] for i in range(n)]
    assert np.allclose(cc, [clu[i] for i in range(n)], rtol=1e-12, atol=0)
    # duplicate edges count per position (multigraph), exactly as the position loops of triangles.rs:74-91
    src2 = np.array([0, 1, 0, 2, 1, 2, 0, 1, 1, 0], np.uint32)
    dst2 = np.array([1, 0, 2, 0, 2, 1, 1, 0, 0, 1], np.uint32)
    cc2, nt2, deg2 = O.OracleGraph(3, src2, dst2).clustering()
    assert deg2.tolist() == [4, 4, 2] and nt2.tolist() == [3, 3, 1]


def test_yen_vs_networkx():
    """k_shortest_path_yen (yen.rs:120-211) on a simple digraph with distinct weights == networkx's
    shortest_simple_paths (costs; with distinct costs the paths are forced too)."""
    import itertools
    import networkx as nx
    rng = np.random.default_rng(3)
    n = 60
    pairs = sorted({(int(a), int(b)) for a, b in zip(rng.integers(0, n, 400), rng.integers(0, n, 400)) if a != b})
    src = np.array([p[0] for p in pairs], np.uint32)
    dst = np.array([p[1] for p in pairs], np.uint32)
    w = (rng.random(src.size) * 10 + 0.5).astype(np.float32)
    g = O.OracleGraph(n, src, dst, w)
    G = nx.DiGraph()
    for a, b, ww in zip(src, dst, w):
        G.add_edge(int(a), int(b), weight=float(ww))
    for s, t in itertools.product(range(0, 60, 13), range(3, 60, 17)):
        if s == t:
            continue
        res = [(c, p) for c, p in g.yen(s, t, 5) if np.isfinite(c)]
        try:
            ref = list(itertools.islice(nx.shortest_simple_paths(G, s, t, weight="weight"), 5))
        except nx.NetworkXNoPath:
            ref = []
        assert [p for _, p in res] == ref


def test_reference_goldens():
    """Pins the oracle against outputs of the REAL reference (tools/run_reference.sh: stock cozo, mem engine).
    The files can only be produced on a machine with cargo; while they are absent the oracle stays
    'parity unpinned'  and this test skips."""
    import json
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ph, pg = os.path.join(gdir, "reference_hnsw.json"), os.path.join(gdir, "reference_graph.json")
    if not (os.path.exists(ph) and os.path.exists(pg)):
        pytest.skip("tests/golden/reference_*.json not generated (needs cargo: tools/run_reference.sh)")
    from tests.util import SEED_DATA, SEED_QUERY, uniform_vectors
    X = uniform_vectors(10_000, 128, SEED_DATA)
    Q = uniform_vectors(1_000, 128, SEED_QUERY)
    ref = json.load(open(ph))
    rows = ref["index_rows"]["rows"]      # layer, fr_k, fr__field, fr__sub_idx, to_k, to__field, to__sub_idx, dist, hash, ignore_link
    top = -min(r[0] for r in rows if r[2] is not None)
    adj = [dict() for _ in range(top + 1)]
    for r in rows:
        if r[2] is None:                  # canary
            continue
        L = -r[0]
        adj[L].setdefault(r[1], [])
        if r[4] != r[1] and not r[9]:     # hnsw_get_neighbours reading rules (hnsw.rs:609, 618-620)
            adj[L][r[1]].append(r[4])
    first = min((r for r in rows if r[2] is not None), key=lambda r: (r[0], r[1]))
    node_ids, row_ptr, col_idx = [None], [], []
    for L in range(top + 1):
        nodes = list(range(10_000)) if L == 0 else sorted(adj[L])
        if L:
            node_ids.append(np.array(nodes, np.uint32))
        rp, ci = [0], []
        for u in nodes:
            ci += sorted(adj[L].get(u, []))
            rp.append(len(ci))
        row_ptr.append(np.array(rp, np.uint64))
        col_idx.append(np.array(ci, np.uint32))
    ix = O.OracleHnsw.from_levels(X, O.HnswLevels(node_ids, row_ptr, col_idx, first[1]))
    ids, dist, cnt, _ = ix.search(Q, 10, 64, n_threads=8)
    same = 0
    for qi, r in enumerate(ref["knn"]):
        got = {int(k) for k in ids[qi, :cnt[qi]]}
        want = {int(row[0]) for row in r["rows"]}
        same += got == want
        assert len(got & want) >= 9
        assert np.allclose(sorted(dist[qi, :cnt[qi]]), sorted(row[1] for row in r["rows"]), rtol=1e-6)
    assert same >= 995
    # graph rules on air-routes
    from tests.test_air_routes_cpu import load_routes
    n, src, dst, w, id_of, _ = load_routes()
    g = json.load(open(pg))
    og = O.OracleGraph(n, src, dst)
    want = {r[0]: r[1] for r in g["pagerank"]["rows"]}
    for variant in ("jacobi", "gs"):
        sc, _, _ = og.pagerank(0.85, 1e-4, 10, variant=variant)
        err = max(abs(sc[id_of[c]] - v) / v for c, v in want.items())
        print(f"PageRank oracle variant {variant}: max rel diff to the reference {err:.3e}")
    sc, _, _ = og.pagerank(0.85, 1e-4, 10, variant="jacobi")
    assert max(abs(sc[id_of[c]] - v) / v for c, v in want.items()) <= 1e-5
    ow = O.OracleGraph(n, src, dst, w)
    oc = ow.closeness(n_threads=8)
    for c, v in g["closeness"]["rows"]:
        assert (not np.isfinite(v) and not np.isfinite(oc[id_of[c]])) or abs(oc[id_of[c]] - v) <= 1e-5 * abs(v)
    ob = ow.betweenness(n_threads=8)
    for c, v in g["betweenness"]["rows"]:
        assert abs(ob[id_of[c]] - v) <= 1e-4 * max(1.0, abs(v))
