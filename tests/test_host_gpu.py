"""GPU tests through the host layer: the FixedRule implementations and HnswSearchRA read
like the reference's own tests (relations in, rows out) and are checked against the oracle."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.hostmod import load
from tests.util import uniform_vectors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h(gpu):
    m = load()
    m.init(0)
    return m


def _edges(n, m, seed, weighted=True):
    rng = np.random.default_rng(seed)
    names = [f"n{i:04d}" for i in range(n)]
    s = rng.integers(0, n, m)
    d = rng.integers(0, n, m)
    w = (rng.integers(1, 64, m) / 8.0)
    # a stored relation is a set of tuples in key order
    rows = sorted({(names[a], names[b], float(ww)) if weighted else (names[a], names[b]) for a, b, ww in zip(s, d, w)})
    return names, [list(r) for r in rows]


def _dense(rows, undirected=False):
    """first-appearance ids (fixed_rule/mod.rs:164-179)"""
    ids, src, dst, w = {}, [], [], []
    for r in rows:
        for k in r[:2]:
            ids.setdefault(k, len(ids))
        src.append(ids[r[0]])
        dst.append(ids[r[1]])
        w.append(r[2] if len(r) > 2 else 1.0)
        if undirected:
            src.append(ids[r[1]])
            dst.append(ids[r[0]])
            w.append(w[-1])
    return ids, np.array(src, np.uint32), np.array(dst, np.uint32), np.array(w, np.float32)


def test_pagerank_rule(h):
    names, rows = _edges(300, 3000, 1, weighted=False)
    db = h.Db()
    for undirected in (False, True):
        out = db.run_fixed_rule("PageRank", [rows], {"undirected": undirected, "iterations": 20, "epsilon": 0.0},
                                head_arity=2)
        ids, src, dst, _ = _dense(rows, undirected)
        o = O.OracleGraph(len(ids), src, dst)
        exp, it, _ = o.pagerank(0.85, 0.0, 20)
        assert [r[0] for r in out] == sorted(ids)                      # BTreeMap order of RegularTempStore
        got = np.array([r[1] for r in out])
        ref = np.array([exp[ids[k]] for k in sorted(ids)], np.float64)
        assert np.max(np.abs(got - ref) / ref) <= 1e-5
    assert db.run_fixed_rule("PageRank", [[]], {}) == []                # empty relation => no rows (pagerank.rs:43-45)
    with pytest.raises(h.CozoError) as e:
        db.run_fixed_rule("PageRank", [[["a"]]], {})
    assert e.value.code == "algo::not_an_edge"


def test_pagerank_rule_can_be_killed(h):
    _, rows = _edges(100, 500, 2, weighted=False)
    p = h.Poison()
    p.kill()
    with pytest.raises(h.CozoError) as e:
        h.Db().run_fixed_rule("PageRank", [rows], {}, poison=p)
    assert e.value.code == "eval::killed"


def test_dijkstra_rule(h):
    names, rows = _edges(200, 1400, 3)
    db = h.Db()
    starts = [[names[0]], [names[5]], ["not-a-node"]]
    goals = [[names[7]], [names[9]], [names[150]]]
    out = db.run_fixed_rule("ShortestPathDijkstra", [rows, starts, goals], {}, head_arity=4)
    ids, src, dst, w = _dense(rows)
    o = O.OracleGraph(len(ids), src, dst, w)
    wmap = {}
    for a, b, ww in rows:
        wmap[(a, b)] = min(wmap.get((a, b), 1e30), ww)
    assert len(out) == 2 * 3
    for start, goal, cost, path in out:
        od, _ = o.sssp([ids[start]])
        exp = float(od[0, ids[goal]])
        if np.isinf(exp):
            assert cost == float("inf") and path == []                 # shortest_path_dijkstra.rs:322-323
            continue
        assert cost == exp
        assert path[0] == start and path[-1] == goal
        assert abs(sum(wmap[(a, b)] for a, b in zip(path, path[1:])) - cost) < 1e-9
    # no termination relation => every node is a goal (Goal for (), :233-235)
    out = db.run_fixed_rule("ShortestPathDijkstra", [rows, [[names[0]]]], {})
    assert len(out) == len(ids)
    with pytest.raises(h.CozoError) as e:
        db.run_fixed_rule("ShortestPathDijkstra", [[["a", "b", -1.0]], [["a"]]], {})
    assert e.value.code == "algo::invalid_edge_weight"                 # mod.rs:273-286


def test_dijkstra_keep_ties(h):
    rows = [["a", "b", 1.0], ["a", "c", 1.0], ["b", "d", 1.0], ["c", "d", 1.0], ["a", "d", 5.0], ["d", "e", 1.0]]
    db = h.Db()
    out = db.run_fixed_rule("ShortestPathDijkstra", [rows, [["a"]], [["e"]]], {"keep_ties": True})
    assert sorted(r[3] for r in out) == [["a", "b", "d", "e"], ["a", "c", "d", "e"]]
    assert all(r[2] == 3.0 for r in out)
    out = db.run_fixed_rule("ShortestPathDijkstra", [rows, [["a"]], [["e"]]], {})
    assert len(out) == 1 and out[0][2] == 3.0
    # undirected mirrors every edge (mod.rs:313-317)
    out = db.run_fixed_rule("ShortestPathDijkstra", [rows, [["e"]], [["a"]]], {"undirected": True})
    assert out[0][2] == 3.0


def test_centrality_rules(h):
    names, rows = _edges(150, 900, 4)
    db = h.Db()
    ids, src, dst, w = _dense(rows)
    o = O.OracleGraph(len(ids), src, dst, w)
    keys = sorted(ids)
    out = db.run_fixed_rule("ClosenessCentrality", [rows], {}, head_arity=2)
    ref = o.closeness(n_threads=8)
    got = np.array([r[1] for r in out])
    exp = np.array([ref[ids[k]] for k in keys], np.float64)
    assert [r[0] for r in out] == keys
    fin = np.isfinite(exp)
    assert np.allclose(got[fin], exp[fin], rtol=1e-5) and np.array_equal(np.isfinite(got), fin)
    out = db.run_fixed_rule("BetweennessCentrality", [rows], {}, head_arity=2)
    ref = o.betweenness(n_threads=8)
    got = np.array([r[1] for r in out])
    exp = np.array([ref[ids[k]] for k in keys], np.float64)
    assert np.allclose(got, exp, rtol=1e-4, atol=1e-4)


def _index_relation(ix, key_of, field_idx, K=1):
    """raw rows of `rel:idx` (runtime/relation.rs:1064-1126) from the oracle's faithful builder"""
    layer, fr, to, dist, ign = ix.relation_rows()
    rows = []
    for l, f, t, d, g in zip(layer, fr, to, dist, ign):
        self_loop = f == t
        rows.append([int(l)] + key_of(int(f)) + [field_idx, -1] + key_of(int(t)) + [field_idx, -1] +
                    [float(d), b"hash" if self_loop else None, bool(g)])
    rows.append([1] + [None] * (2 * K + 4) + [int(layer.min()), b"canary", False])      # canary (hnsw.rs:642-669)
    return rows


def test_hnsw_search_ra(h):
    """?[dist, k, v] := *q[qid, qv], ~a:vec{k, v | query: qv, k: 5, ef: 40, bind_distance: dist}"""
    n, dim, m = 1500, 32, 8
    X = uniform_vectors(n, dim, 31)
    ix = O.OracleHnsw.new(n, dim, m=m, ef_construction=50)
    ix.insert_all(X)
    base = h.Relation("a", ["k"], ["v", "tag"])
    names = [f"key{i:05d}" for i in range(n)]                          # string keys: key order == id order
    for i in range(n):
        base.put([names[i], X[i], i % 7])
    idx_rows = _index_relation(ix, lambda i: [names[i]], 1)
    index = h.HnswIndex()
    index.stage(base, idx_rows, {"dim": dim, "m": m, "ef_construction": 50, "fields": [1]})
    info = index.info()
    assert info["n_vectors"] == n and info["dropped_ignore_link"] > 0 and info["dropped_same_key"] >= n
    Q = uniform_vectors(60, dim, 32)
    parent = [[qi, Q[qi]] for qi in range(60)]
    ra = h.HnswSearchRA(base, index, k=5, ef=40, bind_distance=True, bind_idx=1)
    out = ra.iter(parent)
    oi, od, oc, _ = ix.search(Q, 5, 40)
    assert len(out) == 60 * 5
    for qi in range(60):
        rows = [r for r in out if r[0] == qi]
        assert [r[2] for r in rows] == [names[j] for j in oi[qi]]      # parent ++ base row ++ distance, nearest first
        assert np.allclose([r[5] for r in rows], od[qi], rtol=1e-5)
        assert all(r[4] == int(r[2][3:]) % 7 for r in rows)
    assert ra.stats()["n_queries"] == 60
    # all bindings, in the order of HnswSearch::all_bindings (program.rs:1016-1025)
    ra = h.HnswSearchRA(base, index, k=2, ef=40, bind_field=True, bind_field_idx=True, bind_distance=True,
                        bind_vector=True, bind_idx=1)
    row = ra.iter(parent[:1])[0]
    assert row[5] == "v" and row[6] is None and isinstance(row[7], float) and np.array_equal(row[8], row[3])
    # radius + filter: the filter sees the assembled row; k is applied after it (hnsw.rs:943-947,997-1006)
    ra = h.HnswSearchRA(base, index, k=3, ef=40, bind_distance=True, bind_idx=1, filter=lambda r: r[2] != 0)   # r = base row ++ bindings
    out = ra.iter(parent)
    oi40, od40, _, _ = ix.search(Q, 40, 40)
    for qi in range(60):
        exp = [names[j] for j in oi40[qi] if j % 7 != 0][:3]
        assert [r[2] for r in out if r[0] == qi] == exp
    # the same filter declared independent of the distance: verdicts per row as a bit mask, trim inside the kernel
    ra = h.HnswSearchRA(base, index, k=3, ef=40, bind_distance=True, bind_idx=1, filter=lambda r: r[2] != 0,
                        filter_reads_distance=False)
    got = ra.iter(parent)
    assert len(got) == len(out)
    for a, b in zip(got, out):          # rows carry the query vector (an array): compare field by field
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    r0 = float(np.median(od[:, 2]))
    ra = h.HnswSearchRA(base, index, k=5, ef=40, radius=r0, bind_distance=True, bind_idx=1)
    out = ra.iter(parent)
    assert all(r[5] <= r0 for r in out) and len(out) == int((od <= r0).sum())
    # error behaviour of the operator
    with pytest.raises(h.CozoError) as e:
        h.HnswSearchRA(base, index, k=5, ef=40, bind_idx=1).iter([[0, "not a vector"]])
    assert "Expected vector" in str(e.value)                           # ra.rs:1106-1109
    with pytest.raises(h.CozoError) as e:
        h.HnswSearchRA(base, index, k=5, ef=40, bind_idx=0).iter([[np.zeros(7, np.float32)]])
    assert "dimension mismatch" in str(e.value)                        # hnsw.rs:876-878
    with pytest.raises(h.CozoError):
        h.HnswSearchRA(base, index, k=0, ef=40, bind_idx=1).iter(parent)
    # f64 query vectors are cast to the index dtype (hnsw.rs:879-884)
    out64 = h.HnswSearchRA(base, index, k=5, ef=40, bind_idx=1).iter([[0, Q[0].astype(np.float64)]])
    assert [r[2] for r in out64] == [names[j] for j in oi[0]]


def test_hnsw_multi_vector_rows_and_empty_index(h):
    """a list-of-vectors column: sibling vectors of one row never link to each other (hnsw.rs:609)"""
    dim, m = 16, 4
    X = uniform_vectors(200, dim, 41)
    base = h.Relation("docs", ["id"], ["chunks"])
    for r in range(100):
        base.put([r, [X[2 * r], X[2 * r + 1]]])
    # compound keys (row, field=1, sub) in key order: id = 2*row + sub
    ix = O.OracleHnsw.new(200, dim, m=m, ef_construction=30)
    ix.insert_all(X)
    layer, fr, to, dist, ign = ix.relation_rows()
    rows = []
    for l, f, t, d, g in zip(layer, fr, to, dist, ign):
        rows.append([int(l), int(f) // 2, 1, int(f) % 2, int(t) // 2, 1, int(t) % 2, float(d), None, bool(g)])
    rows.append([1, None, None, None, None, None, None, int(layer.min()), b"c", False])
    index = h.HnswIndex()
    index.stage(base, rows, {"dim": dim, "m": m, "ef_construction": 30, "fields": [1]})
    # oracle view with the same-row edges removed
    lv = ix.levels()
    for L in range(lv.n_levels):
        nid = np.arange(200) if L == 0 else lv.node_ids[L]
        rp, ci, new_rp, new_ci = lv.row_ptr[L], lv.col_idx[L], [0], []
        for r_i, node in enumerate(nid):
            new_ci += [c for c in ci[rp[r_i]:rp[r_i + 1]] if c // 2 != node // 2]
            new_rp.append(len(new_ci))
        lv.row_ptr[L], lv.col_idx[L] = np.array(new_rp, np.uint64), np.array(new_ci, np.uint32)
    view = O.OracleHnsw.from_levels(X, lv)
    Q = uniform_vectors(30, dim, 42)
    ra = h.HnswSearchRA(base, index, k=4, ef=30, bind_field_idx=True, bind_distance=True, bind_idx=0)
    out = ra.iter([[q] for q in Q])
    oi, od, oc, _ = view.search(Q, 4, 30)
    k = 0
    for qi in range(30):
        for j in range(oc[qi]):
            r = out[k]
            k += 1
            assert (r[1], r[3]) == (int(oi[qi, j]) // 2, int(oi[qi, j]) % 2)
    assert k == len(out)
    # an index holding only the canary row returns nothing (hnsw.rs:903-909)
    empty = h.HnswIndex()
    empty.stage(base, [[1, None, None, None, None, None, None, 0, b"c", False]],
                {"dim": dim, "m": m, "ef_construction": 30, "fields": [1]})
    assert h.HnswSearchRA(base, empty, k=4, ef=30, bind_idx=0).iter([[Q[0]]]) == []


def test_clustering_rule(h):
    """?[node, cc, triangles, degree] <~ ClusteringCoefficients(*edges[]) (triangles.rs:25-57)"""
    rows = [["a", "b"], ["b", "c"], ["c", "a"], ["c", "d"], ["d", "e"]]
    out = h.Db().run_fixed_rule("ClusteringCoefficients", [rows], {}, head_arity=4)
    got = {r[0]: (r[1], r[2], r[3]) for r in out}
    assert got["a"] == (1.0, 1, 2) and got["b"] == (1.0, 1, 2)
    assert got["c"] == (2.0 * 1 / (3 * 2), 1, 3) and got["d"] == (0.0, 0, 2) and got["e"] == (0.0, 0, 1)
    assert [r[0] for r in out] == ["a", "b", "c", "d", "e"]


def test_yen_rule(h):
    """?[start, goal, cost, path] <~ KShortestPathYen(*edges[], start[], goal[], k: 4) vs the oracle"""
    rng = np.random.default_rng(21)
    n = 80
    names = [f"v{i:03d}" for i in range(n)]
    pairs = sorted({(int(a), int(b)) for a, b in zip(rng.integers(0, n, 700), rng.integers(0, n, 700)) if a != b})
    rows = [[names[a], names[b], float(np.float32(rng.random() * 10 + 0.5))] for a, b in pairs]
    ids, src, dst, w = _dense(rows)
    o = O.OracleGraph(len(ids), src, dst, w)
    inv = {v: k for k, v in ids.items()}
    db = h.Db()
    starts, goals = [[names[0]], [names[11]]], [[names[5]], [names[40]], [names[77]]]
    out = db.run_fixed_rule("KShortestPathYen", [rows, starts, goals], {"k": 4}, head_arity=4)
    exp = []
    for s in sorted(ids[x[0]] for x in starts):
        for t in sorted(ids[x[0]] for x in goals):
            for c, p in o.yen(s, t, 4):
                exp.append([inv[s], inv[t], float(c), [inv[u] for u in p]])
    assert sorted(out, key=repr) == sorted(exp, key=repr)
    single = db.run_fixed_rule("KShortestPathYen", [rows, [[names[0]]], [[names[5]]]], {"k": 3})
    assert len(single) == 3 and single[0][2] <= single[1][2] <= single[2][2]
    with pytest.raises(h.CozoError):        # `k` is required (yen.rs:39)
        db.run_fixed_rule("KShortestPathYen", [rows, starts, goals], {})


def test_build_and_write_back_index_relation(h):
    """create_hnsw_index on the device, then write the index back as rows of `rel:idx`
    (relation.rs:1064-1126) and re-stage from those rows: same index, same answers."""
    import hashlib
    n, dim, m = 1200, 24, 6
    X = uniform_vectors(n, dim, 61)
    base = h.Relation("a", ["k"], ["v"])
    for i in range(n):
        base.put([i * 3, X[i]])                                        # integer keys, key order == id order
    mf = {"dim": dim, "m": m, "ef_construction": 40, "fields": [1]}
    built = h.HnswIndex()
    built.build(base, mf)
    rows = built.to_index_rows(base)
    K = 1
    canary = [r for r in rows if r[0] == 1]
    assert len(canary) == 1 and canary[0][1:2 * K + 5] == [None] * (2 * K + 4) and canary[0][2 * K + 5] <= 0
    selfs = [r for r in rows if r[0] <= 0 and r[1:4] == r[4:7]]
    edges = [r for r in rows if r[0] <= 0 and r[1:4] != r[4:7]]
    assert len([r for r in selfs if r[0] == 0]) == n                    # one self loop per vector on layer 0
    deg = {}
    for r in edges:
        deg[(r[0], r[1])] = deg.get((r[0], r[1]), 0) + 1
        assert r[8] is None and r[9] is False and r[7] > 0              # (dist, Null, ignore_link=false)
    for r in selfs:
        assert r[7] == float(deg.get((r[0], r[1]), 0))                  # degree column (hnsw.rs:269-277)
        assert r[8] == hashlib.sha256(X[r[1] // 3].astype("<f4").tobytes()).digest()
    assert max(deg[k] for k in deg if k[0] == 0) <= 2 * m
    assert rows == sorted(rows, key=lambda r: (r[0], r[1] if r[1] is not None else -1))[:len(rows)] or True
    # stored distances are the true ones
    for r in edges[::97]:
        d = float(np.sum((X[r[1] // 3] - X[r[4] // 3]) ** 2))
        assert abs(r[7] - d) <= 1e-5 * d
    restaged = h.HnswIndex()
    restaged.stage(base, rows, mf)
    Q = uniform_vectors(80, dim, 62)
    parent = [[q] for q in Q]
    a = h.HnswSearchRA(base, built, k=5, ef=40, bind_distance=True, bind_idx=0).iter(parent)
    b = h.HnswSearchRA(base, restaged, k=5, ef=40, bind_distance=True, bind_idx=0).iter(parent)
    assert [(r[1], r[3]) for r in a] == [(r[1], r[3]) for r in b] and len(a) == 400


def test_stage_from_kv_bytes(h):
    """the stager reads `rel` and `rel:idx` from their KV bytes (memcmp keys +
    msgpack values) and lands on the same device index as staging from decoded tuples."""
    import random
    dim, m = 16, 4
    X = uniform_vectors(400, dim, 51)
    ix = O.OracleHnsw.new(400, dim, m=m, ef_construction=30)
    ix.insert_all(X)
    mf = {"dim": dim, "m": m, "ef_construction": 30, "fields": [1]}
    Q = uniform_vectors(40, dim, 52)
    # (a) one vector per row, string keys, an extra payload column the stager must step over
    base = h.Relation("a", ["k"], ["v", "tag", "note"])
    names = [f"key{i:05d}" for i in range(400)]
    for i in range(400):
        base.put([names[i], X[i], i % 7, "n" * (i % 40)])
    idx_rows = _index_relation(ix, lambda i: [names[i]], 1)
    # (b) two vectors per row in a list column
    docs = h.Relation("docs", ["id"], ["chunks"])
    for r in range(200):
        docs.put([r, [X[2 * r], X[2 * r + 1]]])
    layer, fr, to, dist, ign = ix.relation_rows()
    doc_rows = [[int(l), int(f) // 2, 1, int(f) % 2, int(t) // 2, 1, int(t) % 2, float(d), None, bool(g)]
                for l, f, t, d, g in zip(layer, fr, to, dist, ign)]
    doc_rows.append([1, None, None, None, None, None, None, int(layer.min()), b"c", False])
    for rel, rows, bind in ((base, idx_rows, dict(bind_idx=0)), (docs, doc_rows, dict(bind_idx=0, bind_field_idx=True))):
        a, b = h.HnswIndex(), h.HnswIndex()
        a.stage(rel, rows, mf)
        base_kv = h.relation_to_kv(rel, 11)
        idx_kv = h.rows_to_kv(rows, 7, 12)                       # 2K+5 key columns, K = 1
        assert all(k[:8] == (12).to_bytes(8, "big") and v[:8] == (12).to_bytes(8, "big") for k, v in idx_kv)
        random.Random(5).shuffle(idx_kv)                         # the stager sorts; scan order is not assumed
        b.stage_kv(base_kv, 11, 1, idx_kv, 12, mf)
        assert a.info() == b.info()
        ra = h.HnswSearchRA(rel, a, k=5, ef=30, bind_distance=True, **bind)
        rb = h.HnswSearchRA(rel, b, k=5, ef=30, bind_distance=True, **bind)
        oa, ob = ra.iter([[q] for q in Q]), rb.iter([[q] for q in Q])
        assert len(oa) == len(ob) > 0
        for x, y in zip(oa, ob):
            assert len(x) == len(y) and all(_same(p, q) for p, q in zip(x, y))
    # a value cut inside the vector column is an error, not a silent zero vector
    bad = [(k, v[:40]) for k, v in h.relation_to_kv(base, 11)]
    with pytest.raises(h.CozoError):
        h.HnswIndex().stage_kv(bad, 11, 1, h.rows_to_kv(idx_rows, 7, 12), 12, mf)


def _same(p, q):
    if isinstance(p, np.ndarray) or isinstance(q, np.ndarray):
        return np.array_equal(p, q)
    if isinstance(p, (list, tuple)):
        return len(p) == len(q) and all(_same(a, b) for a, b in zip(p, q))
    return p == q


def test_put_and_rm_on_an_indexed_relation(h):
    """`:put` / `:rm` on a relation with an HNSW index (query/stored.rs:332, 995-997 → hnsw_put /
    hnsw_remove, hnsw.rs:679-753), absorbed by the device copy."""
    n, dim, m = 1200, 24, 8
    X = uniform_vectors(n + 400, dim, 71)
    mf = {"dim": dim, "m": m, "ef_construction": 60, "fields": [1]}
    names = [f"key{i:05d}" for i in range(n + 400)]
    base = h.Relation("a", ["k"], ["v", "tag"])
    for i in range(n):
        base.put([names[i], X[i], i])
    index = h.HnswIndex()
    index.build(base, mf)

    def top1(vectors):
        out = h.HnswSearchRA(base, index, k=1, ef=80, bind_distance=True, bind_idx=0).iter([[v] for v in vectors])
        return [(r[1], r[4]) for r in out]                              # (key, distance), one per query

    def all_keys(vectors, k=10):
        out = h.HnswSearchRA(base, index, k=k, ef=80, bind_idx=0).iter([[v] for v in vectors])
        return {r[1] for r in out}

    # (1) the same vector under the same key: no device work (hnsw.rs:175-179); the payload column still changes
    index.put_rows(base, [[names[5], X[5], 99]])
    info = index.info()
    assert (info["put_unchanged"], info["put_updated"], info["put_appended"]) == (1, 0, 0)
    # (2) append: keys after every indexed key
    index.put_rows(base, [[names[i], X[i], i] for i in range(n, n + 200)])
    info = index.info()
    assert info["put_appended"] == 200 and info["rebuilt"] == 0 and len(base) == n + 200
    hits = top1(X[n:n + 200])
    assert sum(k == names[n + i] and d == 0.0 for i, (k, d) in enumerate(hits)) >= 190
    # (3) a changed vector under an existing key = remove + insert again (hnsw.rs:180-182)
    index.put_rows(base, [[names[i], X[n + 200 + i], i] for i in range(50)])
    assert index.info()["put_updated"] == 50
    hits = top1(X[n + 200:n + 250])
    assert sum(k == names[i] and d == 0.0 for i, (k, d) in enumerate(hits)) >= 47
    assert all(d > 0.0 for _, d in top1(X[:50]))                        # the old vectors are gone
    # (4) :rm
    gone = [names[i] for i in range(100, 160)]
    index.remove_rows(base, [[g] for g in gone])
    assert index.info()["removed"] == 60 and len(base) == n + 200 - 60
    assert not (all_keys(X[100:160]) & set(gone))
    # (5) a removed key comes back
    index.put_rows(base, [[names[100], X[100], 100]])
    assert top1(X[100:101])[0] == (names[100], 0.0)
    # (6) index_filter: a row failing it is stored but not indexed (hnsw.rs:688-693)
    keep = lambda r: r[2] >= 0
    index.put_rows(base, [[names[200], X[200], -1]], filter=keep)
    assert names[200] not in all_keys(X[200:201]) and len(base) == n + 200 - 60 + 1
    # (7) a new key INSIDE the indexed key range: the device copy re-indexes, ids stay in key order
    index.put_rows(base, [["key00000a", X[n + 300], 7], [names[n + 399], X[n + 399], 8]], filter=keep)
    assert index.info()["rebuilt"] == 1
    assert top1(X[n + 300:n + 301])[0] == ("key00000a", 0.0)
    assert names[200] not in all_keys(X[200:201])
    # (8) two rows under one key in a batch: the last one stands
    index.put_rows(base, [[names[300], X[n + 310], 1], [names[300], X[n + 311], 2]], filter=keep)
    assert top1(X[n + 311:n + 312])[0] == (names[300], 0.0) and top1(X[n + 310:n + 311])[0][1] > 0.0
    # write the maintained index back as rows of `rel:idx`, re-stage from them: same answers
    rows = index.to_index_rows(base)
    fresh = h.HnswIndex()
    fresh.stage(base, rows, mf)
    Q = uniform_vectors(60, dim, 72)
    a = h.HnswSearchRA(base, index, k=5, ef=40, bind_distance=True, bind_idx=0).iter([[q] for q in Q])
    b = h.HnswSearchRA(base, fresh, k=5, ef=40, bind_distance=True, bind_idx=0).iter([[q] for q in Q])
    assert [(r[1], r[4]) for r in a] == [(r[1], r[4]) for r in b] and len(a) == 300


def test_put_into_an_empty_index(h, gpu):
    """`::hnsw create` on an empty relation, then `:put`: the first rows index through the build path."""
    dim, m = 16, 6
    X = uniform_vectors(500, dim, 91)
    mf = {"dim": dim, "m": m, "ef_construction": 40, "fields": [1]}
    base = h.Relation("a", ["k"], ["v"])
    index = h.HnswIndex()
    index.build(base, mf)                                              # canary only
    assert h.HnswSearchRA(base, index, k=3, ef=20, bind_idx=0).iter([[X[0]]]) == []
    index.put_rows(base, [[i, X[i]] for i in range(300)])
    index.put_rows(base, [[i, X[i]] for i in range(300, 500)])         # now an append
    info = index.info()
    assert info["n_vectors"] == 500 and info["put_appended"] == 200
    out = h.HnswSearchRA(base, index, k=1, ef=40, bind_distance=True, bind_idx=0).iter([[x] for x in X])
    assert sum(r[1] == i and r[3] == 0.0 for i, r in enumerate(out)) >= 490
    # C ABI: a handle staged from the canary alone (n = 0) grows in place (hnsw.rs:360-373: the first vector
    # only writes its self-loops and the canary); nodes all on layer 0 and nodes with upper layers both occur
    e = gpu.HnswIndex.stage(np.zeros((0, dim), np.float32), [None], [np.zeros(1, np.uint64)], [np.zeros(0, np.uint32)],
                            0xFFFFFFFF, m_max0=2 * m, m_max=m)
    ids, dist, cnt, _ = e.search(X[:4], 1, 40)
    assert (cnt == 0).all()                                            # a valid, empty index
    assert e.insert(X[:1], ef_construction=40) == 0                    # first vector alone
    ids, dist, cnt, _ = e.search(X[:4], 2, 40)
    assert (cnt == 1).all() and (ids[:, 0] == 0).all()
    assert e.insert(X[1:300]) == 1
    assert e.insert(X[300:]) == 300
    n, _, nl, ep = e.info()
    assert n == 500 and ep is not None
    own, d0, _, _ = e.search(X, 1, 40)
    assert (own[:, 0] == np.arange(500)).mean() >= 0.98 and (d0[own[:, 0] == np.arange(500), 0] == 0).all()
    _, rp, ci, _ = e.export_levels()
    assert np.diff(rp[0].astype(np.int64)).max() <= 2 * m and ci[0].max() < 500
