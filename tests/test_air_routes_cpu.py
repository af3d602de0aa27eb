"""The reference's air-routes fixture (cozo-core/tests/air_routes.rs) as a realistic graph input.
The reference's own tests on it only time Dijkstra (air_routes.rs:299-316) and pin nothing but the
relation's row count; here the oracle is cross-checked on it against scipy."""
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "air_routes.npz")


def load_routes():
    z = np.load(GOLDEN)
    codes = [str(c) for c in z["codes"]]
    fr, to, dist = z["fr"].astype(np.int64), z["to"].astype(np.int64), z["dist"]
    # dense ids in first-appearance order over the key-ordered relation (fixed_rule/mod.rs:164-179)
    ids = {}
    for a, b in zip(fr, to):
        ids.setdefault(int(a), len(ids))
        ids.setdefault(int(b), len(ids))
    src = np.array([ids[int(a)] for a in fr], np.uint32)
    dst = np.array([ids[int(b)] for b in to], np.uint32)
    name_of = {v: codes[k] for k, v in ids.items()}
    id_of = {codes[k]: v for k, v in ids.items()}
    return len(ids), src, dst, dist, id_of, name_of


def test_route_count_is_the_references():
    n, src, dst, dist, id_of, _ = load_routes()
    assert src.size == 50637          # air_routes.rs:189-209 pins 50637 rows of `route`
    assert n == 3476 and "JFK" in id_of and "KUL" in id_of


def test_dijkstra_air_routes_vs_scipy():
    import scipy.sparse as sp
    import scipy.sparse.csgraph as cg
    n, src, dst, dist, id_of, name_of = load_routes()
    g = O.OracleGraph(n, src, dst, dist)
    S = sp.csr_matrix((dist.astype(np.float64), (src, dst)), shape=(n, n))
    starts = np.array([id_of[c] for c in ("JFK", "LHR", "SYD", "AUS")], np.uint32)
    d, back = g.sssp(starts, n_threads=4)
    ref = cg.dijkstra(S, indices=starts)
    assert np.array_equal(np.isfinite(d), np.isfinite(ref))
    fin = np.isfinite(ref)
    assert np.array_equal(d[fin], ref[fin].astype(np.float32))      # integer mile counts: f32 sums are exact
    # the query of air_routes.rs:299-316: JFK -> KUL
    t = id_of["KUL"]
    path, cur = [], t
    while cur != starts[0]:
        path.append(cur)
        cur = int(back[0, cur])
    path.append(int(starts[0]))
    assert 2 <= len(path) <= 4 and d[0, t] == ref[0, t]
