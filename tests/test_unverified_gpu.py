"""GPU tests of code that has NOT run on a device yet.

gpurun was closed mid-round (three boxes were lost under long calls while the pod was draining), after these
features were written but before they could be run: the propagation-blocking PageRank engine (pagerank.mode=1), the
compacted-frontier and wide SSSP forms (sssp.frontier / sssp.wide), the zero-weight-cycle refusal of betweenness, the
builder's fidelity mode (batches of one) and extend_candidates.  All of them are off by default (or, for the builder
modes, only reached through explicit arguments).  The tests carry the marker `gpu_unverified` instead of `gpu`, so the
round-end `pytest -m gpu` does not select them, and they skip unless COZO_RUN_UNVERIFIED=1 is set:

    COZO_RUN_UNVERIFIED=1 python -m pytest tests/test_unverified_gpu.py -q

The layout logic of the PageRank engine is additionally covered on the CPU by tests/test_pagerank_model_cpu.py, and this
whole file passes (12 / 12) against the CPU-emulated build of the library (tests/test_emu_lib_cpu.py; record in
profiles/r02_emu_gpu_testfiles.txt) — which is how the extend_candidates interleaving bug was found and fixed.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.test_graph_gpu import _check_tree, _random_graph
from tests.util import recall, rmat_edges, uniform_vectors

pytestmark = [pytest.mark.gpu_unverified,
              pytest.mark.skipif(os.environ.get("COZO_RUN_UNVERIFIED") != "1", reason="not yet run on a GPU; opt in with COZO_RUN_UNVERIFIED=1")]


PR_DEFAULTS = {"pagerank.mode": 1, "pagerank.hub_slots": 16384, "pagerank.group_slots": 32768, "pagerank.window": 24576,
               "pagerank.chunk": 262144}
PR_ENGINES = [
    {"pagerank.mode": 0},                                                      # round-1 gather pull
    {},                                                                        # propagation blocking, defaults
    {"pagerank.hub_slots": 64, "pagerank.group_slots": 256, "pagerank.window": 512, "pagerank.chunk": 1024},
    {"pagerank.hub_slots": 0, "pagerank.group_slots": 64, "pagerank.window": 64},          # no hub table, tiny tiles
    {"pagerank.hub_slots": 1024, "pagerank.group_slots": 4096, "pagerank.window": 3001},   # odd window
]


@pytest.mark.parametrize("engine", range(len(PR_ENGINES)))
def test_pagerank_engines_and_blockings(gpu, engine):
    """every engine / blocking geometry computes the same iteration: <= 1e-5 of the oracle on R-MAT (all code paths of
    the blocking: several groups, many bins, rows that straddle windows, empty rows), bit-identical reruns"""
    opts = {**PR_DEFAULTS, **PR_ENGINES[engine]}
    try:
        for k, v in opts.items():
            gpu.set_option(k, v)
        for scale, iters in ((13, 6), (9, 4)):
            n, src, dst = rmat_edges(scale, 16, 0x5EED0004 + scale)
            g = gpu.Graph(n, src, dst)
            os_, oit, oerr = O.OracleGraph(n, src, dst).pagerank(0.85, 0.0, iters, variant="jacobi", n_threads=8)
            gs, git, gerr, _ = g.pagerank(0.85, 0.0, iters)
            assert git == oit and np.max(np.abs(gs - os_) / os_) <= 1e-5
            assert abs(gerr - oerr) <= 0.02 * oerr + 1e-6
            assert np.array_equal(gs, g.pagerank(0.85, 0.0, iters)[0])
        # a star: one row with 4999 in-edges (many windows long at the small geometries) + reverse edges
        n = 5000
        src = np.concatenate([np.arange(1, n), np.zeros(n - 1)]).astype(np.uint32)
        dst = np.concatenate([np.zeros(n - 1), np.arange(1, n)]).astype(np.uint32)
        gs = gpu.Graph(n, src, dst).pagerank(0.85, 0.0, 5)[0]
        s64 = np.full(n, 1.0 / n)
        outdeg = np.bincount(src, minlength=n).astype(np.float64)
        for _ in range(5):
            nxt = np.full(n, 0.15 / n)
            np.add.at(nxt, dst, 0.85 * (s64 / outdeg)[src])
            s64 = nxt
        assert np.max(np.abs(gs - s64) / s64) <= 4e-6
        # a path graph and a graph whose last ids have no edges at all
        src = np.arange(0, 300, dtype=np.uint32)
        dst = src + 1
        g = gpu.Graph(1000, src, dst)
        os_, _, _ = O.OracleGraph(1000, src, dst).pagerank(0.85, 0.0, 7)
        assert np.allclose(g.pagerank(0.85, 0.0, 7)[0], os_, rtol=1e-6)
    finally:
        for k, v in PR_DEFAULTS.items():
            gpu.set_option(k, v)


def test_sssp_queue_frontier_and_large_graph(gpu):
    """graphs whose per-source state does not fit shared memory use the compacted frontier queues: same fixed
    point, bit-identical distances (forced here on a small graph, natural on RMAT-16)"""
    n = 500
    src, dst, w = _random_graph(n, 4000, 5, dyadic=False)
    g = gpu.Graph(n, src, dst, w)
    o = O.OracleGraph(n, src, dst, w)
    sources = np.arange(0, n, 11, dtype=np.uint32)
    od, _ = o.sssp(sources, n_threads=8)
    gpu.set_option("sssp.frontier", 1)
    try:
        gd, gp, _ = g.sssp(sources)
        assert np.array_equal(gd, od)
        _check_tree(src, dst, w, sources, gd, gp, n)
        gc, _ = g.closeness()
        oc = o.closeness(n_threads=8)
        fin = np.isfinite(oc)
        assert np.array_equal(np.isfinite(gc), fin) and np.allclose(gc[fin], oc[fin], rtol=1e-5)
    finally:
        gpu.set_option("sssp.frontier", 0)
    n, s2, d2 = rmat_edges(16, 8, 4242)                       # 65536 nodes: 10 bytes/node > shared memory
    w2 = (np.random.default_rng(1).integers(1, 64, s2.size) / 8.0).astype(np.float32)
    g2 = gpu.Graph(n, s2, d2, w2)
    o2 = O.OracleGraph(n, s2, d2, w2)
    srcs = np.array([0, 1, 77, 4097, 65535], np.uint32)
    od, _ = o2.sssp(srcs, n_threads=8)
    gpu.set_option("sssp.frontier", 1)
    try:
        gd, gp, _ = g2.sssp(srcs)
    finally:
        gpu.set_option("sssp.frontier", 0)
    assert np.array_equal(gd, od)
    # few sources on a large graph: the "wide" form (many CTAs per source, one launch per round)
    gpu.set_option("sssp.wide", 1)
    try:
        gw, pw, _ = g2.sssp(srcs[:3])
        assert np.array_equal(gw, od[:3])
        _check_tree(s2, d2, w2, srcs[:1], gw[:1], pw[:1], n)
        g1 = gpu.Graph(500, src, dst, w)
        assert np.array_equal(g1.sssp(sources)[0], O.OracleGraph(500, src, dst, w).sssp(sources, n_threads=8)[0])
    finally:
        gpu.set_option("sssp.wide", 0)


def test_betweenness_is_deterministic_and_rejects_zero_weight_cycles(gpu):
    n = 300
    src, dst, w = _random_graph(n, 1800, 13, dyadic=True)
    g = gpu.Graph(n, src, dst, w)
    a, _ = g.betweenness()
    b, _ = g.betweenness()
    assert np.array_equal(a, b)                 # per-source dependencies are added in source order, no atomics
    # a zero-weight edge inside a DAG of ties is fine ...
    src = np.array([0, 1, 0, 2], np.uint32)
    dst = np.array([1, 3, 2, 3], np.uint32)
    w = np.array([0, 1, 0, 1], np.float32)
    gb, _ = gpu.Graph(4, src, dst, w).betweenness()
    ob = O.OracleGraph(4, src, dst, w).betweenness()
    assert np.allclose(gb, ob)
    # ... a zero-weight CYCLE that does not pass through the source makes the tied paths unbounded: refused loudly
    # (the reference's enumeration would not terminate).  0 -> 1 (1.0), 1 <-> 2 at weight 0, 2 -> 3 (1.0)
    src = np.array([0, 1, 2, 2], np.uint32)
    dst = np.array([1, 2, 1, 3], np.uint32)
    w = np.array([1, 0, 0, 1], np.float32)
    with pytest.raises(gpu.CozoGpuError) as e:
        gpu.Graph(4, src, dst, w).betweenness()
    assert e.value.code == gpu.E_UNSUP
    gd, _, _ = gpu.Graph(4, src, dst, w).sssp([0])           # distances are still well defined
    assert gd[0].tolist() == [0, 1, 1, 2]


def test_sssp_paths_in_the_queue_and_wide_forms(gpu):
    rng = np.random.default_rng(8)
    n = 200
    pairs = sorted({(int(a), int(b)) for a, b in zip(rng.integers(0, n, 1600), rng.integers(0, n, 1600)) if a != b})
    src = np.array([p[0] for p in pairs], np.uint32)
    dst = np.array([p[1] for p in pairs], np.uint32)
    w = (rng.random(src.size) * 10 + 0.5).astype(np.float32)
    g = gpu.Graph(n, src, dst, w)
    sources = rng.integers(0, n, 40).astype(np.uint32)
    goals = rng.integers(0, n, 40).astype(np.uint32)
    fn = [[int(x) for x in rng.integers(0, n, rng.integers(0, 4)) if x != sources[i]] for i in range(40)]
    fe = [[pairs[int(j)] for j in rng.integers(0, len(pairs), rng.integers(0, 5))] for i in range(40)]
    res, _ = g.sssp_paths(sources, goals, fn, fe, max_len=64)
    for opt in ("sssp.frontier", "sssp.wide"):                       # the other two frontier forms: same answers
        gpu.set_option(opt, 1)
        try:
            assert g.sssp_paths(sources, goals, fn, fe, max_len=64)[0] == res
        finally:
            gpu.set_option(opt, 0)


@pytest.mark.parametrize("keep_pruned,extend", [(False, False), (True, False), (False, True), (True, True)])
def test_builder_sequential_semantics_equal_the_reference(gpu, keep_pruned, extend):
    """With batches of ONE node the device builder has the reference's sequential visibility (hnsw.rs:155-375): every
    insert sees all earlier ones, a row receives one in-edge at a time.  Given the same levels, its search
    (ef_construction beam carried across layers), heuristic selection (470-538) and shrink (376-469) must then
    produce the reference's graph edge for edge — checked against the oracle's faithful builder."""
    n, dim, m = (1200 if extend else 2500), 32, 8
    X = uniform_vectors(n, dim, 5150)
    g = gpu.HnswIndex.build(X, m=m, ef_construction=40, keep_pruned_connections=keep_pruned, level_seed=77, max_batch=1,
                            extend_candidates=extend)           # extend_candidates (hnsw.rs:499-511) implies batches of one
    ni, rp, ci, ep = g.export_levels()
    level = np.zeros(n, np.int64)
    for L in range(1, len(rp)):
        level[ni[L]] = L
    ix = O.OracleHnsw.new(n, dim, m=m, ef_construction=40, keep_pruned_connections=keep_pruned, extend_candidates=extend)
    for i in range(n):
        ix.insert(i, X[i], forced_level=-int(level[i]))
    lv = ix.levels()
    assert lv.entry == ep and lv.n_levels == len(rp)
    same = total = 0
    for L in range(len(rp)):
        nodes_d = np.arange(n) if L == 0 else ni[L]
        nodes_o = np.arange(n) if L == 0 else lv.node_ids[L]
        assert np.array_equal(nodes_d, nodes_o)
        for r in range(len(nodes_d)):
            a = set(ci[L][int(rp[L][r]):int(rp[L][r + 1])].tolist())
            b = set(lv.col_idx[L][int(lv.row_ptr[L][r]):int(lv.row_ptr[L][r + 1])].tolist())
            same += a == b
            total += 1
    # distances differ in the last f32 bits (summation order), which can flip a strict comparison of the heuristic
    # once in a long while and then propagate; anything systematic would show as a large mismatch
    assert same / total >= 0.995, (same, total)


