"""The three producers of a staging plan — from decoded tuples (`stage`), from KV bytes through tuples,
and from KV bytes directly (`stage_kv`, no index row is ever decoded into DataValues) — must hand
cozo_gpu_hnsw_stage the same arrays.  Host only: no device call is made here."""
import random

import numpy as np
import pytest

from oracle import oracle as O
from tests.hostmod import load
from tests.util import uniform_vectors


@pytest.fixture(scope="module")
def h():
    return load()


def _same_plan(a, b):
    assert a["keys"] == b["keys"] or all(
        x[1:] == y[1:] and len(x[0]) == len(y[0]) and all(_eq(p, q) for p, q in zip(x[0], y[0]))
        for x, y in zip(a["keys"], b["keys"])) and len(a["keys"]) == len(b["keys"])
    for k in ("entry", "n_levels", "edges_kept", "dropped_same_key", "dropped_ignore_link"):
        assert a[k] == b[k], k
    for k in ("node_ids", "row_ptr", "col_idx"):
        assert len(a[k]) == len(b[k])
        for x, y in zip(a[k], b[k]):
            assert np.array_equal(x, y), k
    assert np.array_equal(a["vectors"], b["vectors"])


def _eq(p, q):
    if isinstance(p, np.ndarray) or isinstance(q, np.ndarray):
        return np.array_equal(p, q)
    return p == q and type(p) is type(q)


def _rows(ix, key_of, field_idx, K, drop_some_self_loops=False):
    layer, fr, to, dist, ign = ix.relation_rows()
    rows = []
    for l, f, t, d, g in zip(layer, fr, to, dist, ign):
        self_loop = f == t
        rows.append([int(l)] + key_of(int(f)) + [field_idx, -1] + key_of(int(t)) + [field_idx, -1] +
                    [float(d), b"hash" if self_loop else None, bool(g)])
    rows.append([1] + [None] * (2 * K + 4) + [int(layer.min()), b"canary", False])
    return rows


@pytest.mark.parametrize("keys", ["str", "int_neg", "compound"])
def test_planners_agree_single_vector_rows(h, keys):
    n, dim, m = 600, 12, 5
    X = uniform_vectors(n, dim, 81)
    ix = O.OracleHnsw.new(n, dim, m=m, ef_construction=30)
    ix.insert_all(X)
    for gone in range(40, 55):                                 # removed vectors leave gaps in every layer
        ix.remove(gone)
    if keys == "str":
        K, cols, key_of = 1, ["k"], (lambda i: [f"key{i:05d}" + "x" * (i % 11)])       # lengths straddle the 8-byte groups
    elif keys == "int_neg":
        K, cols, key_of = 1, ["k"], (lambda i: [i * 7 - 2000])                           # negative and positive ints
    else:
        K, cols, key_of = 2, ["a", "b"], (lambda i: [i // 10, f"s{i % 10}"])             # two key columns
    base = h.Relation("a", cols, ["v", "tag", "note"])
    for i in range(n):
        base.put(key_of(i) + [X[i], i % 7, None if i % 3 else "n" * (i % 40)])
    rows = _rows(ix, key_of, K, K)
    mf = {"dim": dim, "m": m, "ef_construction": 30, "fields": [K]}
    p_tuples = h.plan_stage(base, rows, mf)
    base_kv = h.relation_to_kv(base, 11)
    idx_kv = h.rows_to_kv(rows, 2 * K + 5, 12)
    random.Random(3).shuffle(idx_kv)
    p_kv_tuples = h.plan_stage_kv(base_kv, 11, K, idx_kv, 12, mf, bytes_level=False)
    p_kv_bytes = h.plan_stage_kv(base_kv, 11, K, idx_kv, 12, mf, bytes_level=True)
    assert len(p_tuples["keys"]) > 0 and p_tuples["dropped_ignore_link"] >= 0
    _same_plan(p_tuples, p_kv_tuples)
    _same_plan(p_tuples, p_kv_bytes)
    # the plan is what the oracle holds: same adjacency as the oracle's level view
    lv = ix.levels()
    assert p_kv_bytes["n_levels"] == lv.n_levels


def test_planners_agree_multi_vector_rows_and_empty_index(h):
    dim, m = 8, 4
    X = uniform_vectors(300, dim, 82)
    ix = O.OracleHnsw.new(300, dim, m=m, ef_construction=20)
    ix.insert_all(X)
    base = h.Relation("docs", ["id"], ["title", "chunks"])
    for r in range(150):
        base.put([r, f"t{r}", [X[2 * r], X[2 * r + 1]]])
    layer, fr, to, dist, ign = ix.relation_rows()
    rows = [[int(l), int(f) // 2, 2, int(f) % 2, int(t) // 2, 2, int(t) % 2, float(d), None, bool(g)]
            for l, f, t, d, g in zip(layer, fr, to, dist, ign)]
    rows.append([1, None, None, None, None, None, None, int(layer.min()), b"c", False])
    mf = {"dim": dim, "m": m, "ef_construction": 20, "fields": [2]}
    base_kv, idx_kv = h.relation_to_kv(base, 5), h.rows_to_kv(rows, 7, 6)
    a = h.plan_stage(base, rows, mf)
    b = h.plan_stage_kv(base_kv, 5, 1, idx_kv, 6, mf, bytes_level=True)
    _same_plan(a, b)
    assert a["dropped_same_key"] > 300                      # sibling vectors of one row never link (hnsw.rs:609)
    # canary only
    canary = [[1, None, None, None, None, None, None, 0, b"c", False]]
    a = h.plan_stage(base, canary, mf)
    b = h.plan_stage_kv(base_kv, 5, 1, h.rows_to_kv(canary, 7, 6), 6, mf, bytes_level=True)
    _same_plan(a, b)
    assert a["entry"] == 0xFFFFFFFF and a["n_levels"] == 1 and len(a["keys"]) == 0
    # corrupt input is an error in both
    bad = [r for r in rows if not (r[0] == 0 and r[1] == 3)]          # no layer-0 rows FROM base row 3: edges to it dangle
    for bl in (False, True):
        with pytest.raises(h.CozoError):
            h.plan_stage_kv(base_kv, 5, 1, h.rows_to_kv(bad, 7, 6), 6, mf, bytes_level=bl)
    with pytest.raises(h.CozoError):
        h.plan_stage_kv(base_kv, 5, 1, [(k[:-2], v) for k, v in idx_kv], 6, mf, bytes_level=True)


def test_byte_level_planner_is_the_faster_one(h):
    a = h.bench_stage_kv(20000, 16, 16, False)
    b = h.bench_stage_kv(20000, 16, 16, True)
    assert a["index_rows"] == b["index_rows"] and a["edges_kept"] == b["edges_kept"] and a["vectors"] == 20000
    assert b["seconds"] < a["seconds"]
