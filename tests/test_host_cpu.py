"""CPU tests of the host layer that mirrors the reference's plugin interfaces (no GPU calls)."""
import pytest

from tests.hostmod import load


@pytest.fixture(scope="module")
def h():
    return load()


def test_custom_rule_contract(h):
    """the reference's own pinned test of the FixedRule plugin contract:
    runtime/tests.rs:530-577  `?[x] <~ SumCols(rel[], mult: 100)` over [[1,2,3,4],[5,6,7,8]] == [[1000],[2600]]"""
    db = h.Db()

    def sum_cols(inputs, options):
        mult = options.get("mult", 2)
        return [[sum(c if isinstance(c, int) else 0 for c in row) * mult] for row in inputs[0]]

    db.register_fixed_rule("SumCols", 1, sum_cols)
    assert db.run_fixed_rule("SumCols", [[[1, 2, 3, 4], [5, 6, 7, 8]]], {"mult": 100}, head_arity=1) == [[1000], [2600]]
    with pytest.raises(h.CozoError):        # names are unique (runtime/db.rs:764-775)
        db.register_fixed_rule("SumCols", 1, sum_cols)
    with pytest.raises(h.CozoError):        # built-ins cannot be removed (runtime/db.rs:779-784)
        db.unregister_fixed_rule("PageRank")
    assert db.unregister_fixed_rule("SumCols")
    with pytest.raises(h.CozoError) as e:
        db.run_fixed_rule("SumCols", [[[1]]])
    assert e.value.code == "parser::fixed_rule_not_found"


def test_head_arity_must_match(h):
    db = h.Db()
    with pytest.raises(h.CozoError) as e:   # PageRank has arity 2 (pagerank.rs:58-65)
        db.run_fixed_rule("PageRank", [[[1, 2]]], {}, head_arity=3)
    assert "arity" in str(e.value)


def test_option_validation_happens_before_any_device_work(h):
    db = h.Db()
    edges = [[[1, 2], [2, 3]]]
    with pytest.raises(h.CozoError) as e:   # unit_interval_option (mod.rs:492-504)
        db.run_fixed_rule("PageRank", edges, {"theta": 1.5})
    assert "between 0. and 1." in str(e.value)
    with pytest.raises(h.CozoError) as e:   # pos_integer_option (mod.rs:438-450)
        db.run_fixed_rule("PageRank", edges, {"iterations": 0})
    assert "positive integer" in str(e.value)
    with pytest.raises(h.CozoError) as e:   # bool_option (mod.rs:505-533)
        db.run_fixed_rule("PageRank", edges, {"undirected": 1})
    assert "boolean" in str(e.value)
    with pytest.raises(h.CozoError) as e:   # missing positional input (get_input)
        db.run_fixed_rule("ShortestPathDijkstra", edges, {})
    assert e.value.code == "algo::not_enough_args"


def test_value_order(h):
    # derive(Ord) variant order (value.rs:146-174) and Num order (value.rs:575-598)
    assert h.cmp(None, False) < 0 and h.cmp(True, 0) < 0 and h.cmp(10**9, "a") < 0 and h.cmp("z", [0]) < 0
    assert h.cmp(1, 1.0) < 0 and h.cmp(1.0, 1) > 0 and h.cmp(1, 1) == 0 and h.cmp(2, 1.5) > 0
    assert h.cmp([1, 2], [1, 2, 0]) < 0 and h.cmp("ab", "b") < 0


def test_vector_hash_is_sha256_of_le_bytes(h):
    """Vector::get_hash (data/value.rs:333-348) = SHA-256 over the little-endian f32 bytes; checked
    against hashlib (a public known-answer source) on empty, short and multi-block inputs."""
    import hashlib
    import numpy as np
    for n in (0, 1, 13, 14, 16, 128, 768):
        v = (np.arange(n, dtype=np.float32) * 0.37 - 3).astype("<f4")
        assert h.sha256_le_f32(v) == hashlib.sha256(v.tobytes()).digest()
