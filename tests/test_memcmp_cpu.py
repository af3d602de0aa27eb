"""The key codec against the reference's OWN tests (cozo-core/src/data/tests/memcmp.rs): same inputs,
same properties — round trips and `byte order == value order`.  This is the one piece of the host
layer whose parity is pinned by reference tests."""
import math
import random

import numpy as np
import pytest

from tests.hostmod import load

I64_MAX, I64_MIN = 2**63 - 1, -2**63


@pytest.fixture(scope="module")
def h():
    return load()


def _num_key(x):
    """Num order (data/value.rs:575-598): by f64 value (total order), an Int before the equal Float"""
    return x


def test_encode_decode_num(h):
    """memcmp.rs tests: encode_decode_num (lines 15-50)"""
    collected = []

    def test_num(v):
        enc = h.memcmp_encode_value(v)[1:]                      # strip NUM_TAG -> encode_num bytes
        dec, used = h.memcmp_decode_value(b"\x05" + enc)
        assert used == len(enc) + 1
        if isinstance(v, float) and math.isnan(v):
            assert isinstance(dec, float) and math.isnan(dec)
        else:
            assert dec == v and type(dec) is type(v)
        collected.append((enc, v))

    for i in range(54):
        for j in range(0, 1000, 7):                             # the reference uses every j; a stride keeps this quick
            vb = (I64_MAX >> i) - j
            for v in (vb, -vb - 1):
                test_num(v)
    test_num(float("inf"))
    test_num(float("-inf"))
    test_num(float("nan"))
    rng = random.Random(0)
    for _ in range(20000):
        f = (rng.random() - 0.5) * 2.0
        if f != 0.0:
            test_num(f)
            test_num(1.0 / f)
    by_bytes = [v for _, v in sorted(collected, key=lambda t: t[0])]
    import functools
    by_value = sorted((v for _, v in collected), key=functools.cmp_to_key(h.cmp))
    # identical sequences up to elements that compare equal
    assert len(by_bytes) == len(by_value)
    assert all(h.cmp(a, b) == 0 for a, b in zip(by_bytes, by_value))


def test_encode_decode_bytes(h):
    """memcmp.rs tests: encode_decode_bytes (lines 64-96)"""
    target = b"Lorem ipsum dolor sit amet, consectetur adipiscing elit..."
    for i in range(len(target)):
        bs = target[i:]
        enc = h.memcmp_encode_bytes(bs)
        dec, used = h.memcmp_decode_bytes(enc)
        assert used == len(enc) and dec == bs
        enc = h.memcmp_encode_bytes(target) + h.memcmp_encode_bytes(bs) + h.memcmp_encode_bytes(bs) + \
            h.memcmp_encode_bytes(target)
        off = 0
        for exp in (target, bs, bs, target):
            dec, used = h.memcmp_decode_bytes(enc[off:])
            assert dec == exp
            off += used
        assert off == len(enc)


def test_specific_encode(h):
    """memcmp.rs tests: specific_encode (lines 98-111): 2095 then "MSS" back to back"""
    enc = h.memcmp_encode_value(2095) + h.memcmp_encode_value("MSS")
    a, used = h.memcmp_decode_value(enc)
    b, used2 = h.memcmp_decode_value(enc[used:])
    assert used + used2 == len(enc) and a == 2095 and b == "MSS"


def test_encode_decode_datavalues(h):
    """memcmp.rs tests: encode_decode_datavalues (lines 113-137)"""
    dv = [None, False, True, 1, 1.0, I64_MAX, I64_MAX - 1, I64_MAX - 2, I64_MIN, I64_MIN + 1, I64_MIN + 2,
          float("inf"), float("-inf"), []]
    dv.append(list(dv))
    dv.append(list(dv))
    enc = h.memcmp_encode_value(dv)
    dec, used = h.memcmp_decode_value(enc)
    assert used == len(enc)

    def same(a, b):
        if isinstance(a, list):
            return isinstance(b, list) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        return type(a) is type(b) and a == b
    assert same(dec, dv)


def test_known_bytes(h):
    """byte-level anchors derived from the constants of memcmp.rs:20-42, 147-164, 208-215"""
    assert h.memcmp_encode_value(None) == b"\x01" and h.memcmp_encode_value(False) == b"\x02"
    assert h.memcmp_encode_value(True) == b"\x03"
    # "MSS": 3 bytes + 5 bytes of zero padding + marker 0xFF-5
    assert h.memcmp_encode_value("MSS") == b"\x06MSS\x00\x00\x00\x00\x00\xfa"
    # a full group is followed by 0xFF and then by an all-padding group with marker 0xFF-8 (the `index <= len` loop)
    assert h.memcmp_encode_value(b"12345678") == b"\x0712345678\xff" + b"\x00" * 8 + b"\xf7"
    # 1 -> order-encoded f64 bits of 1.0 with the sign bit set, then IS_EXACT_INT; 1.0 -> same bits, IS_FLOAT
    assert h.memcmp_encode_value(1) == b"\x05\xbf\xf0" + b"\x00" * 6 + b"\x00"
    assert h.memcmp_encode_value(1.0) == b"\x05\xbf\xf0" + b"\x00" * 6 + b"\x10"
    # negative floats: all bits flipped
    assert h.memcmp_encode_value(-1.0) == b"\x05\x40\x0f" + b"\xff" * 6 + b"\x10"
    # |i| >= 2^53: IS_APPROX_INT plus the order-encoded i64
    assert h.memcmp_encode_value(2**62) == b"\x05\xc3\xd0" + b"\x00" * 6 + b"\x04\xc0" + b"\x00" * 7
    # a vector key: tag 4, subtype 1, u64 BE length, BE f32 elements (memcmp.rs:51-69)
    assert h.memcmp_encode_value(np.array([1.0, -2.0], np.float32)) == \
        b"\x04\x01" + (2).to_bytes(8, "big") + b"\x3f\x80\x00\x00" + b"\xc0\x00\x00\x00"


def test_tuple_keys_sort_like_tuples(h):
    """the property the storage layer relies on: memcmp order of encoded keys == order of the tuples,
    and the 8-byte relation id prefix (tuple.rs:29-52, relation.rs:63-71)"""
    import functools
    rng = random.Random(5)
    pool = [None, False, True, -3, 0, 7, 2**60, -2.5, 0.0, 7.0, 1e300, "", "a", "ab", "b", "long string " * 3, b"", b"\x00",
            b"\x00\x01", [], [1], [1, "x"], [[None]]]
    tuples = [[rng.choice(pool) for _ in range(rng.randint(1, 4))] for _ in range(600)]
    enc = [h.memcmp_encode_key(t, 42) for t in tuples]
    for t, e in zip(tuples, enc):
        assert e[:8] == (42).to_bytes(8, "big")
        dec = h.memcmp_decode_key(e)
        assert h.cmp(dec, t) == 0
    by_bytes = [t for _, t in sorted(zip(enc, tuples), key=lambda p: p[0])]
    by_value = sorted(tuples, key=functools.cmp_to_key(h.cmp))
    assert all(h.cmp(a, b) == 0 for a, b in zip(by_bytes, by_value))
    # an index-relation key (layer, fr_k, fr__field, fr__sub_idx, to_k, to__field, to__sub_idx) round-trips
    key = [-2, "doc-17", 1, -1, "doc-99", 1, 3]
    assert h.memcmp_decode_key(h.memcmp_encode_key(key, 7)) == key
