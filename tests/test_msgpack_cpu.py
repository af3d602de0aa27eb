"""The value codec (cozo_b200/host/msgpack.hpp): stored-row values = 8-byte BE relation id ++
rmp-serde(Vec<DataValue>) (runtime/relation.rs:275-296, 526-531).

What is pinned here: the msgpack WIRE level, against the independent `msgpack` Python package
building the same tree.  What is not: the mapping of serde enums onto msgpack is rmp-serde 1.2.0's
published convention ({variant name: payload}); the crate is not vendored and the reference has no
byte-level test for values — parity unpinned, said so in the header of msgpack.hpp."""
import math
import random
import struct

import msgpack
import numpy as np
import pytest

from tests.hostmod import load


@pytest.fixture(scope="module")
def h():
    return load()


def tree(v):
    """the serde data-model tree of a DataValue (data/value.rs:146-174, 226-252, 493-499)"""
    if v is None:
        return "Null"
    if isinstance(v, bool):
        return {"Bool": v}
    if isinstance(v, int):
        return {"Num": {"Int": v}}
    if isinstance(v, float):
        return {"Num": {"Float": v}}
    if isinstance(v, str):
        return {"Str": v}
    if isinstance(v, bytes):
        return {"Bytes": v}
    if isinstance(v, np.ndarray):
        return {"Vec": [0, v.astype("<f4").tobytes()]}
    if isinstance(v, (list, tuple)):
        return {"List": [tree(e) for e in v]}
    raise TypeError(v)


def pack(t):
    return msgpack.packb(t, use_bin_type=True)


def same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.array_equal(np.asarray(a, np.float32), np.asarray(b, np.float32))
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    if isinstance(a, float) and math.isnan(a):
        return isinstance(b, float) and math.isnan(b)
    return a == b and type(a) is type(b)


VALUES = [None, True, False, 0, 1, -1, 127, 128, 255, 256, 65535, 65536, 2**32 - 1, 2**32, 2**63 - 1,
          -32, -33, -128, -129, -32768, -32769, -2**31, -2**31 - 1, -2**63,
          0.0, -0.0, 1.5, float("inf"), float("-inf"), float("nan"), 1e-310,
          "", "a", "x" * 31, "x" * 32, "x" * 255, "x" * 256, "x" * 70000, "héllo ✓",
          b"", b"\x00\xff", b"b" * 255, b"b" * 256, b"b" * 65536,
          [], [1, "two", 3.0, None, [True, [b"x"]]], list(range(20)),
          np.arange(5, dtype=np.float32), np.zeros(0, np.float32), np.random.default_rng(0).random(768, dtype=np.float32)]


def test_wire_format_matches_python_msgpack(h):
    for v in VALUES:
        assert h.msgpack_encode_value(v) == pack(tree(v)), repr(v)[:60]


def test_round_trip_and_skip(h):
    for v in VALUES:
        enc = h.msgpack_encode_value(v)
        dec, used = h.msgpack_decode_value(enc)
        assert used == len(enc)
        assert same(v, dec), repr(v)[:60]
        assert h.msgpack_skip(enc) == len(enc)
    rng = random.Random(1)
    for _ in range(3000):
        i = rng.getrandbits(rng.randint(1, 63)) * rng.choice((1, -1))
        enc = h.msgpack_encode_value(i)
        assert enc == pack(tree(i))
        assert h.msgpack_decode_value(enc) == (i, len(enc))
        f = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        enc = h.msgpack_encode_value(f)
        assert enc == pack(tree(f))
        assert same(h.msgpack_decode_value(enc)[0], f)


def test_decoder_accepts_index_variants_and_any_int_width(h):
    """rmp-serde's Deserializer takes a variant by name or by index; DataValue order value.rs:146-174"""
    assert h.msgpack_decode_value(pack({2: {0: 5}}))[0] == 5              # Num::Int by index
    assert h.msgpack_decode_value(pack({2: {1: 2.5}}))[0] == 2.5          # Num::Float
    assert h.msgpack_decode_value(pack(0))[0] is None                     # unit variant Null by index
    assert h.msgpack_decode_value(pack({1: True}))[0] is True
    assert h.msgpack_decode_value(pack({3: "s"}))[0] == "s"
    assert h.msgpack_decode_value(pack({7: [{2: {0: 1}}, "Null"]}))[0] == [1, None]
    wide = b"\x81\xa3Num\x81\xa3Int\xd3" + struct.pack(">q", 7)           # int64 form of a small value
    assert h.msgpack_decode_value(wide)[0] == 7
    f32 = b"\x81\xa3Num\x81\xa5Float\xca" + struct.pack(">f", 0.5)
    assert h.msgpack_decode_value(f32)[0] == 0.5
    v64 = pack({"Vec": [1, np.array([1.0, 2.5], "<f8").tobytes()]})       # F64 vectors narrow to f32
    assert np.array_equal(h.msgpack_decode_value(v64)[0], np.array([1.0, 2.5], np.float32))


def test_lenient_skip_of_kinds_outside_the_host_model(h):
    foreign = [{"Uuid": b"\x01" * 16}, {"Json": {"a": [1, 2, {"b": None}]}}, {"Validity": [[-5], [True]]},
               {"Set": [{"Num": {"Int": 1}}]}, {"Regex": "a+"}]
    for t in foreign:
        enc = pack(t)
        dec, used = h.msgpack_decode_value(enc)
        assert dec is None and used == len(enc)
        with pytest.raises(Exception):
            h.msgpack_decode_value(enc, lenient=False)
    for bad in (b"", b"\x81", b"\x81\xa3Num\x81\xa3Int", b"\x82\xa3Num\x00\x00\x00", b"\xc1"):
        with pytest.raises(Exception):
            h.msgpack_decode_value(bad)


def test_kv_row_round_trip_and_vector_extraction(h):
    rng = np.random.default_rng(3)
    rel_id = 0x0102030405
    for _ in range(50):
        v1, v2 = rng.random(16, dtype=np.float32), rng.random(16, dtype=np.float32)
        row = [int(rng.integers(-10**9, 10**9)), "k%d" % rng.integers(100), "payload", v1, [v2, v1], None, 3.25]
        key = h.memcmp_encode_key(row[:2], rel_id)
        val = h.encode_vals(row, 2, rel_id)
        assert val[:8] == rel_id.to_bytes(8, "big")
        assert val[8:] == pack([tree(x) for x in row[2:]])
        assert same(h.decode_tuple_from_kv(key, val), row)
        assert np.array_equal(h.extract_vector(val, 1, -1, 16), v1)          # value column 1 = row[3]
        assert np.array_equal(h.extract_vector(val, 2, 0, 16), v2)           # list-of-vectors column
        assert np.array_equal(h.extract_vector(val, 2, 1, 16), v1)
        with pytest.raises(Exception):
            h.extract_vector(val, 0, -1, 16)                                 # a Str column
        with pytest.raises(Exception):
            h.extract_vector(val, 1, -1, 8)                                  # wrong dimension
        with pytest.raises(Exception):
            h.extract_vector(val, 2, 2, 16)                                  # list index out of range
    # an index-relation value: (dist Float, hash Bytes | Null, ignore_link Bool)  relation.rs:1113-1126
    val = h.encode_vals([0, 1.25, b"\xaa" * 32, False], 1, 7)
    assert val[8:] == b"\x93" + pack({"Num": {"Float": 1.25}}) + pack({"Bytes": b"\xaa" * 32}) + pack({"Bool": False})
