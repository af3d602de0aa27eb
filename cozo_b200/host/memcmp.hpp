// memcmp.hpp — the order-preserving key codec of cozo's storage layer (data/memcmp.rs:20-371,
// data/tuple.rs:22-86), for the value kinds the hot path meets in the keys of `rel:idx` and of the
// base relation: Null, Bool, Num, Str, Bytes, List, Vec(F32).  This is the KEY half of SURVEY §8f
// rank 1 ("read the index & base relations straight from KV bytes"); the VALUE half (rmp-serde
// msgpack, un-vendored, parity unpinned) is msgpack.hpp.
// Unlike the rest of the host layer this piece is PINNED: data/tests/memcmp.rs's own property
// tests (round trips, byte order == value order) are reproduced in tests/test_memcmp_cpu.py.
#pragma once
#include <cstring>

#include "data_value.hpp"

namespace cozo_host {
namespace memcmp_codec {

constexpr uint8_t INIT_TAG = 0x00, NULL_TAG = 0x01, FALSE_TAG = 0x02, TRUE_TAG = 0x03, VEC_TAG = 0x04, NUM_TAG = 0x05,
                  STR_TAG = 0x06, BYTES_TAG = 0x07, LIST_TAG = 0x0A, BOT_TAG = 0xFF;  // memcmp.rs:20-34
constexpr uint8_t VEC_F32 = 0x01, VEC_F64 = 0x02;                                      // memcmp.rs:36-37
constexpr uint8_t IS_FLOAT = 0b00010000, IS_APPROX_INT = 0b00000100, IS_EXACT_INT = 0;  // memcmp.rs:39-41
constexpr int64_t EXACT_INT_BOUND = 0x20000000000000LL;                                 // memcmp.rs:42
constexpr size_t ENC_GROUP_SIZE = 8;                                                    // memcmp.rs:226
constexpr uint8_t ENC_MARKER = 0xFF;
constexpr uint64_t SIGN_MARK = 0x8000000000000000ULL;
constexpr size_t ENCODED_KEY_MIN_LEN = 8;  // tuple.rs:86: the relation id prefix

inline void put_u64_be(std::string& o, uint64_t v) {
  for (int i = 7; i >= 0; --i) o.push_back((char)((v >> (8 * i)) & 0xff));
}
inline uint64_t get_u64_be(const uint8_t* p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
  return v;
}
inline uint64_t order_encode_i64(int64_t v) { return (uint64_t)v ^ SIGN_MARK; }  // memcmp.rs:200-202
inline int64_t order_decode_i64(uint64_t u) { return (int64_t)(u ^ SIGN_MARK); }
inline uint64_t order_encode_f64(double v) {  // memcmp.rs:208-215
  uint64_t u;
  std::memcpy(&u, &v, 8);
  return std::signbit(v) ? ~u : (u | SIGN_MARK);
}
inline double order_decode_f64(uint64_t u) {  // memcmp.rs:217-224
  u = (u & SIGN_MARK) ? (u & ~SIGN_MARK) : ~u;
  double v;
  std::memcpy(&v, &u, 8);
  return v;
}

// encode_bytes (memcmp.rs:147-164): groups of 8 bytes, each followed by 0xFF - (padding length)
inline void encode_bytes(std::string& o, const std::string& key) {
  const size_t len = key.size();
  for (size_t index = 0; index <= len; index += ENC_GROUP_SIZE) {
    const size_t remain = len - index;
    size_t pad = 0;
    if (remain > ENC_GROUP_SIZE) {
      o.append(key, index, ENC_GROUP_SIZE);
    } else {
      pad = ENC_GROUP_SIZE - remain;
      o.append(key, index, remain);
      o.append(pad, '\0');
    }
    o.push_back((char)(ENC_MARKER - (uint8_t)pad));
  }
}
// decode_bytes (memcmp.rs:167-194); returns the number of input bytes consumed
inline size_t decode_bytes(const uint8_t* data, size_t n, std::string& key) {
  size_t offset = 0;
  for (;;) {
    if (offset + ENC_GROUP_SIZE + 1 > n) throw CozoError("", "truncated memcmp bytes");
    const uint8_t* chunk = data + offset;
    offset += ENC_GROUP_SIZE + 1;
    const size_t pad = (size_t)(ENC_MARKER - chunk[ENC_GROUP_SIZE]);
    if (pad == 0) {
      key.append(reinterpret_cast<const char*>(chunk), ENC_GROUP_SIZE);
      continue;
    }
    if (pad > ENC_GROUP_SIZE) throw CozoError("", "corrupt memcmp bytes");
    key.append(reinterpret_cast<const char*>(chunk), ENC_GROUP_SIZE - pad);
    return offset;
  }
}

// encode_num (memcmp.rs:126-145): order-encoded f64 first, then a tag that separates Int from Float
inline void encode_num(std::string& o, const DataValue& v) {
  double f = v.is_float ? v.f : (double)v.i;
  put_u64_be(o, order_encode_f64(f));
  if (v.is_float) {
    o.push_back((char)IS_FLOAT);
  } else if (v.i > -EXACT_INT_BOUND && v.i < EXACT_INT_BOUND) {
    o.push_back((char)IS_EXACT_INT);
  } else {
    o.push_back((char)IS_APPROX_INT);
    put_u64_be(o, order_encode_i64(v.i));
  }
}

// MemCmpEncoder::encode_datavalue (memcmp.rs:45-125)
inline void encode_datavalue(std::string& o, const DataValue& v) {
  switch (v.kind) {
    case DataValue::Null: o.push_back((char)NULL_TAG); break;
    case DataValue::Bool: o.push_back((char)(v.b ? TRUE_TAG : FALSE_TAG)); break;
    case DataValue::Vec: {  // memcmp.rs:51-69: big-endian elements
      o.push_back((char)VEC_TAG);
      o.push_back((char)VEC_F32);
      put_u64_be(o, v.v->size());
      for (float e : *v.v) {
        uint32_t u;
        std::memcpy(&u, &e, 4);
        for (int i = 3; i >= 0; --i) o.push_back((char)((u >> (8 * i)) & 0xff));
      }
      break;
    }
    case DataValue::Num:
      o.push_back((char)NUM_TAG);
      encode_num(o, v);
      break;
    case DataValue::Str:
      o.push_back((char)STR_TAG);
      encode_bytes(o, v.s);
      break;
    case DataValue::Bytes:
      o.push_back((char)BYTES_TAG);
      encode_bytes(o, v.s);
      break;
    case DataValue::List:
      o.push_back((char)LIST_TAG);
      for (auto& el : v.list) encode_datavalue(o, el);
      o.push_back((char)INIT_TAG);
      break;
    case DataValue::Bot: o.push_back((char)BOT_TAG); break;
  }
}

// DataValue::decode_from_key (memcmp.rs:257-368); returns bytes consumed
inline size_t decode_datavalue(const uint8_t* p, size_t n, DataValue& out) {
  if (n == 0) throw CozoError("", "truncated key");
  const uint8_t tag = p[0];
  switch (tag) {
    case NULL_TAG: out = DataValue::null(); return 1;
    case FALSE_TAG: out = DataValue::from_bool(false); return 1;
    case TRUE_TAG: out = DataValue::from_bool(true); return 1;
    case BOT_TAG: out = DataValue::bot(); return 1;
    case NUM_TAG: {  // Num::decode_from_key (memcmp.rs:229-254)
      if (n < 10) throw CozoError("", "truncated number");
      const double f = order_decode_f64(get_u64_be(p + 1));
      const uint8_t t = p[9];
      if (t == IS_FLOAT) {
        out = DataValue::from_float(f);
        return 10;
      }
      if (t == IS_EXACT_INT) {
        out = DataValue::from_int((int64_t)f);
        return 10;
      }
      if (t == IS_APPROX_INT) {
        if (n < 18) throw CozoError("", "truncated number");
        out = DataValue::from_int(order_decode_i64(get_u64_be(p + 10)));
        return 18;
      }
      throw CozoError("", "corrupt number tag");
    }
    case STR_TAG:
    case BYTES_TAG: {
      std::string s;
      size_t used = decode_bytes(p + 1, n - 1, s);
      out = tag == STR_TAG ? DataValue::from_str(std::move(s)) : DataValue::from_bytes(std::move(s));
      return 1 + used;
    }
    case LIST_TAG: {
      std::vector<DataValue> coll;
      size_t off = 1;
      for (;;) {
        if (off >= n) throw CozoError("", "truncated list");
        if (p[off] == INIT_TAG) break;
        DataValue el;
        off += decode_datavalue(p + off, n - off, el);
        coll.push_back(std::move(el));
      }
      out = DataValue::from_list(std::move(coll));
      return off + 1;
    }
    case VEC_TAG: {
      if (n < 10) throw CozoError("", "truncated vector");
      const uint8_t t = p[1];
      const uint64_t len = get_u64_be(p + 2);
      const size_t esz = t == VEC_F32 ? 4 : 8;
      if (t != VEC_F32 && t != VEC_F64) throw CozoError("", "corrupt vector tag");
      if (len > (n - 10) / esz) throw CozoError("", "truncated vector");  // overflow-safe: len is untrusted
      std::vector<float> v(len);
      for (uint64_t i = 0; i < len; ++i) {
        const uint8_t* e = p + 10 + i * esz;
        if (t == VEC_F32) {
          uint32_t u = ((uint32_t)e[0] << 24) | ((uint32_t)e[1] << 16) | ((uint32_t)e[2] << 8) | e[3];
          std::memcpy(&v[i], &u, 4);
        } else {  // F64 keys are narrowed: the device path is f32 (query vectors are cast the same way, hnsw.rs:883)
          uint64_t u = get_u64_be(e);
          double d;
          std::memcpy(&d, &u, 8);
          v[i] = (float)d;
        }
      }
      out = DataValue::from_vec(std::move(v));
      return 10 + len * esz;
    }
    default: throw CozoError("", "unsupported memcmp tag " + std::to_string((int)tag));
  }
}

// Encoded length of one value without materialising it (the stager walks index keys with this)
inline size_t skip_datavalue(const uint8_t* p, size_t n) {
  if (n == 0) throw CozoError("", "truncated key");
  switch (p[0]) {
    case NULL_TAG:
    case FALSE_TAG:
    case TRUE_TAG:
    case BOT_TAG: return 1;
    case NUM_TAG: {
      if (n < 10) throw CozoError("", "truncated number");
      const uint8_t t = p[9];
      if (t == IS_FLOAT || t == IS_EXACT_INT) return 10;
      if (t == IS_APPROX_INT) {
        if (n < 18) throw CozoError("", "truncated number");
        return 18;
      }
      throw CozoError("", "corrupt number tag");
    }
    case STR_TAG:
    case BYTES_TAG: {
      size_t off = 1;
      for (;;) {
        if (off + ENC_GROUP_SIZE + 1 > n) throw CozoError("", "truncated memcmp bytes");
        const uint8_t marker = p[off + ENC_GROUP_SIZE];
        off += ENC_GROUP_SIZE + 1;
        if (marker != ENC_MARKER) {
          if ((size_t)(ENC_MARKER - marker) > ENC_GROUP_SIZE) throw CozoError("", "corrupt memcmp bytes");
          return off;
        }
      }
    }
    case LIST_TAG: {
      size_t off = 1;
      for (;;) {
        if (off >= n) throw CozoError("", "truncated list");
        if (p[off] == INIT_TAG) return off + 1;
        off += skip_datavalue(p + off, n - off);
      }
    }
    case VEC_TAG: {
      if (n < 10) throw CozoError("", "truncated vector");
      if (p[1] != VEC_F32 && p[1] != VEC_F64) throw CozoError("", "corrupt vector tag");
      const uint64_t len = get_u64_be(p + 2);
      const size_t esz = p[1] == VEC_F32 ? 4 : 8;
      if (len > (n - 10) / esz) throw CozoError("", "truncated vector");  // overflow-safe: len is untrusted
      return 10 + (size_t)len * esz;
    }
    default: throw CozoError("", "unsupported memcmp tag " + std::to_string((int)p[0]));
  }
}

// TupleT::encode_as_key (tuple.rs:29-38): 8-byte big-endian relation id, then the values
inline std::string encode_as_key(const Tuple& t, uint64_t relation_id) {
  std::string o;
  put_u64_be(o, relation_id);
  for (auto& v : t) encode_datavalue(o, v);
  return o;
}
// decode_tuple_from_key (tuple.rs:41-52)
inline Tuple decode_tuple_from_key(const std::string& key) {
  if (key.size() < ENCODED_KEY_MIN_LEN) throw CozoError("", "key shorter than the relation id prefix");
  const uint8_t* p = reinterpret_cast<const uint8_t*>(key.data());
  size_t off = ENCODED_KEY_MIN_LEN;
  Tuple t;
  while (off < key.size()) {
    DataValue v;
    off += decode_datavalue(p + off, key.size() - off, v);
    t.push_back(std::move(v));
  }
  return t;
}

}  // namespace memcmp_codec
}  // namespace cozo_host
