// msgpack.hpp — the VALUE half of cozo's KV row format: a stored row's value is
//   8-byte big-endian relation id  ++  rmp-serde(Vec<DataValue>)       (runtime/relation.rs:275-296)
// and is read back by `rmp_serde::from_slice(&val[ENCODED_KEY_MIN_LEN..])` (relation.rs:528).
//
// PARITY UNPINNED.  The byte format comes from a third-party crate that is NOT vendored under
// /root/reference — rmp-serde 1.2.0 over rmp 0.8.14 (Cargo.lock:3213-3230), used with
// `Serializer::new` defaults (relation.rs:284,294) — and the reference has no test that pins value
// bytes.  This file restates the crate's published encoding of serde's data model:
//   * enum unit variant      -> the variant NAME as a msgpack str          ("Null", "Bot")
//   * enum newtype variant   -> fixmap(1) { variant NAME : payload }       ({"Num": {"Int": 3}})
//   * seq / tuple            -> array header + elements
//   * integers               -> the shortest msgpack form (rmp::encode::write_sint: non-negative
//                               values use the unsigned family)
//   * f64 -> 0xcb, bool -> 0xc2/0xc3, str -> fixstr/str8/16/32, bytes -> bin8/16/32
// DataValue's own serde shape is in the reference: derive on the enum (data/value.rs:146-174), Num
// derive (493-499), `#[serde(with = "serde_bytes")]` on Bytes (158-159), and Vector's hand-written
// impl = tuple(2) of (u8 tag 0=F32/1=F64, raw native-endian element bytes) (226-252).
// The msgpack wire level itself IS checked: tests/test_msgpack_cpu.py compares every encoding with
// the `msgpack` Python package building the same tree.  The DECODER is deliberately permissive —
// it accepts a variant given by name or by index and any integer width — so it reads the
// reference's bytes under either enum convention rmp-serde has used; only the ENCODER depends on
// the restated convention.
#pragma once
#include "data_value.hpp"

namespace cozo_host {
namespace msgpack_codec {

// ---- writer -------------------------------------------------------------------------------
inline void put_be(std::string& o, uint64_t v, int nbytes) {
  for (int i = nbytes - 1; i >= 0; --i) o.push_back((char)((v >> (8 * i)) & 0xff));
}
inline void write_uint(std::string& o, uint64_t v) {  // rmp::encode::write_uint
  if (v < 128) o.push_back((char)v);
  else if (v <= 0xff) { o.push_back((char)0xcc); put_be(o, v, 1); }
  else if (v <= 0xffff) { o.push_back((char)0xcd); put_be(o, v, 2); }
  else if (v <= 0xffffffffull) { o.push_back((char)0xce); put_be(o, v, 4); }
  else { o.push_back((char)0xcf); put_be(o, v, 8); }
}
inline void write_sint(std::string& o, int64_t v) {  // rmp::encode::write_sint
  if (v >= 0) return write_uint(o, (uint64_t)v);
  if (v >= -32) o.push_back((char)(uint8_t)v);
  else if (v >= -128) { o.push_back((char)0xd0); put_be(o, (uint64_t)v, 1); }
  else if (v >= -32768) { o.push_back((char)0xd1); put_be(o, (uint64_t)v, 2); }
  else if (v >= -2147483648LL) { o.push_back((char)0xd2); put_be(o, (uint64_t)v, 4); }
  else { o.push_back((char)0xd3); put_be(o, (uint64_t)v, 8); }
}
inline void write_f64(std::string& o, double f) {
  uint64_t u;
  std::memcpy(&u, &f, 8);
  o.push_back((char)0xcb);
  put_be(o, u, 8);
}
inline void write_str(std::string& o, const std::string& s) {
  const size_t n = s.size();
  if (n < 32) o.push_back((char)(0xa0 | n));
  else if (n <= 0xff) { o.push_back((char)0xd9); put_be(o, n, 1); }
  else if (n <= 0xffff) { o.push_back((char)0xda); put_be(o, n, 2); }
  else { o.push_back((char)0xdb); put_be(o, n, 4); }
  o.append(s);
}
inline void write_bin(std::string& o, const char* p, size_t n) {
  if (n <= 0xff) { o.push_back((char)0xc4); put_be(o, n, 1); }
  else if (n <= 0xffff) { o.push_back((char)0xc5); put_be(o, n, 2); }
  else { o.push_back((char)0xc6); put_be(o, n, 4); }
  o.append(p, n);
}
inline void write_array_len(std::string& o, size_t n) {
  if (n < 16) o.push_back((char)(0x90 | n));
  else if (n <= 0xffff) { o.push_back((char)0xdc); put_be(o, n, 2); }
  else { o.push_back((char)0xdd); put_be(o, n, 4); }
}
inline void write_variant(std::string& o, const char* name) {  // newtype variant header
  o.push_back((char)0x81);
  write_str(o, name);
}

inline void encode_datavalue(std::string& o, const DataValue& v) {
  switch (v.kind) {
    case DataValue::Null: write_str(o, "Null"); break;
    case DataValue::Bot: write_str(o, "Bot"); break;
    case DataValue::Bool:
      write_variant(o, "Bool");
      o.push_back((char)(v.b ? 0xc3 : 0xc2));
      break;
    case DataValue::Num:
      write_variant(o, "Num");
      if (v.is_float) {
        write_variant(o, "Float");
        write_f64(o, v.f);
      } else {
        write_variant(o, "Int");
        write_sint(o, v.i);
      }
      break;
    case DataValue::Str:
      write_variant(o, "Str");
      write_str(o, v.s);
      break;
    case DataValue::Bytes:
      write_variant(o, "Bytes");
      write_bin(o, v.s.data(), v.s.size());
      break;
    case DataValue::List:
      write_variant(o, "List");
      write_array_len(o, v.list.size());
      for (auto& e : v.list) encode_datavalue(o, e);
      break;
    case DataValue::Vec:  // value.rs:226-252: (0u8, raw f32 bytes in native = little-endian order)
      write_variant(o, "Vec");
      write_array_len(o, 2);
      write_uint(o, 0);
      write_bin(o, reinterpret_cast<const char*>(v.v->data()), v.v->size() * sizeof(float));
      break;
  }
}

// encode_val_for_store / encode_val_only_for_store (relation.rs:275-296)
inline std::string encode_vals(const Tuple& t, size_t start, uint64_t relation_id) {
  std::string o;
  put_be(o, relation_id, 8);
  write_array_len(o, t.size() - start);
  for (size_t i = start; i < t.size(); ++i) encode_datavalue(o, t[i]);
  return o;
}

// ---- reader -------------------------------------------------------------------------------
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  Reader(const uint8_t* b, size_t n) : p(b), end(b + n) {}
  void need(size_t n) const {
    if ((size_t)(end - p) < n) throw CozoError("", "truncated msgpack value");
  }
  uint8_t peek() const {
    need(1);
    return *p;
  }
  uint64_t be(int nbytes) {
    need((size_t)nbytes);
    uint64_t v = 0;
    for (int i = 0; i < nbytes; ++i) v = (v << 8) | *p++;
    return v;
  }
  bool is_int() const {
    const uint8_t t = peek();
    return t < 0x80 || t >= 0xe0 || (t >= 0xcc && t <= 0xd3);
  }
  int64_t read_int() {
    const uint8_t t = (uint8_t)be(1);
    if (t < 0x80) return t;
    if (t >= 0xe0) return (int8_t)t;
    switch (t) {
      case 0xcc: return (int64_t)be(1);
      case 0xcd: return (int64_t)be(2);
      case 0xce: return (int64_t)be(4);
      case 0xcf: return (int64_t)be(8);
      case 0xd0: return (int8_t)be(1);
      case 0xd1: return (int16_t)be(2);
      case 0xd2: return (int32_t)be(4);
      case 0xd3: return (int64_t)be(8);
    }
    throw CozoError("", "msgpack: expected an integer");
  }
  double read_float() {
    const uint8_t t = peek();
    if (t == 0xcb) {
      ++p;
      uint64_t u = be(8);
      double d;
      std::memcpy(&d, &u, 8);
      return d;
    }
    if (t == 0xca) {
      ++p;
      uint32_t u = (uint32_t)be(4);
      float f;
      std::memcpy(&f, &u, 4);
      return f;
    }
    return (double)read_int();
  }
  bool is_str() const {
    const uint8_t t = peek();
    return (t >= 0xa0 && t <= 0xbf) || (t >= 0xd9 && t <= 0xdb);
  }
  std::string read_str() {
    const uint8_t t = (uint8_t)be(1);
    size_t n;
    if (t >= 0xa0 && t <= 0xbf) n = t & 0x1f;
    else if (t == 0xd9) n = be(1);
    else if (t == 0xda) n = be(2);
    else if (t == 0xdb) n = be(4);
    else throw CozoError("", "msgpack: expected a string");
    need(n);
    std::string s(reinterpret_cast<const char*>(p), n);
    p += n;
    return s;
  }
  // bin, or (serde_bytes leniency) str
  std::pair<const uint8_t*, size_t> read_bin_view() {
    const uint8_t t = peek();
    size_t n;
    if (t == 0xc4) { ++p; n = be(1); }
    else if (t == 0xc5) { ++p; n = be(2); }
    else if (t == 0xc6) { ++p; n = be(4); }
    else if (is_str()) {
      ++p;
      if (t >= 0xa0 && t <= 0xbf) n = t & 0x1f;
      else n = be(t == 0xd9 ? 1 : t == 0xda ? 2 : 4);
    } else throw CozoError("", "msgpack: expected bytes");
    need(n);
    auto r = std::make_pair(p, n);
    p += n;
    return r;
  }
  size_t read_array_len() {
    const uint8_t t = (uint8_t)be(1);
    if (t >= 0x90 && t <= 0x9f) return t & 0x0f;
    if (t == 0xdc) return be(2);
    if (t == 0xdd) return be(4);
    throw CozoError("", "msgpack: expected an array");
  }
  size_t read_map_len() {
    const uint8_t t = (uint8_t)be(1);
    if (t >= 0x80 && t <= 0x8f) return t & 0x0f;
    if (t == 0xde) return be(2);
    if (t == 0xdf) return be(4);
    throw CozoError("", "msgpack: expected a map");
  }
  // skip one complete msgpack object of any type
  void skip() {
    const uint8_t t = (uint8_t)be(1);
    auto adv = [&](size_t n) {
      need(n);
      p += n;
    };
    if (t < 0x80 || t >= 0xe0) return;
    if (t <= 0x8f) { for (size_t i = 0, n = 2 * (size_t)(t & 0x0f); i < n; ++i) skip(); return; }
    if (t <= 0x9f) { for (size_t i = 0, n = t & 0x0f; i < n; ++i) skip(); return; }
    if (t <= 0xbf) return adv(t & 0x1f);
    switch (t) {
      case 0xc0: case 0xc2: case 0xc3: return;
      case 0xc4: return adv(be(1));
      case 0xc5: return adv(be(2));
      case 0xc6: return adv(be(4));
      case 0xc7: { size_t n = be(1); return adv(n + 1); }
      case 0xc8: { size_t n = be(2); return adv(n + 1); }
      case 0xc9: { size_t n = be(4); return adv(n + 1); }
      case 0xca: return adv(4);
      case 0xcb: return adv(8);
      case 0xcc: case 0xd0: return adv(1);
      case 0xcd: case 0xd1: return adv(2);
      case 0xce: case 0xd2: return adv(4);
      case 0xcf: case 0xd3: return adv(8);
      case 0xd4: return adv(2);
      case 0xd5: return adv(3);
      case 0xd6: return adv(5);
      case 0xd7: return adv(9);
      case 0xd8: return adv(17);
      case 0xd9: return adv(be(1));
      case 0xda: return adv(be(2));
      case 0xdb: return adv(be(4));
      case 0xdc: { for (size_t i = 0, n = be(2); i < n; ++i) skip(); return; }
      case 0xdd: { for (size_t i = 0, n = be(4); i < n; ++i) skip(); return; }
      case 0xde: { for (size_t i = 0, n = 2 * be(2); i < n; ++i) skip(); return; }
      case 0xdf: { for (size_t i = 0, n = 2 * be(4); i < n; ++i) skip(); return; }
    }
    throw CozoError("", "msgpack: reserved type byte");
  }
};

// variant order of the reference enums (value.rs:146-174, 493-499): the index form of a variant key
inline const char* const* datavalue_variants() {
  static const char* const names[] = {"Null", "Bool", "Num",  "Str", "Bytes",    "Uuid", "Regex",
                                      "List", "Set",  "Vec",  "Json", "Validity", "Bot"};
  return names;
}
inline std::string read_variant(Reader& r, const char* const* names, size_t n_names) {
  if (r.is_str()) return r.read_str();
  int64_t idx = r.read_int();
  if (idx < 0 || (size_t)idx >= n_names) throw CozoError("", "msgpack: variant index out of range");
  return names[idx];
}

// Decode one DataValue.  Kinds outside this host's value model (Uuid, Regex, Set, Json, Validity)
// are skipped and come back as Null when `lenient`, else raise: the stager only needs the key
// columns and the vector columns, but must be able to step over everything else in a base row.
inline DataValue decode_datavalue(Reader& r, bool lenient = true) {
  std::string var;
  bool has_payload = true;
  if (r.is_str() || r.is_int()) {  // unit variant
    var = read_variant(r, datavalue_variants(), 13);
    has_payload = false;
  } else {
    if (r.read_map_len() != 1) throw CozoError("", "msgpack: enum must be a 1-entry map");
    var = read_variant(r, datavalue_variants(), 13);
  }
  if (!has_payload) {
    if (var == "Null") return DataValue::null();
    if (var == "Bot") return DataValue::bot();
    throw CozoError("", "msgpack: variant " + var + " needs a payload");
  }
  if (var == "Bool") {
    const uint8_t t = (uint8_t)r.be(1);
    if (t != 0xc2 && t != 0xc3) throw CozoError("", "msgpack: expected a bool");
    return DataValue::from_bool(t == 0xc3);
  }
  if (var == "Num") {
    static const char* const num_names[] = {"Int", "Float"};
    if (r.read_map_len() != 1) throw CozoError("", "msgpack: Num must be a 1-entry map");
    std::string nv = read_variant(r, num_names, 2);
    if (nv == "Int") return DataValue::from_int(r.read_int());
    if (nv == "Float") return DataValue::from_float(r.read_float());
    throw CozoError("", "msgpack: unknown Num variant " + nv);
  }
  if (var == "Str") return DataValue::from_str(r.read_str());
  if (var == "Bytes") {
    auto b = r.read_bin_view();
    return DataValue::from_bytes(std::string(reinterpret_cast<const char*>(b.first), b.second));
  }
  if (var == "List") {
    const size_t n = r.read_array_len();
    std::vector<DataValue> l;
    l.reserve(n);
    for (size_t i = 0; i < n; ++i) l.push_back(decode_datavalue(r, lenient));
    return DataValue::from_list(std::move(l));
  }
  if (var == "Vec") {  // VectorVisitor (value.rs:263-308)
    if (r.read_array_len() != 2) throw CozoError("", "msgpack: vector representation");
    const int64_t tag = r.read_int();
    auto b = r.read_bin_view();
    if (tag == 0) {
      std::vector<float> v(b.second / sizeof(float));
      std::memcpy(v.data(), b.first, v.size() * sizeof(float));
      return DataValue::from_vec(std::move(v));
    }
    if (tag == 1) {  // F64 narrowed to the device's f32 (hnsw.rs:883 casts queries the same way)
      const size_t n = b.second / sizeof(double);
      std::vector<float> v(n);
      for (size_t i = 0; i < n; ++i) {
        double d;
        std::memcpy(&d, b.first + 8 * i, 8);
        v[i] = (float)d;
      }
      return DataValue::from_vec(std::move(v));
    }
    throw CozoError("", "msgpack: bad vector tag");
  }
  if (var == "Bot") return DataValue::bot();
  if (var == "Null") return DataValue::null();
  if (!lenient) throw CozoError("", "msgpack: DataValue::" + var + " is outside the host value model");
  r.skip();
  return DataValue::null();
}

// extend_tuple_from_v (relation.rs:526-531): append the value columns to an already decoded key
inline void extend_tuple_from_v(Tuple& key, const std::string& val, bool lenient = true) {
  if (val.empty()) return;
  if (val.size() < 8) throw CozoError("", "value shorter than the relation id prefix");
  Reader r(reinterpret_cast<const uint8_t*>(val.data()) + 8, val.size() - 8);
  const size_t n = r.read_array_len();
  for (size_t i = 0; i < n; ++i) key.push_back(decode_datavalue(r, lenient));
  if (r.p != r.end) throw CozoError("", "trailing bytes after the value columns");
}

// One Bool value column of a stored value without materialising the row (the stager reads `ignore_link`
// with it).  A column that is not a Bool reads as false, like DataValue::get_bool on the decoded tuple.
inline bool read_bool_column(const std::string& val, size_t col) {
  if (val.size() < 8) throw CozoError("", "value shorter than the relation id prefix");
  Reader r(reinterpret_cast<const uint8_t*>(val.data()) + 8, val.size() - 8);
  const size_t n = r.read_array_len();
  if (col >= n) throw CozoError("", "value column out of range");
  for (size_t i = 0; i < col; ++i) r.skip();
  const uint8_t t = r.peek();
  if (!((t >= 0x80 && t <= 0x8f) || t == 0xde || t == 0xdf)) return false;  // a unit variant (Null, Bot)
  if (r.read_map_len() != 1) throw CozoError("", "msgpack: enum must be a 1-entry map");
  if (read_variant(r, datavalue_variants(), 13) != "Bool") return false;
  const uint8_t b = (uint8_t)r.be(1);
  if (b != 0xc2 && b != 0xc3) throw CozoError("", "msgpack: expected a bool");
  return b == 0xc3;
}

// Fast path of the stager: copy ONE f32 vector column (and optionally one element of a
// list-of-vectors column, sub_idx >= 0) of a stored value into `out[dim]` without materialising the
// row — VectorCache::ensure_key (hnsw.rs:122-151) against raw KV bytes.  `col` counts value columns.
inline void extract_vector(const std::string& val, size_t col, int32_t sub_idx, float* out, size_t dim) {
  if (val.size() < 8) throw CozoError("", "value shorter than the relation id prefix");
  Reader r(reinterpret_cast<const uint8_t*>(val.data()) + 8, val.size() - 8);
  const size_t n = r.read_array_len();
  if (col >= n) throw CozoError("", "vector column out of range");
  for (size_t i = 0; i < col; ++i) r.skip();
  auto open_variant = [&](const char* want) {
    if (r.read_map_len() != 1) throw CozoError("", std::string("Cannot interpret value as ") + want);
    std::string var = read_variant(r, datavalue_variants(), 13);
    if (var != want) throw CozoError("", "Cannot interpret " + var + " as " + want);
  };
  if (sub_idx >= 0) {
    open_variant("List");
    const size_t ln = r.read_array_len();
    if ((size_t)sub_idx >= ln) throw CozoError("", "list index out of range");
    for (int32_t i = 0; i < sub_idx; ++i) r.skip();
  }
  open_variant("Vec");
  if (r.read_array_len() != 2) throw CozoError("", "msgpack: vector representation");
  const int64_t tag = r.read_int();
  auto b = r.read_bin_view();
  if (tag == 0) {
    if (b.second != dim * sizeof(float)) throw CozoError("", "vector dimension mismatch for the index");
    std::memcpy(out, b.first, b.second);
  } else if (tag == 1) {
    if (b.second != dim * sizeof(double)) throw CozoError("", "vector dimension mismatch for the index");
    for (size_t i = 0; i < dim; ++i) {
      double d;
      std::memcpy(&d, b.first + 8 * i, 8);
      out[i] = (float)d;
    }
  } else {
    throw CozoError("", "msgpack: bad vector tag");
  }
}

}  // namespace msgpack_codec
}  // namespace cozo_host
