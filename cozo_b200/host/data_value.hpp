// data_value.hpp — the slice of cozo's value model the hot path touches.
//
// Host language note: the reference host is Rust (no toolchain in this image), so the
// host side above the C ABI is C++ mirroring the reference's operator interface.  This
// file mirrors data/value.rs:146-174 (DataValue), 208-213 (Vector), 575-598 (Num order).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cozo_host {

// miette::Report stand-in: message + diagnostic code (e.g. "algo::not_an_edge")
struct CozoError : std::runtime_error {
  std::string code;
  CozoError(const std::string& code_, const std::string& msg) : std::runtime_error(msg), code(code_) {}
};

struct DataValue;
using Tuple = std::vector<DataValue>;

struct DataValue {
  // variant order == derive(Ord) order of the reference enum (value.rs:146-174)
  enum Kind : uint8_t { Null = 0, Bool = 1, Num = 2, Str = 3, Bytes = 4, List = 7, Vec = 9, Bot = 12 };
  Kind kind = Null;
  bool is_float = false;  // Num::Int / Num::Float
  bool b = false;
  int64_t i = 0;
  double f = 0.0;
  std::string s;                          // Str / Bytes
  std::vector<DataValue> list;            // List
  std::shared_ptr<std::vector<float>> v;  // Vec (Vector::F32); F64 vectors are cast on entry

  DataValue() = default;
  static DataValue null() { return DataValue(); }
  static DataValue bot() {
    DataValue d;
    d.kind = Bot;
    return d;
  }
  static DataValue from_bool(bool x) {
    DataValue d;
    d.kind = Bool;
    d.b = x;
    return d;
  }
  static DataValue from_int(int64_t x) {
    DataValue d;
    d.kind = Num;
    d.i = x;
    return d;
  }
  static DataValue from_float(double x) {
    DataValue d;
    d.kind = Num;
    d.is_float = true;
    d.f = x;
    return d;
  }
  static DataValue from_str(std::string x) {
    DataValue d;
    d.kind = Str;
    d.s = std::move(x);
    return d;
  }
  static DataValue from_bytes(std::string x) {
    DataValue d;
    d.kind = Bytes;
    d.s = std::move(x);
    return d;
  }
  static DataValue from_list(std::vector<DataValue> x) {
    DataValue d;
    d.kind = List;
    d.list = std::move(x);
    return d;
  }
  static DataValue from_vec(std::vector<float> x) {
    DataValue d;
    d.kind = Vec;
    d.v = std::make_shared<std::vector<float>>(std::move(x));
    return d;
  }

  // DataValue::get_int (value.rs): Int, or Float with integral value
  bool get_int(int64_t& out) const {
    if (kind != Num) return false;
    if (!is_float) {
      out = i;
      return true;
    }
    if (std::floor(f) == f && std::isfinite(f)) {
      out = (int64_t)f;
      return true;
    }
    return false;
  }
  bool get_float(double& out) const {
    if (kind != Num) return false;
    out = is_float ? f : (double)i;
    return true;
  }
  bool get_bool(bool& out) const {
    if (kind != Bool) return false;
    out = b;
    return true;
  }

  std::string repr() const {
    std::ostringstream o;
    switch (kind) {
      case Null: o << "null"; break;
      case Bool: o << (b ? "true" : "false"); break;
      case Num:
        if (is_float) o << f;
        else o << i;
        break;
      case Str: o << '"' << s << '"'; break;
      case Bytes: o << "bytes(" << s.size() << ")"; break;
      case List: {
        o << "[";
        for (size_t k = 0; k < list.size(); ++k) o << (k ? ", " : "") << list[k].repr();
        o << "]";
        break;
      }
      case Vec: o << "vec(" << (v ? v->size() : 0) << ")"; break;
      case Bot: o << "bot"; break;
    }
    return o.str();
  }
};

inline int total_cmp(double a, double b) {  // f64::total_cmp
  auto key = [](double x) {
    int64_t bits;
    static_assert(sizeof(bits) == sizeof(x), "");
    std::memcpy(&bits, &x, 8);
    bits ^= (int64_t)((uint64_t)(bits >> 63) >> 1);
    return bits;
  };
  int64_t ka = key(a), kb = key(b);
  return ka < kb ? -1 : (ka > kb ? 1 : 0);
}

inline int cmp(const DataValue& a, const DataValue& b);
inline int cmp_tuple(const Tuple& a, const Tuple& b) {
  size_t n = a.size() < b.size() ? a.size() : b.size();
  for (size_t k = 0; k < n; ++k) {
    int c = cmp(a[k], b[k]);
    if (c) return c;
  }
  return a.size() < b.size() ? -1 : (a.size() > b.size() ? 1 : 0);
}

inline int cmp(const DataValue& a, const DataValue& b) {
  if (a.kind != b.kind) return a.kind < b.kind ? -1 : 1;
  switch (a.kind) {
    case DataValue::Bool: return (int)a.b - (int)b.b;
    case DataValue::Num: {  // value.rs:575-598: an Int sorts before the equal Float
      if (!a.is_float && !b.is_float) return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
      if (a.is_float && b.is_float) return total_cmp(a.f, b.f);
      if (!a.is_float) {
        int c = total_cmp((double)a.i, b.f);
        return c == 0 ? -1 : c;
      }
      int c = total_cmp(a.f, (double)b.i);
      return c == 0 ? 1 : c;
    }
    case DataValue::Str:
    case DataValue::Bytes: return a.s < b.s ? -1 : (a.s > b.s ? 1 : 0);
    case DataValue::List: return cmp_tuple(a.list, b.list);
    case DataValue::Vec: {
      const auto &x = *a.v, &y = *b.v;
      size_t n = x.size() < y.size() ? x.size() : y.size();
      for (size_t k = 0; k < n; ++k) {
        int c = total_cmp(x[k], y[k]);
        if (c) return c;
      }
      return x.size() < y.size() ? -1 : (x.size() > y.size() ? 1 : 0);
    }
    default: return 0;
  }
}

struct DataValueLess {
  bool operator()(const DataValue& a, const DataValue& b) const { return cmp(a, b) < 0; }
};
struct TupleLess {
  bool operator()(const Tuple& a, const Tuple& b) const { return cmp_tuple(a, b) < 0; }
};
inline bool operator==(const DataValue& a, const DataValue& b) { return cmp(a, b) == 0; }

}  // namespace cozo_host
