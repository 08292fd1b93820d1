// Minimal JSON reader for the reference's wire format (src/Types.hs aeson instances).
//
// One pass over the text builds a flat TAPE of 24-byte nodes (no per-value allocation): numbers and strings
// are (pointer, length) views INTO the parsed buffer, arrays/objects carry their child count and the index of
// the node after their subtree.  `JValue` is a (tape, index) handle with the accessors the decoders need.
// Numbers are kept as exact decimal tokens and reduced mod p on demand: field elements must never go through
// `double`.  A 335 KB standard-recursion proof (16 k numbers) decodes in well under a
// millisecond per core; p2v_parse_proofs spreads a batch over host threads.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace p2vhost {

struct JsonError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// eight ASCII digits at once (the SWAR trick of fast_float / simdjson)
inline bool eightDigits(const char *p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return ((v & 0xF0F0F0F0F0F0F0F0ULL) | (((v + 0x0606060606060606ULL) & 0xF0F0F0F0F0F0F0F0ULL) >> 4)) == 0x3333333333333333ULL;
}
inline uint32_t parseEightDigits(const char *p) {  // little-endian load: p[0] is the most significant digit
  uint64_t v;
  memcpy(&v, p, 8);
  v -= 0x3030303030303030ULL;
  v = (v * 10) + (v >> 8);
  v = (((v & 0x000000FF000000FFULL) * 0x000F424000000064ULL) + (((v >> 16) & 0x000000FF000000FFULL) * 0x0000271000000001ULL)) >> 32;
  return (uint32_t)v;
}

struct JNode {
  enum Kind : uint8_t { Null, False, True, Number, String, Array, Object, Key };
  const char *p = nullptr;  // Number/String/Key: the token (String/Key: between the quotes, escapes undecoded)
  uint32_t len = 0;         // token length; Array/Object: number of children (Object: key/value PAIRS)
  uint32_t next = 0;        // index of the node following this value's subtree
  uint8_t kind = Null;
  uint8_t escaped = 0;      // String/Key: contains a backslash
};

class JList;

class JValue {
 public:
  // kept for source compatibility with code that switches on the kind
  enum Kind { Null, Bool, Number, String, Array, Object };

  JValue() = default;
  JValue(const std::vector<JNode> *tape, uint32_t idx) : tape_(tape), idx_(idx) {}
  bool valid() const { return tape_ != nullptr; }
  Kind kind() const {
    switch (node().kind) {
      case JNode::Null: return Null;
      case JNode::False: case JNode::True: return Bool;
      case JNode::Number: return Number;
      case JNode::String: return String;
      case JNode::Array: return Array;
      default: return Object;
    }
  }
  bool isObject() const { return node().kind == JNode::Object; }
  size_t objSize() const { return isObject() ? node().len : 0; }
  // i-th (key, value) of an object, in document order
  std::string objKey(size_t i) const { return decode((*tape_)[memberIndex(i)]); }
  JValue objVal(size_t i) const { return JValue(tape_, memberIndex(i) + 1); }

  // first member with this key, or an invalid handle
  JValue find(const char *key) const {
    if (!isObject()) return JValue();
    size_t klen = strlen(key);
    uint32_t i = idx_ + 1;
    for (uint32_t m = 0; m < node().len; m++) {
      const JNode &k = (*tape_)[i];
      bool eq = k.escaped ? decode(k) == key : (k.len == klen && memcmp(k.p, key, klen) == 0);
      if (eq) return JValue(tape_, i + 1);
      i = (*tape_)[i + 1].next;
    }
    return JValue();
  }
  JValue at(const char *key) const {
    JValue v = find(key);
    if (!v.valid()) throw JsonError(std::string("missing key \"") + key + "\"");
    return v;
  }
  JValue find(const std::string &key) const { return find(key.c_str()); }
  JValue at(const std::string &key) const { return at(key.c_str()); }
  inline JList list() const;

  std::string token() const { return std::string(node().p, node().len); }
  // integer token -> value mod p (p = 2^64 - 2^32 + 1), like `mkGoldilocks <$> parseJSON`.  Exact for any length:
  // up to 19 digits accumulate in a u64, the 19 before them in a second u64 (value = hi * 10^k + lo < 2^128, one
  // reduction); longer tokens fall back to digit-by-digit reduction.
  uint64_t felt() const {
    const JNode &nd = node();
    if (nd.kind != JNode::Number) throw JsonError("expected a number");
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    const char *t = nd.p, *e = nd.p + nd.len;
    bool negative = false;
    if (t < e && *t == '-') { negative = true; t++; }
    if (t >= e) throw JsonError("bad number");
    static const uint64_t POW10[20] = {1ULL, 10ULL, 100ULL, 1000ULL, 10000ULL, 100000ULL, 1000000ULL, 10000000ULL, 100000000ULL,
                                       1000000000ULL, 10000000000ULL, 100000000000ULL, 1000000000000ULL, 10000000000000ULL,
                                       100000000000000ULL, 1000000000000000ULL, 10000000000000000ULL, 100000000000000000ULL,
                                       1000000000000000000ULL, 10000000000000000000ULL};
    size_t ndig = (size_t)(e - t);
    uint64_t v;
    if (ndig <= 38) {
      size_t n1 = ndig > 19 ? ndig - 19 : 0;  // leading digits
      uint64_t hi = 0, lo = 0;
      size_t i = 0;
      for (; i < n1; i++) { unsigned d = (unsigned)(t[i] - '0'); if (d > 9) bad(); hi = hi * 10 + d; }
      for (; i + 8 <= ndig && eightDigits(t + i); i += 8) lo = lo * 100000000ULL + parseEightDigits(t + i);
      for (; i < ndig; i++) { unsigned d = (unsigned)(t[i] - '0'); if (d > 9) bad(); lo = lo * 10 + d; }
      if (n1 == 0) v = lo >= P ? lo - P : lo;  // lo < 10^19 < 2p
      else {
        unsigned __int128 full = (unsigned __int128)hi * POW10[ndig - n1] + lo;  // < 10^38 < 2^128
        v = reduce128(full);
      }
    } else {
      unsigned __int128 acc = 0;
      for (; t < e; t++) { unsigned d = (unsigned)(*t - '0'); if (d > 9) bad(); acc = (acc * 10 + d) % P; }
      v = (uint64_t)acc;
    }
    if (negative && v) v = P - v;
    return v;
  }
  // x mod p through 2^64 = 2^32 - 1 and 2^96 = -1 (mod p) instead of a 128-bit division
  static uint64_t reduce128(unsigned __int128 x) {
    const uint64_t P = 0xFFFFFFFF00000001ULL, EPS = 0xFFFFFFFFULL;
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64), hh = hi >> 32, hl = hi & EPS;
    uint64_t t = lo - hh;
    if (lo < hh) t -= EPS;  // borrowed 2^64: add p instead
    uint64_t m = hl * EPS, r = t + m;
    if (r < t) r += EPS;    // carried 2^64 = 2^32 - 1
    return r >= P ? r - P : r;
  }
  long long integer() const {
    const JNode &nd = node();
    if (nd.kind != JNode::Number) throw JsonError("expected a number");
    const char *t = nd.p, *e = nd.p + nd.len;
    bool negative = false;
    if (t < e && *t == '-') { negative = true; t++; }
    long long acc = 0;
    if (t >= e) throw JsonError("bad number");
    for (; t < e; t++) {
      char c = *t;
      if (c < '0' || c > '9') throw JsonError("non-integer number \"" + token() + "\"");
      if (acc > (long long)4e17) throw JsonError("integer out of range");
      acc = acc * 10 + (c - '0');
    }
    return negative ? -acc : acc;
  }
  bool boolean() const {
    uint8_t k = node().kind;
    if (k != JNode::True && k != JNode::False) throw JsonError("expected a boolean");
    return k == JNode::True;
  }
  std::string str() const {
    if (node().kind != JNode::String) throw JsonError("expected a string");
    return decode(node());
  }

 private:
  friend class JList;
  const std::vector<JNode> *tape_ = nullptr;
  uint32_t idx_ = 0;
  const JNode &node() const { return (*tape_)[idx_]; }
  [[noreturn]] void bad() const { throw JsonError("non-integer number \"" + token() + "\" where a field element is expected"); }
  uint32_t memberIndex(size_t i) const {
    if (i >= objSize()) throw JsonError("object member out of range");
    uint32_t k = idx_ + 1;
    for (size_t m = 0; m < i; m++) k = (*tape_)[k + 1].next;
    return k;
  }
  static std::string decode(const JNode &n) {
    if (!n.escaped) return std::string(n.p, n.len);
    std::string out;
    const char *p = n.p, *end = n.p + n.len;
    while (p < end) {
      char c = *p++;
      if (c != '\\') { out += c; continue; }
      if (p >= end) throw JsonError("bad escape");
      char e = *p++;
      switch (e) {
        case '"': out += '"'; break;
        case '\\': out += '\\'; break;
        case '/': out += '/'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'n': out += '\n'; break;
        case 'r': out += '\r'; break;
        case 't': out += '\t'; break;
        case 'u': {
          if (end - p < 4) throw JsonError("bad \\u escape");
          unsigned cp = 0;
          for (int i = 0; i < 4; i++) {
            char h = *p++;
            cp <<= 4;
            if (h >= '0' && h <= '9') cp |= h - '0';
            else if (h >= 'a' && h <= 'f') cp |= h - 'a' + 10;
            else if (h >= 'A' && h <= 'F') cp |= h - 'A' + 10;
            else throw JsonError("bad \\u escape");
          }
          if (cp < 0x80) out += (char)cp;
          else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
          else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
          break;
        }
        default: throw JsonError("bad escape");
      }
    }
    return out;
  }
};

// The elements of an array: size() is O(1), iteration follows the subtree links, [i] walks i links.
class JList {
 public:
  class iterator {
   public:
    iterator(const std::vector<JNode> *t, uint32_t i) : cur_(t, i) {}
    const JValue &operator*() const { return cur_; }
    iterator &operator++() { cur_ = JValue(cur_.tape_, (*cur_.tape_)[cur_.idx_].next); return *this; }
    bool operator!=(const iterator &o) const { return cur_.idx_ != o.cur_.idx_; }
   private:
    JValue cur_;
  };
  JList(const std::vector<JNode> *t, uint32_t arr) : tape_(t), arr_(arr) {}
  size_t size() const { return (*tape_)[arr_].len; }
  iterator begin() const { return iterator(tape_, arr_ + 1); }
  iterator end() const { return iterator(tape_, (*tape_)[arr_].next); }
  JValue operator[](size_t i) const {
    if (i >= size()) throw JsonError("array index out of range");
    uint32_t k = arr_ + 1;
    for (size_t m = 0; m < i; m++) k = (*tape_)[k].next;
    return JValue(tape_, k);
  }
 private:
  const std::vector<JNode> *tape_;
  uint32_t arr_;
};

inline JList JValue::list() const {
  if (node().kind != JNode::Array) throw JsonError("expected an array");
  return JList(tape_, idx_);
}

// Owns the tape; the text must outlive every JValue taken from it.
class JsonDoc {
 public:
  JsonDoc(const char *p, size_t n) : p_(p), end_(p + n) {
    if (n > 0x7fffffffu) throw JsonError("document larger than 2 GiB");  // node indices and token lengths are 32-bit
    tape_.reserve(n / 12 + 16);
    value(0);
    ws();
    if (p_ != end_) throw JsonError("trailing characters after JSON value");
  }
  JValue root() const { return JValue(&tape_, 0); }

 private:
  const char *p_, *end_;
  std::vector<JNode> tape_;
  static constexpr int kMaxDepth = 256;

  void ws() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) p_++;
  }
  bool lit(const char *s, size_t n) {
    if ((size_t)(end_ - p_) >= n && memcmp(p_, s, n) == 0) { p_ += n; return true; }
    return false;
  }
  uint32_t push(uint8_t kind) {
    tape_.emplace_back();
    tape_.back().kind = kind;
    return (uint32_t)tape_.size() - 1;
  }
  void stringToken(uint32_t at) {  // p_ at the opening quote
    p_++;
    const char *s = p_;
    uint8_t esc = 0;
    while (p_ < end_ && *p_ != '"') {
      if (*p_ == '\\') { esc = 1; p_++; if (p_ >= end_) break; }
      p_++;
    }
    if (p_ >= end_) throw JsonError("unterminated string");
    tape_[at].p = s;
    tape_[at].len = (uint32_t)(p_ - s);
    tape_[at].escaped = esc;
    p_++;  // closing quote
  }
  void value(int depth) {
    if (depth > kMaxDepth) throw JsonError("nesting too deep");
    ws();
    if (p_ >= end_) throw JsonError("unexpected end of input");
    char c = *p_;
    if (c == '{') {
      p_++;
      uint32_t me = push(JNode::Object);
      uint32_t count = 0;
      ws();
      if (p_ < end_ && *p_ == '}') { p_++; }
      else
        for (;;) {
          ws();
          if (p_ >= end_ || *p_ != '"') throw JsonError("expected object key");
          uint32_t k = push(JNode::Key);
          stringToken(k);
          tape_[k].next = k + 1;
          ws();
          if (p_ >= end_ || *p_ != ':') throw JsonError("expected ':'");
          p_++;
          value(depth + 1);
          count++;
          ws();
          if (p_ < end_ && *p_ == ',') { p_++; continue; }
          if (p_ < end_ && *p_ == '}') { p_++; break; }
          throw JsonError("expected ',' or '}'");
        }
      tape_[me].len = count;
      tape_[me].next = (uint32_t)tape_.size();
    } else if (c == '[') {
      p_++;
      uint32_t me = push(JNode::Array);
      uint32_t count = 0;
      ws();
      if (p_ < end_ && *p_ == ']') { p_++; }
      else
        for (;;) {
          value(depth + 1);
          count++;
          ws();
          if (p_ < end_ && *p_ == ',') { p_++; continue; }
          if (p_ < end_ && *p_ == ']') { p_++; break; }
          throw JsonError("expected ',' or ']'");
        }
      tape_[me].len = count;
      tape_[me].next = (uint32_t)tape_.size();
    } else if (c == '"') {
      uint32_t me = push(JNode::String);
      stringToken(me);
      tape_[me].next = me + 1;
    } else if (c == 't' && lit("true", 4)) {
      uint32_t me = push(JNode::True); tape_[me].next = me + 1;
    } else if (c == 'f' && lit("false", 5)) {
      uint32_t me = push(JNode::False); tape_[me].next = me + 1;
    } else if (c == 'n' && lit("null", 4)) {
      uint32_t me = push(JNode::Null); tape_[me].next = me + 1;
    } else if (c == '-' || (c >= '0' && c <= '9')) {
      const char *s = p_;
      if (*p_ == '-') p_++;
      // JSON grammar (and aeson): no leading zeros — "0" is a number, "007" is not
      if (end_ - p_ >= 2 && p_[0] == '0' && (unsigned)(p_[1] - '0') < 10) throw JsonError("number with a leading zero");
      while (end_ - p_ >= 8 && eightDigits(p_)) p_ += 8;
      while (p_ < end_ && (unsigned)(*p_ - '0') < 10) p_++;
      while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) p_++;
      uint32_t me = push(JNode::Number);
      tape_[me].p = s;
      tape_[me].len = (uint32_t)(p_ - s);
      tape_[me].next = me + 1;
    } else {
      throw JsonError(std::string("unexpected character '") + c + "'");
    }
  }
};

}  // namespace p2vhost
