// Minimal JSON DOM for the reference's wire format (src/Types.hs aeson instances).
// Numbers are kept as exact decimal integers reduced mod 2^64 / mod p on demand: field elements
// must never go through `double`.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace p2vhost {

struct JsonError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct JValue {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  std::string text;  // Number: the token; String: the decoded string
  std::vector<JValue> arr;
  std::vector<std::pair<std::string, JValue>> obj;

  const JValue *find(const std::string &key) const {
    if (kind != Object) return nullptr;
    for (auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const JValue &at(const std::string &key) const {
    const JValue *v = find(key);
    if (!v) throw JsonError("missing key \"" + key + "\"");
    return *v;
  }
  const std::vector<JValue> &list() const {
    if (kind != Array) throw JsonError("expected an array");
    return arr;
  }
  // non-negative integer token -> value mod p (p = 2^64 - 2^32 + 1), like `mkGoldilocks <$> parseJSON`
  uint64_t felt() const {
    if (kind != Number) throw JsonError("expected a number");
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    size_t i = 0;
    bool negative = false;
    if (i < text.size() && text[i] == '-') { negative = true; i++; }
    if (i >= text.size()) throw JsonError("bad number");
    unsigned __int128 acc = 0;
    for (; i < text.size(); i++) {
      char c = text[i];
      if (c < '0' || c > '9') throw JsonError("non-integer number \"" + text + "\" where a field element is expected");
      acc = (acc * 10 + (unsigned)(c - '0')) % P;
    }
    uint64_t v = (uint64_t)acc;
    if (negative && v) v = P - v;
    return v;
  }
  long long integer() const {
    if (kind != Number) throw JsonError("expected a number");
    size_t i = 0;
    bool negative = false;
    if (i < text.size() && text[i] == '-') { negative = true; i++; }
    long long acc = 0;
    if (i >= text.size()) throw JsonError("bad number");
    for (; i < text.size(); i++) {
      char c = text[i];
      if (c < '0' || c > '9') throw JsonError("non-integer number \"" + text + "\"");
      if (acc > (long long)4e17) throw JsonError("integer out of range");
      acc = acc * 10 + (c - '0');
    }
    return negative ? -acc : acc;
  }
  bool boolean() const {
    if (kind != Bool) throw JsonError("expected a boolean");
    return b;
  }
  const std::string &str() const {
    if (kind != String) throw JsonError("expected a string");
    return text;
  }
};

class JsonParser {
 public:
  JsonParser(const char *p, size_t n) : p_(p), end_(p + n) {}
  JValue parse() {
    JValue v = value();
    ws();
    if (p_ != end_) throw JsonError("trailing characters after JSON value");
    return v;
  }

 private:
  const char *p_, *end_;
  void ws() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) p_++;
  }
  bool lit(const char *s) {
    size_t n = strlen_(s);
    if ((size_t)(end_ - p_) >= n && std::string(p_, n) == s) { p_ += n; return true; }
    return false;
  }
  static size_t strlen_(const char *s) { size_t n = 0; while (s[n]) n++; return n; }
  JValue value() {
    ws();
    if (p_ >= end_) throw JsonError("unexpected end of input");
    JValue v;
    char c = *p_;
    if (c == '{') {
      p_++;
      v.kind = JValue::Object;
      ws();
      if (p_ < end_ && *p_ == '}') { p_++; return v; }
      for (;;) {
        ws();
        if (p_ >= end_ || *p_ != '"') throw JsonError("expected object key");
        std::string k = string_();
        ws();
        if (p_ >= end_ || *p_ != ':') throw JsonError("expected ':'");
        p_++;
        v.obj.emplace_back(std::move(k), value());
        ws();
        if (p_ < end_ && *p_ == ',') { p_++; continue; }
        if (p_ < end_ && *p_ == '}') { p_++; break; }
        throw JsonError("expected ',' or '}'");
      }
    } else if (c == '[') {
      p_++;
      v.kind = JValue::Array;
      ws();
      if (p_ < end_ && *p_ == ']') { p_++; return v; }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (p_ < end_ && *p_ == ',') { p_++; continue; }
        if (p_ < end_ && *p_ == ']') { p_++; break; }
        throw JsonError("expected ',' or ']'");
      }
    } else if (c == '"') {
      v.kind = JValue::String;
      v.text = string_();
    } else if (c == 't' && lit("true")) {
      v.kind = JValue::Bool; v.b = true;
    } else if (c == 'f' && lit("false")) {
      v.kind = JValue::Bool; v.b = false;
    } else if (c == 'n' && lit("null")) {
      v.kind = JValue::Null;
    } else if (c == '-' || (c >= '0' && c <= '9')) {
      const char *s = p_;
      if (*p_ == '-') p_++;
      while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) p_++;
      v.kind = JValue::Number;
      v.text.assign(s, p_ - s);
    } else {
      throw JsonError(std::string("unexpected character '") + c + "'");
    }
    return v;
  }
  std::string string_() {
    std::string out;
    p_++;  // opening quote
    while (p_ < end_ && *p_ != '"') {
      char c = *p_++;
      if (c == '\\') {
        if (p_ >= end_) throw JsonError("bad escape");
        char e = *p_++;
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
            if (end_ - p_ < 4) throw JsonError("bad \\u escape");
            unsigned cp = 0;
            for (int i = 0; i < 4; i++) {
              char h = *p_++;
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
      } else {
        out += c;
      }
    }
    if (p_ >= end_) throw JsonError("unterminated string");
    p_++;  // closing quote
    return out;
  }
};

}  // namespace p2vhost
