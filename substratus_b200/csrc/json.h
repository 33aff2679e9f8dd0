// json.h — minimal JSON DOM (config.json, params.json, safetensors headers, HTTP bodies).  Header-only, no deps.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ssb {

struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept (safetensors tensors)

  bool is_null() const { return kind == Null; }
  const Json* find(const std::string& k) const {
    if (kind != Obj) return nullptr;
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  bool has(const std::string& k) const {
    const Json* j = find(k);
    return j && !j->is_null();
  }
  double get_num(const std::string& k, double dflt) const {
    const Json* j = find(k);
    if (!j) return dflt;
    if (j->kind == Num) return j->num;
    if (j->kind == Bool) return j->b ? 1 : 0;
    if (j->kind == Str) {  // params.json values arrive as strings when set through `env`-style maps
      char* end = nullptr;
      double v = strtod(j->str.c_str(), &end);
      if (end && *end == 0 && !j->str.empty()) return v;
    }
    return dflt;
  }
  int64_t get_int(const std::string& k, int64_t dflt) const { return (int64_t)get_num(k, (double)dflt); }
  std::string get_str(const std::string& k, const std::string& dflt) const {
    const Json* j = find(k);
    return (j && j->kind == Str) ? j->str : dflt;
  }
};

class JsonParser {
 public:
  JsonParser(const char* s, size_t n) : p_(s), e_(s + n) {}
  Json parse() {
    Json j = value();
    ws();
    if (p_ != e_) fail("trailing characters");
    return j;
  }

 private:
  const char *p_, *e_;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m); }
  void ws() {
    while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
  }
  bool lit(const char* s) {
    size_t n = strlen_(s);
    if ((size_t)(e_ - p_) >= n && std::equal(s, s + n, p_)) {
      p_ += n;
      return true;
    }
    return false;
  }
  static size_t strlen_(const char* s) {
    size_t n = 0;
    while (s[n]) ++n;
    return n;
  }
  // nesting guard: params.json comes from the user's CRD (.spec.params) and request bodies from the network; a few
  // hundred thousand '[' must end in an error, not in a stack overflow of this recursive descent
  int depth_ = 0;
  struct Depth {
    int& d;
    explicit Depth(int& x) : d(x) { ++d; }
    ~Depth() { --d; }
  };
  Json value() {
    Depth guard(depth_);
    if (depth_ > 128) fail("nesting too deep");
    ws();
    if (p_ >= e_) fail("unexpected end");
    Json j;
    char c = *p_;
    if (c == '{') {
      ++p_;
      j.kind = Json::Obj;
      ws();
      if (p_ < e_ && *p_ == '}') {
        ++p_;
        return j;
      }
      for (;;) {
        ws();
        if (p_ >= e_ || *p_ != '"') fail("expected key");
        std::string k = string();
        ws();
        if (p_ >= e_ || *p_ != ':') fail("expected ':'");
        ++p_;
        j.obj.emplace_back(std::move(k), value());
        ws();
        if (p_ < e_ && *p_ == ',') {
          ++p_;
          continue;
        }
        if (p_ < e_ && *p_ == '}') {
          ++p_;
          return j;
        }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      ++p_;
      j.kind = Json::Arr;
      ws();
      if (p_ < e_ && *p_ == ']') {
        ++p_;
        return j;
      }
      for (;;) {
        j.arr.push_back(value());
        ws();
        if (p_ < e_ && *p_ == ',') {
          ++p_;
          continue;
        }
        if (p_ < e_ && *p_ == ']') {
          ++p_;
          return j;
        }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') {
      j.kind = Json::Str;
      j.str = string();
      return j;
    }
    if (lit("true")) {
      j.kind = Json::Bool;
      j.b = true;
      return j;
    }
    if (lit("false")) {
      j.kind = Json::Bool;
      return j;
    }
    if (lit("null")) return j;
    if (lit("NaN") || lit("Infinity") || lit("-Infinity")) {  // python json.dump may emit these
      j.kind = Json::Num;
      return j;
    }
    char* end = nullptr;
    std::string tmp(p_, (size_t)std::min<ptrdiff_t>(e_ - p_, 64));
    double v = strtod(tmp.c_str(), &end);
    if (end == tmp.c_str()) fail("bad value");
    p_ += end - tmp.c_str();
    j.kind = Json::Num;
    j.num = v;
    return j;
  }
  static void utf8(std::string& o, unsigned cp) {
    if (cp < 0x80)
      o += (char)cp;
    else if (cp < 0x800) {
      o += (char)(0xC0 | (cp >> 6));
      o += (char)(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
      o += (char)(0xE0 | (cp >> 12));
      o += (char)(0x80 | ((cp >> 6) & 0x3F));
      o += (char)(0x80 | (cp & 0x3F));
    } else {
      o += (char)(0xF0 | (cp >> 18));
      o += (char)(0x80 | ((cp >> 12) & 0x3F));
      o += (char)(0x80 | ((cp >> 6) & 0x3F));
      o += (char)(0x80 | (cp & 0x3F));
    }
  }
  unsigned hex4() {
    if (e_ - p_ < 4) fail("bad \\u");
    unsigned v = 0;
    for (int i = 0; i < 4; ++i) {
      char c = *p_++;
      v <<= 4;
      if (c >= '0' && c <= '9')
        v |= c - '0';
      else if (c >= 'a' && c <= 'f')
        v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F')
        v |= c - 'A' + 10;
      else
        fail("bad hex");
    }
    return v;
  }
  std::string string() {
    ++p_;  // opening quote
    std::string o;
    while (p_ < e_) {
      char c = *p_++;
      if (c == '"') return o;
      if (c != '\\') {
        o += c;
        continue;
      }
      if (p_ >= e_) break;
      char x = *p_++;
      switch (x) {
        case 'n': o += '\n'; break;
        case 't': o += '\t'; break;
        case 'r': o += '\r'; break;
        case 'b': o += '\b'; break;
        case 'f': o += '\f'; break;
        case 'u': {
          unsigned cp = hex4();
          if (cp >= 0xD800 && cp < 0xDC00 && e_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
            p_ += 2;
            unsigned lo = hex4();
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8(o, cp);
          break;
        }
        default: o += x;
      }
    }
    fail("unterminated string");
  }
};

inline Json json_parse(const std::string& s) { return JsonParser(s.data(), s.size()).parse(); }

inline std::string json_escape(const std::string& s) {
  std::string o;
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) {
          char buf[8];
          snprintf(buf, sizeof buf, "\\u%04x", c);
          o += buf;
        } else
          o += (char)c;
    }
  }
  return o;
}

}  // namespace ssb
