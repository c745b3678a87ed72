// Tiny JSON reader/writer: /etc/nvidia/gpu_config.json and the Kubernetes API objects the health checker edits
// (objects, arrays, strings with \uXXXX escapes, numbers kept as their source text so a GET -> edit -> PUT round trip
// cannot lose int64 precision, true/false/null). No dependencies.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace json {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Value> arr;
  std::map<std::string, Value> obj;

  const Value* get(const std::string& key) const {
    if (kind != Object) return nullptr;
    auto it = obj.find(key);
    return it == obj.end() ? nullptr : &it->second;
  }
  std::string get_string(const std::string& key, const std::string& dflt = "") const {
    const Value* v = get(key);
    return (v && v->kind == String) ? v->str : dflt;
  }
  long get_int(const std::string& key, long dflt = 0) const {
    const Value* v = get(key);
    return (v && v->kind == Number) ? (long)v->num : dflt;
  }
  Value* find(const std::string& key) {
    if (kind != Object) return nullptr;
    auto it = obj.find(key);
    return it == obj.end() ? nullptr : &it->second;
  }
  Value& at(const std::string& key) {            // object member, created (and this value turned into an object) on demand
    if (kind != Object) { *this = Value(); kind = Object; }
    return obj[key];
  }
  static Value of(const std::string& s) { Value v; v.kind = String; v.str = s; return v; }
  static Value of(const char* s) { return of(std::string(s)); }
  static Value of(bool b) { Value v; v.kind = Bool; v.b = b; return v; }
  static Value of(long n) { Value v; v.kind = Number; v.num = (double)n; v.str = std::to_string(n); return v; }
  static Value object() { Value v; v.kind = Object; return v; }
  static Value array() { Value v; v.kind = Array; return v; }
};

inline void escape_to(const std::string& s, std::string* out) {
  out->push_back('"');
  for (unsigned char c : s) {
    switch (c) {
      case '"': *out += "\\\""; break;
      case '\\': *out += "\\\\"; break;
      case '\n': *out += "\\n"; break;
      case '\r': *out += "\\r"; break;
      case '\t': *out += "\\t"; break;
      case '\b': *out += "\\b"; break;
      case '\f': *out += "\\f"; break;
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); *out += b; }
        else out->push_back((char)c);                     // UTF-8 passes through untouched
    }
  }
  out->push_back('"');
}

inline void dump_to(const Value& v, std::string* out) {
  switch (v.kind) {
    case Value::Null: *out += "null"; break;
    case Value::Bool: *out += v.b ? "true" : "false"; break;
    case Value::Number:
      if (!v.str.empty()) *out += v.str;
      else { char b[40]; snprintf(b, sizeof b, "%.17g", v.num); *out += b; }
      break;
    case Value::String: escape_to(v.str, out); break;
    case Value::Array: {
      out->push_back('[');
      for (size_t i = 0; i < v.arr.size(); i++) { if (i) out->push_back(','); dump_to(v.arr[i], out); }
      out->push_back(']');
      break;
    }
    case Value::Object: {
      out->push_back('{');
      bool first = true;
      for (auto& kv : v.obj) { if (!first) out->push_back(','); first = false; escape_to(kv.first, out); out->push_back(':'); dump_to(kv.second, out); }
      out->push_back('}');
      break;
    }
  }
}

inline std::string dump(const Value& v) { std::string s; dump_to(v, &s); return s; }

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s) {}
  bool parse(Value* out, std::string* err) {
    *out = Value();          // a reused Value must not keep members of the previous document
    ws();
    if (!value(out)) { *err = err_.empty() ? "invalid JSON" : err_; return false; }
    ws();
    if (i_ != s_.size()) { *err = "trailing characters after JSON value"; return false; }
    return true;
  }

 private:
  void ws() { while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t' || s_[i_] == '\n' || s_[i_] == '\r')) i_++; }
  bool lit(const char* w) { size_t n = strlen(w); if (s_.compare(i_, n, w) == 0) { i_ += n; return true; } return false; }
  struct Depth { int& d; explicit Depth(int& x) : d(x) { d++; } ~Depth() { d--; } };
  bool value(Value* v) {
    if (i_ >= s_.size()) return false;
    const char c = s_[i_];
    if (c == '{' || c == '[') {
      Depth guard(depth_);
      if (depth_ > kMaxDepth) { err_ = "JSON nested too deeply"; return false; }
      return c == '{' ? object(v) : array(v);
    }
    if (c == '"') { v->kind = Value::String; return string(&v->str); }
    if (lit("true")) { v->kind = Value::Bool; v->b = true; return true; }
    if (lit("false")) { v->kind = Value::Bool; v->b = false; return true; }
    if (lit("null")) { v->kind = Value::Null; return true; }
    char* end = nullptr;
    v->num = strtod(s_.c_str() + i_, &end);
    if (end == s_.c_str() + i_) return false;
    v->str.assign(s_.c_str() + i_, (size_t)(end - (s_.c_str() + i_)));              // source text, re-emitted verbatim by dump()
    i_ = (size_t)(end - s_.c_str());
    v->kind = Value::Number;
    return true;
  }
  static void utf8(unsigned cp, std::string* out) {
    if (cp < 0x80) out->push_back((char)cp);
    else if (cp < 0x800) { out->push_back((char)(0xC0 | (cp >> 6))); out->push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) { out->push_back((char)(0xE0 | (cp >> 12))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
    else { out->push_back((char)(0xF0 | (cp >> 18))); out->push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
  }
  bool hex4(unsigned* cp) {
    if (i_ + 4 > s_.size()) return false;
    unsigned v = 0;
    for (int k = 0; k < 4; k++) {
      const char c = s_[i_ + k];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
      else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
      else return false;
    }
    i_ += 4; *cp = v;
    return true;
  }
  bool string(std::string* out) {
    i_++;   // opening quote
    while (i_ < s_.size() && s_[i_] != '"') {
      if (s_[i_] == '\\' && i_ + 1 < s_.size()) {
        const char e = s_[i_ + 1];
        i_ += 2;
        if (e == 'u') {
          unsigned cp = 0;
          if (!hex4(&cp)) return false;
          if (cp >= 0xD800 && cp < 0xDC00 && i_ + 1 < s_.size() && s_[i_] == '\\' && s_[i_ + 1] == 'u') {   // surrogate pair
            i_ += 2;
            unsigned lo = 0;
            if (!hex4(&lo)) return false;
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8(cp, out);
        } else {
          switch (e) {
            case 'n': out->push_back('\n'); break;
            case 't': out->push_back('\t'); break;
            case 'r': out->push_back('\r'); break;
            case 'b': out->push_back('\b'); break;
            case 'f': out->push_back('\f'); break;
            case '"': case '\\': case '/': out->push_back(e); break;
            default: err_ = "invalid escape in JSON string"; return false;
          }
        }
      } else out->push_back(s_[i_++]);
    }
    if (i_ >= s_.size()) return false;
    i_++;
    return true;
  }
  bool array(Value* v) {
    v->kind = Value::Array; i_++; ws();
    if (i_ < s_.size() && s_[i_] == ']') { i_++; return true; }
    while (true) {
      Value e; ws(); if (!value(&e)) return false;
      v->arr.push_back(std::move(e)); ws();
      if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
      if (i_ < s_.size() && s_[i_] == ']') { i_++; return true; }
      return false;
    }
  }
  bool object(Value* v) {
    v->kind = Value::Object; i_++; ws();
    if (i_ < s_.size() && s_[i_] == '}') { i_++; return true; }
    while (true) {
      ws(); if (i_ >= s_.size() || s_[i_] != '"') return false;
      std::string k; if (!string(&k)) return false;
      ws(); if (i_ >= s_.size() || s_[i_] != ':') return false;
      i_++; ws();
      Value e; if (!value(&e)) return false;
      v->obj[k] = std::move(e); ws();
      if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
      if (i_ < s_.size() && s_[i_] == '}') { i_++; return true; }
      return false;
    }
  }
  static constexpr int kMaxDepth = 128;   // Kubernetes objects nest < 20 levels; bounds the recursion on hostile input
  const std::string& s_;
  size_t i_ = 0;
  int depth_ = 0;
  std::string err_;
};

}  // namespace json
