// Tiny JSON reader for /etc/nvidia/gpu_config.json (objects, arrays, strings, numbers, true/false/null). No dependencies.
#pragma once
#include <stdlib.h>

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
};

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s) {}
  bool parse(Value* out, std::string* err) {
    ws();
    if (!value(out)) { *err = err_.empty() ? "invalid JSON" : err_; return false; }
    ws();
    if (i_ != s_.size()) { *err = "trailing characters after JSON value"; return false; }
    return true;
  }

 private:
  void ws() { while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t' || s_[i_] == '\n' || s_[i_] == '\r')) i_++; }
  bool lit(const char* w) { size_t n = strlen(w); if (s_.compare(i_, n, w) == 0) { i_ += n; return true; } return false; }
  bool value(Value* v) {
    if (i_ >= s_.size()) return false;
    const char c = s_[i_];
    if (c == '{') return object(v);
    if (c == '[') return array(v);
    if (c == '"') { v->kind = Value::String; return string(&v->str); }
    if (lit("true")) { v->kind = Value::Bool; v->b = true; return true; }
    if (lit("false")) { v->kind = Value::Bool; v->b = false; return true; }
    if (lit("null")) { v->kind = Value::Null; return true; }
    char* end = nullptr;
    v->num = strtod(s_.c_str() + i_, &end);
    if (end == s_.c_str() + i_) return false;
    i_ = (size_t)(end - s_.c_str());
    v->kind = Value::Number;
    return true;
  }
  bool string(std::string* out) {
    i_++;   // opening quote
    while (i_ < s_.size() && s_[i_] != '"') {
      if (s_[i_] == '\\' && i_ + 1 < s_.size()) {
        const char e = s_[i_ + 1];
        out->push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e == 'r' ? '\r' : e);
        i_ += 2;
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
  const std::string& s_;
  size_t i_ = 0;
  std::string err_;
};

}  // namespace json
