// mini_json.hpp — a small recursive-descent JSON reader, just enough for the reference's
// scene files (objects, arrays, numbers, strings, booleans, null).  Numbers are kept as
// double (strtod, correctly rounded) plus an "is integer" flag; callers narrow with
// static_cast<float>, which is what the reference's vendored JSON library does for get<float>().
#pragma once

#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace minijson {

struct Value {
  enum class Kind { Null, Bool, Number, String, Array, Object } kind = Kind::Null;
  bool boolean = false;
  double number = 0.0;
  bool is_integer = false;
  long long integer = 0;
  std::string string;
  std::vector<Value> array;
  std::vector<std::pair<std::string, Value>> object;  // insertion order kept

  bool is_object() const { return kind == Kind::Object; }
  bool is_array() const { return kind == Kind::Array; }
  bool is_number() const { return kind == Kind::Number; }
  bool is_string() const { return kind == Kind::String; }
  const Value* find(const std::string& key) const {
    if (kind != Kind::Object) return nullptr;
    for (const auto& kv : object)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool has(const std::string& key) const { return find(key) != nullptr; }
  const Value& at(const std::string& key) const {
    const Value* v = find(key);
    if (!v) throw std::runtime_error("key '" + key + "' not found");
    return *v;
  }
  float as_float() const {
    if (kind != Kind::Number) throw std::runtime_error("type must be number");
    return is_integer ? static_cast<float>(integer) : static_cast<float>(number);
  }
  int as_int() const {
    if (kind != Kind::Number) throw std::runtime_error("type must be number");
    return is_integer ? static_cast<int>(integer) : static_cast<int>(number);
  }
  const std::string& as_string() const {
    if (kind != Kind::String) throw std::runtime_error("type must be string");
    return string;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& text) : s_(text) {}
  Value parse() {
    Value v = value();
    skip_ws();
    if (pos_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  std::size_t pos_ = 0;

  [[noreturn]] void fail(const std::string& why) const {
    throw std::runtime_error("parse error at offset " + std::to_string(pos_) + ": " + why);
  }
  void skip_ws() {
    while (pos_ < s_.size() && (s_[pos_] == ' ' || s_[pos_] == '\t' || s_[pos_] == '\n' || s_[pos_] == '\r')) ++pos_;
  }
  bool consume(char c) {
    skip_ws();
    if (pos_ < s_.size() && s_[pos_] == c) {
      ++pos_;
      return true;
    }
    return false;
  }
  void expect(char c) {
    if (!consume(c)) fail(std::string("expected '") + c + "'");
  }
  Value value() {
    skip_ws();
    if (pos_ >= s_.size()) fail("unexpected end of input");
    const char c = s_[pos_];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') {
      Value v;
      v.kind = Value::Kind::String;
      v.string = string();
      return v;
    }
    if (c == 't' || c == 'f' || c == 'n') return literal();
    return number();
  }
  Value object() {
    Value v;
    v.kind = Value::Kind::Object;
    expect('{');
    if (consume('}')) return v;
    do {
      skip_ws();
      if (pos_ >= s_.size() || s_[pos_] != '"') fail("expected string key");
      std::string key = string();
      expect(':');
      v.object.emplace_back(std::move(key), value());
    } while (consume(','));
    expect('}');
    return v;
  }
  Value array() {
    Value v;
    v.kind = Value::Kind::Array;
    expect('[');
    if (consume(']')) return v;
    do {
      v.array.push_back(value());
    } while (consume(','));
    expect(']');
    return v;
  }
  std::string string() {
    std::string out;
    ++pos_;  // opening quote
    while (pos_ < s_.size() && s_[pos_] != '"') {
      char c = s_[pos_++];
      if (c == '\\') {
        if (pos_ >= s_.size()) fail("bad escape");
        const char e = s_[pos_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (pos_ + 4 > s_.size()) fail("bad \\u escape");
            const unsigned cp = static_cast<unsigned>(std::strtoul(s_.substr(pos_, 4).c_str(), nullptr, 16));
            pos_ += 4;
            if (cp < 0x80) out += static_cast<char>(cp);
            else if (cp < 0x800) { out += static_cast<char>(0xC0 | (cp >> 6)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
            else { out += static_cast<char>(0xE0 | (cp >> 12)); out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e;
        }
      } else {
        out += c;
      }
    }
    if (pos_ >= s_.size()) fail("unterminated string");
    ++pos_;
    return out;
  }
  Value literal() {
    Value v;
    if (s_.compare(pos_, 4, "true") == 0) { v.kind = Value::Kind::Bool; v.boolean = true; pos_ += 4; }
    else if (s_.compare(pos_, 5, "false") == 0) { v.kind = Value::Kind::Bool; pos_ += 5; }
    else if (s_.compare(pos_, 4, "null") == 0) { pos_ += 4; }
    else fail("unknown literal");
    return v;
  }
  Value number() {
    const std::size_t start = pos_;
    bool integral = true;
    if (pos_ < s_.size() && (s_[pos_] == '-' || s_[pos_] == '+')) ++pos_;
    while (pos_ < s_.size()) {
      const char c = s_[pos_];
      if (c >= '0' && c <= '9') { ++pos_; continue; }
      if (c == '.' || c == 'e' || c == 'E' || c == '-' || c == '+') { integral = false; ++pos_; continue; }
      break;
    }
    if (pos_ == start) fail("expected a value");
    const std::string tok = s_.substr(start, pos_ - start);
    Value v;
    v.kind = Value::Kind::Number;
    char* end = nullptr;
    v.number = std::strtod(tok.c_str(), &end);
    if (end == tok.c_str()) fail("bad number");
    if (integral && tok.size() < 18) {
      v.is_integer = true;
      v.integer = std::strtoll(tok.c_str(), nullptr, 10);
    }
    return v;
  }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace minijson
