// Minimal JSON reader/writer for the JSON chunk of .scene files. Mirrors the behaviour of the reference's hand-written
// JSON class that matters for this path (reference src/JSON.cpp:167-192): a number containing '.', 'e' or 'E' is a float
// (parsed as fp32), anything else an int32; the writer prints floats with std::to_string (6 decimals, src/JSON.hpp:27-32).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace vkx {

class Json {
  public:
    enum class Type { Null, Bool, Int, Real, String, Array, Object };
    Json() = default;
    Json(bool b) : _type(Type::Bool), _b(b) {}
    Json(int i) : _type(Type::Int), _i(i) {}
    Json(float f) : _type(Type::Real), _f(f) {}
    Json(const std::string& s) : _type(Type::String), _s(s) {}
    Json(const char* s) : _type(Type::String), _s(s) {}
    static Json array() { Json j; j._type = Type::Array; return j; }
    static Json object() { Json j; j._type = Type::Object; return j; }

    static Json parse(const char* data, size_t length);
    std::string toString() const;

    Type type() const { return _type; }
    bool contains(const std::string& key) const { return _type == Type::Object && _keys.count(key); }
    const Json& operator[](const std::string& key) const;
    Json& operator[](const std::string& key);
    const Json& operator[](size_t i) const { return _items.at(i); }
    size_t size() const { return _items.size(); }
    const std::vector<Json>& items() const { return _items; }
    void push(const Json& v) { _type = Type::Array; _items.push_back(v); }

    int asInt() const { if (_type == Type::Int) return _i; if (_type == Type::Real) return int(_f); throw std::runtime_error("JSON: not a number"); }
    float asFloat() const { if (_type == Type::Real) return _f; if (_type == Type::Int) return float(_i); throw std::runtime_error("JSON: not a number"); }
    const std::string& asString() const { if (_type != Type::String) throw std::runtime_error("JSON: not a string"); return _s; }
    int get(const std::string& key, int def) const { return contains(key) ? (*this)[key].asInt() : def; }
    float get(const std::string& key, float def) const { return contains(key) ? (*this)[key].asFloat() : def; }

  private:
    Type _type = Type::Null;
    bool _b = false; int _i = 0; float _f = 0.f; std::string _s;
    std::vector<Json> _items;                 // array items, or object values in insertion order
    std::vector<std::string> _names;          // object keys in insertion order
    std::map<std::string, size_t> _keys;      // key -> index into _items
    friend struct JsonParser;
};

} // namespace vkx
