#include "Json.hpp"
#include <charconv>
#include <cstdlib>
#include <cstring>

namespace vkx {

const Json& Json::operator[](const std::string& key) const {
    auto it = _keys.find(key);
    if (it == _keys.end()) throw std::runtime_error("JSON: missing key '" + key + "'");
    return _items[it->second];
}
Json& Json::operator[](const std::string& key) {
    if (_type != Type::Object) { _type = Type::Object; _items.clear(); _names.clear(); _keys.clear(); }
    auto it = _keys.find(key);
    if (it != _keys.end()) return _items[it->second];
    _keys[key] = _items.size(); _names.push_back(key); _items.emplace_back();
    return _items.back();
}

struct JsonParser {
    const char* p; const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("JSON parse error: ") + what); }
    Json value() {
        ws();
        if (p >= end) fail("unexpected end");
        switch (*p) {
            case '{': return object();
            case '[': return array();
            case '"': return Json(string());
            case 't': expect("true"); return Json(true);
            case 'f': expect("false"); return Json(false);
            case 'n': expect("null"); return Json();
            default: return number();
        }
    }
    void expect(const char* lit) { size_t n = std::strlen(lit); if (size_t(end - p) < n || std::strncmp(p, lit, n) != 0) fail("bad literal"); p += n; }
    std::string string() {
        std::string s; ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                ++p; if (p >= end) fail("bad escape");
                switch (*p) { case 'n': s += '\n'; break; case 'r': s += '\r'; break; case 't': s += '\t'; break; case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                              case 'u': fail("\\u escapes are not supported (nor by the reference, src/JSON.cpp:152-155)"); default: s += *p; }
                ++p;
            } else s += *p++;
        }
        if (p >= end) fail("unterminated string");
        ++p; return s;
    }
    Json number() { // reference src/JSON.cpp:167-192
        const char* b = p; bool isFloat = false;
        while (p < end && (*p == '-' || *p == '+' || (*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E')) { if (*p == '.' || *p == 'e' || *p == 'E') isFloat = true; ++p; }
        if (p == b) fail("bad number");
        if (isFloat) { float f = 0.f; std::from_chars(b, p, f); return Json(f); }
        int i = 0; std::from_chars(b, p, i); return Json(i);
    }
    Json array() {
        Json a = Json::array(); ++p; ws();
        if (p < end && *p == ']') { ++p; return a; }
        for (;;) { a._items.push_back(value()); ws(); if (p >= end) fail("unterminated array"); if (*p == ',') { ++p; continue; } if (*p == ']') { ++p; break; } fail("expected , or ]"); }
        return a;
    }
    Json object() {
        Json o = Json::object(); ++p; ws();
        if (p < end && *p == '}') { ++p; return o; }
        for (;;) {
            ws(); if (p >= end || *p != '"') fail("expected key");
            std::string k = string(); ws();
            if (p >= end || *p != ':') fail("expected :");
            ++p;
            Json v = value();
            o[k] = v; ws();
            if (p >= end) fail("unterminated object");
            if (*p == ',') { ++p; continue; }
            if (*p == '}') { ++p; break; }
            fail("expected , or }");
        }
        return o;
    }
};

Json Json::parse(const char* data, size_t length) { JsonParser ps{data, data + length}; return ps.value(); }

static void escape(const std::string& s, std::string& out) {
    out += '"';
    for (char c : s) { if (c == '"' || c == '\\') { out += '\\'; out += c; } else if (c == '\n') out += "\\n"; else if (c == '\t') out += "\\t"; else if (c == '\r') out += "\\r"; else out += c; }
    out += '"';
}
std::string Json::toString() const {
    std::string o;
    switch (_type) {
        case Type::Null: return "null";
        case Type::Bool: return _b ? "true" : "false";
        case Type::Int: return std::to_string(_i);
        case Type::Real: return std::to_string(_f);
        case Type::String: escape(_s, o); return o;
        case Type::Array: o = "["; for (size_t i = 0; i < _items.size(); ++i) { if (i) o += ","; o += _items[i].toString(); } return o + "]";
        case Type::Object: o = "{"; for (size_t i = 0; i < _items.size(); ++i) { if (i) o += ","; escape(_names[i], o); o += ":"; o += _items[i].toString(); } return o + "}";
    }
    return o;
}

} // namespace vkx
