// Minimal value tree shared by the msgpack (engine-core IPC) and JSON (HTTP) codecs of the native front-end.
// The engine core speaks msgspec-msgpack: structs are string-keyed maps with defaults omitted
// (vllm_b200/core/request.py EngineCoreRequest / EngineCoreOutputs), outputs arrive as [client_index, {...}].
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace wire {

struct Val {
  enum T { NIL, BOOL, INT, FLT, STR, BIN, ARR, MAP } t = NIL;
  bool b = false;
  int64_t i = 0;
  double f = 0;
  std::string s;
  std::vector<Val> a;
  std::vector<std::pair<std::string, Val>> m;

  const Val* get(const char* k) const {
    if (t != MAP) return nullptr;
    for (auto& kv : m) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  bool is_num() const { return t == INT || t == FLT; }
  double num() const { return t == INT ? double(i) : f; }
  int64_t integer() const { return t == INT ? i : int64_t(f); }
  bool truthy() const { return (t == BOOL && b) || (t == INT && i) || (t == FLT && f != 0); }
};

// ------------------------------------------------------------------ msgpack reader
struct MpReader {
  const uint8_t* p; const uint8_t* e;
  MpReader(const void* d, size_t n) : p((const uint8_t*)d), e((const uint8_t*)d + n) {}
  void need(size_t n) { if (size_t(e - p) < n) throw std::runtime_error("msgpack: truncated"); }
  uint64_t be(int n) { need(n); uint64_t v = 0; for (int k = 0; k < n; k++) v = (v << 8) | *p++; return v; }
  std::string bytes(size_t n) { need(n); std::string r((const char*)p, n); p += n; return r; }
  Val arr(size_t n) { Val v; v.t = Val::ARR; v.a.reserve(n); for (size_t k = 0; k < n; k++) v.a.push_back(read()); return v; }
  Val map(size_t n) {
    Val v; v.t = Val::MAP; v.m.reserve(n);
    for (size_t k = 0; k < n; k++) { Val key = read(); std::string ks = key.t == Val::STR ? key.s : std::to_string(key.i); v.m.emplace_back(std::move(ks), read()); }
    return v;
  }
  Val read() {
    need(1);
    uint8_t c = *p++;
    Val v;
    if (c <= 0x7f) { v.t = Val::INT; v.i = c; return v; }
    if (c >= 0xe0) { v.t = Val::INT; v.i = int8_t(c); return v; }
    if ((c & 0xf0) == 0x80) return map(c & 0x0f);
    if ((c & 0xf0) == 0x90) return arr(c & 0x0f);
    if ((c & 0xe0) == 0xa0) { v.t = Val::STR; v.s = bytes(c & 0x1f); return v; }
    switch (c) {
      case 0xc0: return v;
      case 0xc2: v.t = Val::BOOL; v.b = false; return v;
      case 0xc3: v.t = Val::BOOL; v.b = true; return v;
      case 0xc4: v.t = Val::BIN; v.s = bytes(be(1)); return v;
      case 0xc5: v.t = Val::BIN; v.s = bytes(be(2)); return v;
      case 0xc6: v.t = Val::BIN; v.s = bytes(be(4)); return v;
      case 0xc7: { size_t n = be(1); be(1); v.t = Val::BIN; v.s = bytes(n); return v; }
      case 0xc8: { size_t n = be(2); be(1); v.t = Val::BIN; v.s = bytes(n); return v; }
      case 0xc9: { size_t n = be(4); be(1); v.t = Val::BIN; v.s = bytes(n); return v; }
      case 0xca: { uint32_t u = be(4); float f; memcpy(&f, &u, 4); v.t = Val::FLT; v.f = f; return v; }
      case 0xcb: { uint64_t u = be(8); double f; memcpy(&f, &u, 8); v.t = Val::FLT; v.f = f; return v; }
      case 0xcc: v.t = Val::INT; v.i = be(1); return v;
      case 0xcd: v.t = Val::INT; v.i = be(2); return v;
      case 0xce: v.t = Val::INT; v.i = be(4); return v;
      case 0xcf: v.t = Val::INT; v.i = int64_t(be(8)); return v;
      case 0xd0: v.t = Val::INT; v.i = int8_t(be(1)); return v;
      case 0xd1: v.t = Val::INT; v.i = int16_t(be(2)); return v;
      case 0xd2: v.t = Val::INT; v.i = int32_t(be(4)); return v;
      case 0xd3: v.t = Val::INT; v.i = int64_t(be(8)); return v;
      case 0xd4: case 0xd5: case 0xd6: case 0xd7: case 0xd8: { be(1); v.t = Val::BIN; v.s = bytes(size_t(1) << (c - 0xd4)); return v; }
      case 0xd9: v.t = Val::STR; v.s = bytes(be(1)); return v;
      case 0xda: v.t = Val::STR; v.s = bytes(be(2)); return v;
      case 0xdb: v.t = Val::STR; v.s = bytes(be(4)); return v;
      case 0xdc: return arr(be(2));
      case 0xdd: return arr(be(4));
      case 0xde: return map(be(2));
      case 0xdf: return map(be(4));
    }
    throw std::runtime_error("msgpack: bad type byte");
  }
};

// ------------------------------------------------------------------ msgpack writer
struct MpWriter {
  std::string o;
  void raw(uint64_t v, int n) { for (int k = n - 1; k >= 0; k--) o.push_back(char((v >> (8 * k)) & 0xff)); }
  void nil() { o.push_back(char(0xc0)); }
  void boolean(bool b) { o.push_back(char(b ? 0xc3 : 0xc2)); }
  void integer(int64_t v) {
    if (v >= 0) {
      if (v < 128) o.push_back(char(v));
      else if (v < 256) { o.push_back(char(0xcc)); raw(v, 1); }
      else if (v < 65536) { o.push_back(char(0xcd)); raw(v, 2); }
      else if (v < (int64_t(1) << 32)) { o.push_back(char(0xce)); raw(v, 4); }
      else { o.push_back(char(0xcf)); raw(v, 8); }
    } else {
      if (v >= -32) o.push_back(char(v));
      else if (v >= -128) { o.push_back(char(0xd0)); raw(uint8_t(v), 1); }
      else if (v >= -32768) { o.push_back(char(0xd1)); raw(uint16_t(v), 2); }
      else if (v >= -(int64_t(1) << 31)) { o.push_back(char(0xd2)); raw(uint32_t(v), 4); }
      else { o.push_back(char(0xd3)); raw(uint64_t(v), 8); }
    }
  }
  void real(double d) { o.push_back(char(0xcb)); uint64_t u; memcpy(&u, &d, 8); raw(u, 8); }
  void str(const std::string& s) {
    size_t n = s.size();
    if (n < 32) o.push_back(char(0xa0 | n));
    else if (n < 256) { o.push_back(char(0xd9)); raw(n, 1); }
    else if (n < 65536) { o.push_back(char(0xda)); raw(n, 2); }
    else { o.push_back(char(0xdb)); raw(n, 4); }
    o += s;
  }
  void array(size_t n) {
    if (n < 16) o.push_back(char(0x90 | n));
    else if (n < 65536) { o.push_back(char(0xdc)); raw(n, 2); }
    else { o.push_back(char(0xdd)); raw(n, 4); }
  }
  void map(size_t n) {
    if (n < 16) o.push_back(char(0x80 | n));
    else if (n < 65536) { o.push_back(char(0xde)); raw(n, 2); }
    else { o.push_back(char(0xdf)); raw(n, 4); }
  }
  void value(const Val& v) {          // JSON-born values forwarded verbatim (e.g. kv_transfer_params)
    switch (v.t) {
      case Val::NIL: nil(); break;
      case Val::BOOL: boolean(v.b); break;
      case Val::INT: integer(v.i); break;
      case Val::FLT: real(v.f); break;
      case Val::STR: case Val::BIN: str(v.s); break;
      case Val::ARR: array(v.a.size()); for (auto& x : v.a) value(x); break;
      case Val::MAP: map(v.m.size()); for (auto& kv : v.m) { str(kv.first); value(kv.second); } break;
    }
  }
};

// ------------------------------------------------------------------ JSON
struct JsonReader {
  const char* p; const char* e;
  JsonReader(const std::string& s) : p(s.data()), e(s.data() + s.size()) {}
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m); }
  static void utf8(std::string& o, uint32_t c) {
    if (c < 0x80) o.push_back(char(c));
    else if (c < 0x800) { o.push_back(char(0xc0 | (c >> 6))); o.push_back(char(0x80 | (c & 0x3f))); }
    else if (c < 0x10000) { o.push_back(char(0xe0 | (c >> 12))); o.push_back(char(0x80 | ((c >> 6) & 0x3f))); o.push_back(char(0x80 | (c & 0x3f))); }
    else { o.push_back(char(0xf0 | (c >> 18))); o.push_back(char(0x80 | ((c >> 12) & 0x3f))); o.push_back(char(0x80 | ((c >> 6) & 0x3f))); o.push_back(char(0x80 | (c & 0x3f))); }
  }
  uint32_t hex4() {
    if (e - p < 4) fail("bad \\u escape");
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) {
      char c = *p++; v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0'; else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10; else fail("bad hex digit");
    }
    return v;
  }
  std::string string() {
    if (p >= e || *p != '"') fail("expected string");
    p++;
    std::string o;
    while (true) {
      if (p >= e) fail("unterminated string");
      char c = *p++;
      if (c == '"') break;
      if (c != '\\') { o.push_back(c); continue; }
      if (p >= e) fail("bad escape");
      char x = *p++;
      switch (x) {
        case 'n': o.push_back('\n'); break; case 't': o.push_back('\t'); break; case 'r': o.push_back('\r'); break;
        case 'b': o.push_back('\b'); break; case 'f': o.push_back('\f'); break;
        case 'u': {
          uint32_t c1 = hex4();
          if (c1 >= 0xd800 && c1 < 0xdc00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') { p += 2; uint32_t c2 = hex4(); c1 = 0x10000 + ((c1 - 0xd800) << 10) + (c2 - 0xdc00); }
          utf8(o, c1); break;
        }
        default: o.push_back(x);
      }
    }
    return o;
  }
  Val read(int depth = 0) {
    if (depth > 64) fail("too deep");
    ws();
    if (p >= e) fail("unexpected end");
    Val v;
    char c = *p;
    if (c == '{') {
      p++; v.t = Val::MAP; ws();
      if (p < e && *p == '}') { p++; return v; }
      while (true) {
        ws(); std::string k = string(); ws();
        if (p >= e || *p != ':') fail("expected ':'");
        p++;
        v.m.emplace_back(std::move(k), read(depth + 1)); ws();
        if (p < e && *p == ',') { p++; continue; }
        if (p < e && *p == '}') { p++; return v; }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      p++; v.t = Val::ARR; ws();
      if (p < e && *p == ']') { p++; return v; }
      while (true) {
        v.a.push_back(read(depth + 1)); ws();
        if (p < e && *p == ',') { p++; continue; }
        if (p < e && *p == ']') { p++; return v; }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') { v.t = Val::STR; v.s = string(); return v; }
    if (e - p >= 4 && !strncmp(p, "true", 4)) { p += 4; v.t = Val::BOOL; v.b = true; return v; }
    if (e - p >= 5 && !strncmp(p, "false", 5)) { p += 5; v.t = Val::BOOL; return v; }
    if (e - p >= 4 && !strncmp(p, "null", 4)) { p += 4; return v; }
    const char* s = p; bool flt = false;
    if (p < e && (*p == '-' || *p == '+')) p++;
    while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) { if (*p == '.' || *p == 'e' || *p == 'E') flt = true; p++; }
    if (p == s) fail("unexpected character");
    std::string num(s, p - s);
    if (flt) { v.t = Val::FLT; v.f = strtod(num.c_str(), nullptr); } else { v.t = Val::INT; v.i = strtoll(num.c_str(), nullptr, 10); }
    return v;
  }
};

inline void json_escape(std::string& o, const std::string& s) {
  o.push_back('"');
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break; case '\\': o += "\\\\"; break; case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break; case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; } else o.push_back(char(c));
    }
  }
  o.push_back('"');
}

inline void json_write(std::string& o, const Val& v) {
  switch (v.t) {
    case Val::NIL: o += "null"; break;
    case Val::BOOL: o += v.b ? "true" : "false"; break;
    case Val::INT: o += std::to_string(v.i); break;
    case Val::FLT: { if (!std::isfinite(v.f)) { o += "null"; break; } char b[40]; snprintf(b, sizeof b, "%.17g", v.f); o += b; break; }
    case Val::STR: case Val::BIN: json_escape(o, v.s); break;
    case Val::ARR: { o.push_back('['); bool f = true; for (auto& x : v.a) { if (!f) o.push_back(','); f = false; json_write(o, x); } o.push_back(']'); break; }
    case Val::MAP: { o.push_back('{'); bool f = true; for (auto& kv : v.m) { if (!f) o.push_back(','); f = false; json_escape(o, kv.first); o.push_back(':'); json_write(o, kv.second); } o.push_back('}'); break; }
  }
}

}  // namespace wire
