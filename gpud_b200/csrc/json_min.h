// json_min.h — the little JSON reading the host side needs (stored ExtraInfo maps, the xidErrorEventDetail payload): strings with
// every escape encoding/json accepts, and skipping over values that are not looked at.  Header-only, no allocation besides the
// output strings.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>

namespace jsonmin {

inline void utf8_put(std::string& o, uint32_t c) {
  if (c < 0x80) o += (char)c;
  else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 63)); }
  else if (c < 0x10000) { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
  else { o += (char)(0xF0 | (c >> 18)); o += (char)(0x80 | ((c >> 12) & 63)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
}
inline void ws(const char*& p) { while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p; }
inline bool hex4(const char*& p, uint32_t* v) {
  uint32_t x = 0;
  for (int i = 0; i < 4; ++i, ++p) {
    const char c = *p;
    if (c >= '0' && c <= '9') x = x * 16 + (uint32_t)(c - '0');
    else if (c >= 'a' && c <= 'f') x = x * 16 + (uint32_t)(c - 'a' + 10);
    else if (c >= 'A' && c <= 'F') x = x * 16 + (uint32_t)(c - 'A' + 10);
    else return false;
  }
  *v = x;
  return true;
}
// p at the opening quote; on success p is past the closing quote and *out holds the decoded UTF-8
inline bool string(const char*& p, std::string* out) {
  if (*p != '"') return false;
  ++p;
  out->clear();
  while (*p && *p != '"') {
    if ((unsigned char)*p < 0x20) return false;
    if (*p != '\\') { *out += *p++; continue; }
    ++p;
    switch (*p) {
      case '"': *out += '"'; ++p; break;
      case '\\': *out += '\\'; ++p; break;
      case '/': *out += '/'; ++p; break;
      case 'b': *out += '\b'; ++p; break;
      case 'f': *out += '\f'; ++p; break;
      case 'n': *out += '\n'; ++p; break;
      case 'r': *out += '\r'; ++p; break;
      case 't': *out += '\t'; ++p; break;
      case 'u': {
        ++p;
        uint32_t c;
        if (!hex4(p, &c)) return false;
        if (c >= 0xD800 && c < 0xDC00 && p[0] == '\\' && p[1] == 'u') {       // surrogate pair; a lone half becomes U+FFFD like encoding/json
          const char* q = p + 2;
          uint32_t lo;
          if (hex4(q, &lo) && lo >= 0xDC00 && lo < 0xE000) { c = 0x10000 + ((c - 0xD800) << 10) + (lo - 0xDC00); p = q; }
          else c = 0xFFFD;
        } else if (c >= 0xD800 && c < 0xE000) c = 0xFFFD;
        utf8_put(*out, c);
        break;
      }
      default: return false;
    }
  }
  if (*p != '"') return false;
  ++p;
  return true;
}
// skip one value of any type; p ends past it
inline bool skip(const char*& p, int depth = 0) {
  ws(p);
  if (depth > 64) return false;
  if (*p == '"') { std::string t; return string(p, &t); }
  if (*p == '{' || *p == '[') {
    const char close = *p == '{' ? '}' : ']';
    const bool obj = *p == '{';
    ++p;
    ws(p);
    if (*p == close) { ++p; return true; }
    for (;;) {
      ws(p);
      if (obj) {
        std::string k;
        if (!string(p, &k)) return false;
        ws(p);
        if (*p++ != ':') return false;
      }
      if (!skip(p, depth + 1)) return false;
      ws(p);
      if (*p == ',') { ++p; continue; }
      if (*p == close) { ++p; return true; }
      return false;
    }
  }
  const char* q = p;
  while (*p && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') ++p;
  return p > q;
}

}  // namespace jsonmin
