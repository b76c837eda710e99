// Byte-level BPE encoder for the native front-end: special-token splitting, the GPT-4 / Llama-3 / Qwen-2 / GPT-2
// pre-tokenisation patterns as a hand-written scanner (no regex engine), rank-ordered pair merging.
// Role of the reference's rust/src/tokenizer crate (HF `tokenizers` / tiktoken bindings).  The Unicode classes
// (\p{L}, \p{N}, \s) come from the table file written by the Python launcher, so both sides agree by construction.
//
//   style 0: (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+
//   style 1: same with \p{N} (one digit per piece)                                                         (Qwen-2)
//   style 2: 's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+                    (GPT-2)
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

namespace tok {

struct Ranges {
  std::vector<std::pair<uint32_t, uint32_t>> r;   // sorted, inclusive
  bool has(uint32_t c) const {
    size_t lo = 0, hi = r.size();
    while (lo < hi) { size_t m = (lo + hi) / 2; if (c < r[m].first) hi = m; else if (c > r[m].second) lo = m + 1; else return true; }
    return false;
  }
};

struct Encoder {
  int style = 0;
  bool add_bos = false;
  int64_t bos_id = -1;
  Ranges L, N, S;
  std::unordered_map<std::string, std::pair<uint32_t, uint32_t>> ranks;   // bytes → (priority, id)
  std::vector<std::pair<std::string, uint32_t>> specials;                  // sorted by length desc
  // HF `tokenizers` semantics: explicit (left id, right id) → (rank, merged id) merges; empty → tiktoken semantics
  // (any adjacent pair whose concatenation is a known token may merge, priority = its rank)
  std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> pairs;
  bool whole_piece_lookup = true;
  bool ready = false;

  // ------------------------------------------------------------------ utf-8
  struct Cp { uint32_t c; uint32_t off, len; };
  static std::vector<Cp> decode(const std::string& s) {
    std::vector<Cp> out; out.reserve(s.size());
    size_t i = 0, n = s.size();
    while (i < n) {
      unsigned char b = s[i];
      uint32_t c; size_t len;
      if (b < 0x80) { c = b; len = 1; }
      else if ((b >> 5) == 6 && i + 1 < n) { c = ((b & 0x1f) << 6) | (s[i + 1] & 0x3f); len = 2; }
      else if ((b >> 4) == 14 && i + 2 < n) { c = ((b & 0x0f) << 12) | ((s[i + 1] & 0x3f) << 6) | (s[i + 2] & 0x3f); len = 3; }
      else if ((b >> 3) == 30 && i + 3 < n) { c = ((b & 0x07) << 18) | ((s[i + 1] & 0x3f) << 12) | ((s[i + 2] & 0x3f) << 6) | (s[i + 3] & 0x3f); len = 4; }
      else { c = 0xFFFD; len = 1; }
      out.push_back({c, (uint32_t)i, (uint32_t)len});
      i += len;
    }
    return out;
  }

  bool isL(uint32_t c) const { return L.has(c); }
  bool isN(uint32_t c) const { return N.has(c); }
  bool isS(uint32_t c) const { return S.has(c); }
  static bool isNL(uint32_t c) { return c == '\r' || c == '\n'; }
  static uint32_t lower(uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

  // length (in code points) of a contraction at i, or 0
  size_t contraction(const std::vector<Cp>& t, size_t i, bool icase) const {
    if (t[i].c != '\'' || i + 1 >= t.size()) return 0;
    auto ch = [&](size_t k) { return icase ? lower(t[k].c) : t[k].c; };
    uint32_t a = ch(i + 1);
    uint32_t b = i + 2 < t.size() ? ch(i + 2) : 0;
    if (a == 's' || a == 't') return 2;
    if (a == 'r' && b == 'e') return 3;
    if (a == 'v' && b == 'e') return 3;
    if (a == 'm') return 2;
    if (a == 'l' && b == 'l') return 3;
    if (a == 'd') return 2;
    return 0;
  }

  // end index (exclusive) of the piece starting at i
  size_t next_piece(const std::vector<Cp>& t, size_t i) const {
    const size_t n = t.size();
    auto other = [&](size_t k) { return !isS(t[k].c) && !isL(t[k].c) && !isN(t[k].c); };
    if (size_t k = contraction(t, i, style != 2)) return i + k;
    if (style != 2) {
      // [^\r\n\p{L}\p{N}]?\p{L}+
      size_t j = i;
      if (isL(t[j].c)) { while (j < n && isL(t[j].c)) j++; return j; }
      if (!isNL(t[j].c) && !isN(t[j].c) && j + 1 < n && isL(t[j + 1].c)) { j++; while (j < n && isL(t[j].c)) j++; return j; }
      // \p{N}{1,3} / \p{N}
      if (isN(t[i].c)) { size_t lim = style == 0 ? 3 : 1; j = i; while (j < n && j - i < lim && isN(t[j].c)) j++; return j; }
      //  ?[^\s\p{L}\p{N}]+[\r\n]*
      size_t st = n;
      if (t[i].c == ' ' && i + 1 < n && other(i + 1)) st = i + 1; else if (other(i)) st = i;
      if (st != n) { j = st; while (j < n && other(j)) j++; while (j < n && isNL(t[j].c)) j++; return j; }
    } else {
      size_t st = (t[i].c == ' ' && i + 1 < n) ? i + 1 : i;
      size_t j;
      if (isL(t[st].c)) { j = st; while (j < n && isL(t[j].c)) j++; return j; }
      if (isN(t[st].c)) { j = st; while (j < n && isN(t[j].c)) j++; return j; }
      if (other(st)) { j = st; while (j < n && other(j)) j++; return j; }
    }
    // whitespace alternatives
    if (isS(t[i].c)) {
      size_t e = i; while (e < n && isS(t[e].c)) e++;
      if (style != 2) {                       // \s*[\r\n]+ : up to and including the last newline of the run
        size_t last = n;
        for (size_t k = i; k < e; k++) if (isNL(t[k].c)) last = k;
        if (last != n) return last + 1;
      }
      if (e == n) return e;                   // \s+(?!\S) at end of text
      if (e - 1 > i) return e - 1;            // leave one whitespace char for the next piece
      return e;                               // \s+
    }
    return i + 1;
  }

  // ------------------------------------------------------------------ BPE over one piece
  void bpe_pairs(const std::string& piece, std::vector<int64_t>& out) const {
    std::vector<uint32_t> ids; ids.reserve(piece.size());
    for (unsigned char c : piece) {
      auto it = ranks.find(std::string(1, (char)c));
      ids.push_back(it != ranks.end() ? it->second.second : 0);
    }
    while (ids.size() > 1) {
      uint32_t best = UINT32_MAX, merged = 0; size_t at = 0;
      for (size_t k = 0; k + 1 < ids.size(); k++) {
        auto it = pairs.find((uint64_t(ids[k]) << 32) | ids[k + 1]);
        if (it != pairs.end() && it->second.first < best) { best = it->second.first; merged = it->second.second; at = k; }
      }
      if (best == UINT32_MAX) break;
      ids[at] = merged;
      ids.erase(ids.begin() + at + 1);
    }
    for (auto id : ids) out.push_back(id);
  }

  void bpe(const std::string& piece, std::vector<int64_t>& out) const {
    auto whole = ranks.find(piece);
    if (whole_piece_lookup && whole != ranks.end()) { out.push_back(whole->second.second); return; }
    if (!pairs.empty()) { bpe_pairs(piece, out); return; }
    struct Sym { uint32_t start, len; };
    std::vector<Sym> sy; sy.reserve(piece.size());
    for (uint32_t k = 0; k < piece.size(); k++) sy.push_back({k, 1});
    while (sy.size() > 1) {
      uint32_t best = UINT32_MAX; size_t at = 0;
      for (size_t k = 0; k + 1 < sy.size(); k++) {
        auto it = ranks.find(piece.substr(sy[k].start, sy[k].len + sy[k + 1].len));
        if (it != ranks.end() && it->second.first < best) { best = it->second.first; at = k; }
      }
      if (best == UINT32_MAX) break;
      sy[at].len += sy[at + 1].len;
      sy.erase(sy.begin() + at + 1);
    }
    for (auto& s : sy) {
      auto it = ranks.find(piece.substr(s.start, s.len));
      out.push_back(it != ranks.end() ? (int64_t)it->second.second : 0);
    }
  }

  void encode_plain(const std::string& text, std::vector<int64_t>& out) const {
    if (text.empty()) return;
    auto t = decode(text);
    size_t i = 0;
    while (i < t.size()) {
      size_t j = next_piece(t, i);
      if (j <= i) j = i + 1;
      size_t b0 = t[i].off, b1 = j < t.size() ? t[j].off : text.size();
      bpe(text.substr(b0, b1 - b0), out);
      i = j;
    }
  }

  std::vector<int64_t> encode(const std::string& text, bool with_bos) const {
    std::vector<int64_t> out;
    if (with_bos && add_bos && bos_id >= 0) out.push_back(bos_id);
    size_t pos = 0, seg = 0;
    while (pos < text.size()) {
      bool hit = false;
      if (!specials.empty() && text[pos] == '<') {
        for (auto& sp : specials) {
          if (text.compare(pos, sp.first.size(), sp.first) == 0) {
            encode_plain(text.substr(seg, pos - seg), out);
            out.push_back(sp.second);
            pos += sp.first.size(); seg = pos; hit = true;
            break;
          }
        }
      }
      if (!hit) pos++;
    }
    encode_plain(text.substr(seg), out);
    return out;
  }

  // ------------------------------------------------------------------ table file section
  // u32 style, u32 add_bos, i64 bos_id, 3 x (u32 n, n x (u32 lo, u32 hi)), u32 n_special (u32 id, u32 len, bytes),
  // u32 n_rank (u32 priority, u32 id, u32 len, bytes), [u32 n_pairs, u32 whole_piece_lookup, n_pairs x (u32 left, right, rank, merged)]
  bool load(FILE* f) {
    auto rd32 = [&](uint32_t& v) { return fread(&v, 4, 1, f) == 1; };
    uint32_t st, ab; int64_t bos;
    if (!rd32(st) || !rd32(ab) || fread(&bos, 8, 1, f) != 1) return false;
    style = (int)st; add_bos = ab != 0; bos_id = bos;
    for (Ranges* R : {&L, &N, &S}) {
      uint32_t n; if (!rd32(n)) return false;
      R->r.resize(n);
      for (auto& p : R->r) if (!rd32(p.first) || !rd32(p.second)) return false;
    }
    uint32_t ns; if (!rd32(ns)) return false;
    for (uint32_t k = 0; k < ns; k++) {
      uint32_t id, len; if (!rd32(id) || !rd32(len)) return false;
      std::string s(len, '\0'); if (len && fread(&s[0], 1, len, f) != len) return false;
      specials.emplace_back(std::move(s), id);
    }
    std::sort(specials.begin(), specials.end(), [](auto& a, auto& b) { return a.first.size() > b.first.size(); });
    uint32_t nr; if (!rd32(nr)) return false;
    ranks.reserve(nr * 2);
    for (uint32_t k = 0; k < nr; k++) {
      uint32_t pr, id, len; if (!rd32(pr) || !rd32(id) || !rd32(len)) return false;
      std::string s(len, '\0'); if (len && fread(&s[0], 1, len, f) != len) return false;
      auto it = ranks.find(s);
      if (it == ranks.end() || pr < it->second.first) ranks[s] = {pr, id};
    }
    uint32_t np = 0, wl = 1;
    if (rd32(np)) {                                                        // optional pair-merge section
      if (!rd32(wl)) return false;
      whole_piece_lookup = wl != 0;
      pairs.reserve(np * 2);
      for (uint32_t k = 0; k < np; k++) {
        uint32_t l, r, rank, merged; if (!rd32(l) || !rd32(r) || !rd32(rank) || !rd32(merged)) return false;
        uint64_t key = (uint64_t(l) << 32) | r;
        if (!pairs.count(key)) pairs[key] = {rank, merged};
      }
    }
    ready = true;
    return true;
  }
};

}  // namespace tok
