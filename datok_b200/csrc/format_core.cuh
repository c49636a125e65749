// format_core.cuh -- the text half of the TokenWriter (token_writer.go:36-175) on the device, per-item bodies.
//
// The reference's writer appends to one stream while the events arrive.  Here every output item knows its
// place from three exclusive prefix sums, so all items are written in parallel:
//     P_tok[k]  bytes of the surfaces (+ "\n") of the tokens before token k          (TOKENS, :81-84,91-93)
//     P_pos[k]  bytes of the `pos` entries (digits + separator) of the tokens before k   (TOKEN_POS, :133-140)
//     P_sp[j]   bytes of the `sent` entries before entry j                               (SENTENCE_POS, :144-151)
//     P_x[d]    bytes written by the TextEnd calls before text d: per text the `pos` line, the `sent` line,
//               or a single "\n" when no position flag is set                            (:130-167)
// In stream order a token is preceded by the tokens before it, by the SentenceEnd events with sent_tok <= k
// (each a "\n" under SENTENCES, :112-114,119-122) and by the TextEnds of the texts that ended before it:
//     out_token(k)    = [T] P_tok[k]               + [S] #{s: sent_tok[s] <= k} + P_x[text of k]
//     out_sentence(s) = [T] P_tok[sent_tok[s]]     + [S] s                      + P_x[text of s]
//     out_textend(d)  = [T] P_tok[text_tok_end[d]] + [S] text_sent_end[d]       + P_x[d]
// The arrays are the ones the compaction kernels leave in HBM (absolute form, indices relative to the piece).
// Surfaces are copied verbatim: an input with malformed UTF-8 (Go re-encodes those bytes as U+FFFD) is left to
// the host formatter (format.cpp).
//
// __host__ __device__ like the other kernel bodies: tests/emul runs them on the CPU against the oracle's text.
#pragma once
#include <stdint.h>

#include "walk_core.cuh"

namespace datok {

constexpr uint32_t FMT_TILE_SHIFT = 10;            // items per scan tile
constexpr uint32_t FMT_TILE = 1u << FMT_TILE_SHIFT;
constexpr uint32_t FMT_F_TOKENS = 1, FMT_F_SENTENCES = 2, FMT_F_TOKEN_POS = 4, FMT_F_SENTENCE_POS = 8;

// exclusive prefix sum in two levels: P(i) = base[i >> FMT_TILE_SHIFT] + local[i], for i in 0..n (n + 1 entries)
struct FmtScan {
  uint32_t* local;
  unsigned long long* base;
};
DATOK_HD unsigned long long fmt_P(const FmtScan& s, uint32_t i) { return s.base[i >> FMT_TILE_SHIFT] + s.local[i]; }

struct FmtCtx {
  const uint8_t* in;
  const uint32_t* tok_bytes;       // 2 per token
  const int32_t* tok_pos;          // 2 per token (TOKEN_POS)
  const int32_t* sent_pos;         // (SENTENCE_POS)
  const uint32_t* sent_tok;        // per SentenceEnd event: tokens before it
  const uint32_t *text_tok_end, *text_sent_end, *text_sentpos_end;
  uint32_t n_tok, n_sent, n_sentpos, n_text, flags;
  FmtScan ptok, ppos, psp, px;
  uint8_t* out;
};

DATOK_HD uint32_t fmt_digits(int32_t v) {  // strconv.Itoa length
  const uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
  return (v < 0 ? 1u : 0u) + 1u + (u >= 10u) + (u >= 100u) + (u >= 1000u) + (u >= 10000u) + (u >= 100000u) + (u >= 1000000u) +
         (u >= 10000000u) + (u >= 100000000u) + (u >= 1000000000u);
}
// strconv.Itoa(v) followed by `sep`; returns the bytes written
DATOK_HD uint32_t fmt_itoa_sep(uint8_t* dst, int32_t v, uint8_t sep) {
  const uint32_t n = fmt_digits(v);
  uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
  uint32_t i = n;
  do { dst[--i] = (uint8_t)('0' + u % 10u); u /= 10u; } while (u);
  if (v < 0) dst[0] = '-';
  dst[n] = sep;
  return n + 1;
}
// number of elements <= key in the ascending array a[0..n)
DATOK_HD uint32_t fmt_count_le(const uint32_t* a, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] <= key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- item lengths (the values the prefix sums run over) ----
DATOK_HD uint32_t fmt_len_tok(const FmtCtx& c, uint32_t k) {
  return (c.flags & FMT_F_TOKENS) ? c.tok_bytes[2 * k + 1] - c.tok_bytes[2 * k] + 1u : 0u;
}
DATOK_HD uint32_t fmt_len_pos(const FmtCtx& c, uint32_t k) {
  return (c.flags & FMT_F_TOKEN_POS) ? fmt_digits(c.tok_pos[2 * k]) + fmt_digits(c.tok_pos[2 * k + 1]) + 2u : 0u;
}
DATOK_HD uint32_t fmt_len_sp(const FmtCtx& c, uint32_t j) {
  return (c.flags & FMT_F_SENTENCE_POS) ? fmt_digits(c.sent_pos[j]) + 1u : 0u;
}
// what TextEnd d writes (token_writer.go:130-167); needs ppos and psp
DATOK_HD uint32_t fmt_len_text(const FmtCtx& c, uint32_t d) {
  if (!(c.flags & (FMT_F_TOKEN_POS | FMT_F_SENTENCE_POS))) return 1u;
  uint32_t n = 0;
  if (c.flags & FMT_F_TOKEN_POS) {
    const uint32_t t0 = d ? c.text_tok_end[d - 1] : 0u, t1 = c.text_tok_end[d];
    const uint32_t l = (uint32_t)(fmt_P(c.ppos, t1) - fmt_P(c.ppos, t0));
    n += l ? l : 1u;  // (an empty list is outside the parity domain: the reference panics; one "\n" keeps the sizes sane)
  }
  if (c.flags & FMT_F_SENTENCE_POS) {
    const uint32_t p0 = d ? c.text_sentpos_end[d - 1] : 0u, p1 = c.text_sentpos_end[d];
    const uint32_t l = (uint32_t)(fmt_P(c.psp, p1) - fmt_P(c.psp, p0));
    n += l ? l : 1u;
  }
  return n;
}

DATOK_HD unsigned long long fmt_out_text(const FmtCtx& c, uint32_t d) {
  return ((c.flags & FMT_F_TOKENS) ? fmt_P(c.ptok, c.text_tok_end[d]) : 0ull) +
         ((c.flags & FMT_F_SENTENCES) ? (unsigned long long)c.text_sent_end[d] : 0ull) + fmt_P(c.px, d);
}
DATOK_HD unsigned long long fmt_total(const FmtCtx& c) {
  return ((c.flags & FMT_F_TOKENS) ? fmt_P(c.ptok, c.n_tok) : 0ull) + ((c.flags & FMT_F_SENTENCES) ? (unsigned long long)c.n_sent : 0ull) +
         fmt_P(c.px, c.n_text);
}

// ---- writers ----
// token k: its surface (TOKENS) and its two entries of the `pos` line of its text (TOKEN_POS)
DATOK_HD void fmt_write_token(const FmtCtx& c, uint32_t k) {
  const uint32_t d = fmt_count_le(c.text_tok_end, c.n_text, k);  // texts that ended before token k = index of its text
  if (c.flags & FMT_F_TOKENS) {
    const uint32_t sb = (c.flags & FMT_F_SENTENCES) ? fmt_count_le(c.sent_tok, c.n_sent, k) : 0u;
    uint8_t* o = c.out + fmt_P(c.ptok, k) + sb + fmt_P(c.px, d);
    const uint32_t lo = c.tok_bytes[2 * k], hi = c.tok_bytes[2 * k + 1];
    for (uint32_t i = lo; i < hi; i++) *o++ = c.in[i];
    *o = '\n';
  }
  if ((c.flags & FMT_F_TOKEN_POS) && d < c.n_text) {
    const uint32_t t0 = d ? c.text_tok_end[d - 1] : 0u, t1 = c.text_tok_end[d];
    uint8_t* o = c.out + fmt_out_text(c, d) + (fmt_P(c.ppos, k) - fmt_P(c.ppos, t0));
    o += fmt_itoa_sep(o, c.tok_pos[2 * k], ' ');
    fmt_itoa_sep(o, c.tok_pos[2 * k + 1], k + 1 == t1 ? '\n' : ' ');
  }
}
// SentenceEnd event s (SENTENCES): "\n"
DATOK_HD void fmt_write_sentence(const FmtCtx& c, uint32_t s) {
  if (!(c.flags & FMT_F_SENTENCES)) return;
  const uint32_t d = fmt_count_le(c.text_sent_end, c.n_text, s);
  c.out[((c.flags & FMT_F_TOKENS) ? fmt_P(c.ptok, c.sent_tok[s]) : 0ull) + s + fmt_P(c.px, d)] = '\n';
}
// TextEnd d: the "\n" of a writer without position flags, or the "\n" of an empty list
DATOK_HD void fmt_write_text(const FmtCtx& c, uint32_t d) {
  uint8_t* o = c.out + fmt_out_text(c, d);
  if (!(c.flags & (FMT_F_TOKEN_POS | FMT_F_SENTENCE_POS))) { *o = '\n'; return; }
  if (c.flags & FMT_F_TOKEN_POS) {
    const uint32_t t0 = d ? c.text_tok_end[d - 1] : 0u, t1 = c.text_tok_end[d];
    const uint32_t l = (uint32_t)(fmt_P(c.ppos, t1) - fmt_P(c.ppos, t0));
    if (!l) *o = '\n';
    o += l ? l : 1u;
  }
  if (c.flags & FMT_F_SENTENCE_POS) {
    const uint32_t p0 = d ? c.text_sentpos_end[d - 1] : 0u, p1 = c.text_sentpos_end[d];
    if (p0 == p1) *o = '\n';
  }
}
// entry j of the `sent` lists (SENTENCE_POS)
DATOK_HD void fmt_write_sentpos(const FmtCtx& c, uint32_t j) {
  if (!(c.flags & FMT_F_SENTENCE_POS)) return;
  const uint32_t d = fmt_count_le(c.text_sentpos_end, c.n_text, j);
  if (d >= c.n_text) return;  // (entries behind the last TextEnd are never printed)
  const uint32_t p0 = d ? c.text_sentpos_end[d - 1] : 0u, p1 = c.text_sentpos_end[d];
  uint8_t* o = c.out + fmt_out_text(c, d);
  if (c.flags & FMT_F_TOKEN_POS) {
    const uint32_t t0 = d ? c.text_tok_end[d - 1] : 0u, t1 = c.text_tok_end[d];
    const uint32_t l = (uint32_t)(fmt_P(c.ppos, t1) - fmt_P(c.ppos, t0));
    o += l ? l : 1u;
  }
  o += fmt_P(c.psp, j) - fmt_P(c.psp, p0);
  fmt_itoa_sep(o, c.sent_pos[j], j + 1 == p1 ? '\n' : ' ');
}

}  // namespace datok
