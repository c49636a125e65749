// format.cpp -- host half of the TokenWriter (token_writer.go:36-175): turns the
// offset arrays of a result into the exact text NewTokenWriter(w, flags) writes,
// or replays the events into caller-supplied callbacks (custom TokenWriters,
// token_writer.go:27-33).  Pure formatting: all boundaries and offsets were
// computed on the GPU.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/datok_b200.h"
#include "model.hpp"

namespace {

static const char kDigits2[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";

struct Sink {
  uint8_t* dst;      // nullptr: only count
  size_t cap, n = 0;
  const uint8_t* src_end = nullptr;  // end of the input buffer: lets short surfaces be copied as one 16-byte block
  // in_input: p points into the input buffer (only then may the 16-byte block copy read past p + len)
  inline void put(const uint8_t* p, size_t len, bool in_input = true) {
    if (n + len <= cap) {
      uint8_t* d = dst + n;
      if (len <= 16) {  // token surfaces are short
        if (in_input && n + 16 <= cap && p + 16 <= src_end) std::memcpy(d, p, 16);  // fixed size: two moves; the tail is overwritten next
        else for (size_t i = 0; i < len; i++) d[i] = p[i];
      } else {
        std::memcpy(d, p, len);
      }
    } else if (n < cap) {
      std::memcpy(dst + n, p, cap - n);
    }
    n += len;
  }
  inline void byte(uint8_t b) {
    if (n < cap) dst[n] = b;
    n++;
  }
  inline void itoa(int32_t v) {  // strconv.Itoa
    uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    if (!dst) {  // counting pass: digits only
      n += (v < 0) + 1 + (u >= 10) + (u >= 100) + (u >= 1000) + (u >= 10000) + (u >= 100000) + (u >= 1000000) +
           (u >= 10000000) + (u >= 100000000) + (u >= 1000000000);
      return;
    }
    uint8_t tmp[12];
    int i = 12;
    while (u >= 100) { const uint32_t q = u / 100, r = u - q * 100; tmp[--i] = (uint8_t)kDigits2[2 * r + 1]; tmp[--i] = (uint8_t)kDigits2[2 * r]; u = q; }
    if (u >= 10) { tmp[--i] = (uint8_t)kDigits2[2 * u + 1]; tmp[--i] = (uint8_t)kDigits2[2 * u]; }
    else tmp[--i] = (uint8_t)('0' + u);
    if (v < 0) tmp[--i] = '-';
    const size_t len = (size_t)(12 - i);
    if (n + len <= cap) { for (size_t k = 0; k < len; k++) dst[n + k] = tmp[i + k]; n += len; }
    else { for (size_t k = 0; k < len; k++) byte(tmp[i + k]); }
  }
};

// string(buf[offset:]) -- Go re-encodes runes, so malformed bytes come out as U+FFFD
void put_surface(Sink& s, const uint8_t* in, size_t lo, size_t hi, bool reencode) {
  if (!reencode) { s.put(in + lo, hi - lo); return; }
  static const uint8_t kRep[3] = {0xEF, 0xBF, 0xBD};
  size_t p = lo;
  while (p < hi) {
    int w;
    int32_t r = datok::decode_rune(in + p, hi - p, &w);
    if (r == 0xFFFD && w == 1) s.put(kRep, 3, false); else s.put(in + p, (size_t)w);
    p += (size_t)w;
  }
}

// Token spans of a result in stream order, from the absolute arrays or from the DATOK_COMPACT deltas.
struct TokenCursor {
  const datok_view* v;
  uint64_t k = 0, text = 0;
  uint64_t byte_end = 0;   // end of the previous token of this text, or the start of the text
  int64_t rune_end = 0;
  uint64_t esc = 0;        // next entry of the DATOK_COMPACT8 escape list
  uint32_t lo = 0, hi = 0;
  int32_t ps = 0, pe = 0;
  explicit TokenCursor(const datok_view* view) : v(view) {}
  // positions the cursor on the first token of text d
  void seek_text(uint64_t d) {
    k = d ? v->text_tok_end[d - 1] : 0;
    text = d;
    byte_end = d ? v->text_byte_end[d - 1] : 0;
    rune_end = 0;
    if (v->tok_delta8) {  // first escape entry of a token >= k
      uint64_t a = 0, b = v->n_esc;
      while (a < b) { const uint64_t mid = (a + b) / 2; if (v->tok_esc[2 * mid] < k) a = mid + 1; else b = mid; }
      esc = a;
    }
  }
  void next() {  // token k
    if (v->tok_delta || v->tok_delta8) {
      while (text < v->n_texts && k >= v->text_tok_end[text]) {  // a new text: both cursors restart
        byte_end = v->text_byte_end[text];
        rune_end = 0;
        text++;
      }
      uint32_t d[4];
      if (v->tok_delta) {
        const uint16_t* p = v->tok_delta + 4 * k;
        d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
      } else {
        const uint8_t* p = v->tok_delta8 + 4 * k;
        d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
        if ((d[0] == 255) | (d[1] == 255) | (d[2] == 255) | (d[3] == 255)) {  // rare: values from the escape list
          while (esc < v->n_esc && v->tok_esc[2 * esc] < k) esc++;
          for (; esc < v->n_esc && v->tok_esc[2 * esc] == k; esc++) d[(v->tok_esc[2 * esc + 1] >> 16) & 3u] = v->tok_esc[2 * esc + 1] & 0xFFFFu;
        }
      }
      lo = (uint32_t)(byte_end + d[0]); hi = lo + d[1];
      ps = (int32_t)(rune_end + d[2]); pe = ps + (int32_t)d[3];
      byte_end = hi; rune_end = pe;
    } else {
      if (v->tok_bytes) { lo = v->tok_bytes[2 * k]; hi = v->tok_bytes[2 * k + 1]; }
      if (v->tok_pos) { ps = v->tok_pos[2 * k]; pe = v->tok_pos[2 * k + 1]; }
    }
    k++;
  }
};

}  // namespace

extern "C" {

int datok_expand(const datok_result* r, uint32_t* tok_bytes, int32_t* tok_pos) {
  const datok_view* v = datok_result_view(r);
  if (!v || (!v->tok_delta && !v->tok_delta8 && v->n_tokens)) return DATOK_ERR_INVALID_ARG;
  TokenCursor c(v);
  for (uint64_t k = 0; k < v->n_tokens; k++) {
    c.next();
    if (tok_bytes) { tok_bytes[2 * k] = c.lo; tok_bytes[2 * k + 1] = c.hi; }
    if (tok_pos) { tok_pos[2 * k] = c.ps; tok_pos[2 * k + 1] = c.pe; }
  }
  return DATOK_OK;
}

}  // extern "C"

namespace {

// Output of the texts [d0, d1) (and, for the last range, of the events after the last TextEnd).
// Texts are independent: token, sentence and `sent` indices restart from the per-text bounds.
void format_range(const datok_view* v, const uint8_t* in, size_t n_in, uint32_t flags, uint64_t d0, uint64_t d1, bool tail,
                  Sink& s) {
  s.src_end = in + n_in;
  const bool tokens = flags & DATOK_TOKENS, sentences = flags & DATOK_SENTENCES;
  const bool tpos = flags & DATOK_TOKEN_POS, spos = flags & DATOK_SENTENCE_POS;
  const bool re = v->has_invalid_utf8 != 0;
  uint64_t tok = d0 ? v->text_tok_end[d0 - 1] : 0, sen = d0 ? v->text_sent_end[d0 - 1] : 0,
           sp = d0 ? v->text_sentpos_end[d0 - 1] : 0;
  TokenCursor cur(v);
  cur.seek_text(d0);
  // the `pos` line (token_writer.go:131-143) of a delta-coded result: a second cursor decodes the text's
  // rune offsets again (cheaper than buffering them while the surfaces are written)
  const bool delta_pos = tpos && (v->tok_delta || v->tok_delta8);
  auto emit_tokens = [&](uint64_t upto) {
    if (tokens)
      for (; tok < upto; tok++) {
        cur.next();
        put_surface(s, in, cur.lo, cur.hi, re);
        s.byte('\n');
      }
    tok = upto;
  };
  auto emit_sentences = [&](uint64_t upto) {
    if (sentences) {
      for (; sen < upto; sen++) {
        emit_tokens(v->sent_tok[sen]);
        s.byte('\n');  // token_writer.go:112-114,119-122
      }
    }
    sen = upto;
  };
  for (uint64_t d = d0; d < d1; d++) {
    const uint64_t t0 = tok, t1 = v->text_tok_end[d];
    emit_sentences(v->text_sent_end[d]);
    emit_tokens(t1);
    if (tpos || spos) {  // token_writer.go:131-159
      if (tpos) {
        if (delta_pos) {
          TokenCursor pc(v);
          pc.seek_text(d);
          for (uint64_t k = t0; k < t1; k++) {
            pc.next();
            if (k != t0) s.byte(' ');
            s.itoa(pc.ps);
            s.byte(' ');
            s.itoa(pc.pe);
          }
        } else {
          for (uint64_t k = 2 * t0; k < 2 * t1; k++) {
            if (k != 2 * t0) s.byte(' ');
            s.itoa(v->tok_pos[k]);
          }
        }
        s.byte('\n');
      }
      if (spos) {
        const uint64_t p1 = v->text_sentpos_end[d];
        for (uint64_t k = sp; k < p1; k++) {
          if (k != sp) s.byte(' ');
          s.itoa(v->sent_pos[k]);
        }
        s.byte('\n');
        sp = p1;
      }
    } else {
      s.byte('\n');  // token_writer.go:163-166
    }
  }
  if (tail) emit_sentences(v->n_sentences);  // SentenceEnd events after the last TextEnd
}

// Size of format_range()'s output without touching the text: token lengths, digit counts and
// separators only (valid UTF-8: surfaces are copied verbatim).
inline size_t digits10(int32_t v) {
  uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
  return (v < 0) + 1 + (u >= 10) + (u >= 100) + (u >= 1000) + (u >= 10000) + (u >= 100000) + (u >= 1000000) +
         (u >= 10000000) + (u >= 100000000) + (u >= 1000000000);
}
size_t count_range(const datok_view* v, uint32_t flags, uint64_t d0, uint64_t d1, bool tail) {
  const bool tokens = flags & DATOK_TOKENS, sentences = flags & DATOK_SENTENCES;
  const bool tpos = flags & DATOK_TOKEN_POS, spos = flags & DATOK_SENTENCE_POS;
  const uint64_t t0 = d0 ? v->text_tok_end[d0 - 1] : 0, t1 = d1 ? v->text_tok_end[d1 - 1] : 0;
  const uint64_t s0 = d0 ? v->text_sent_end[d0 - 1] : 0, s1 = d1 ? v->text_sent_end[d1 - 1] : 0;
  const uint64_t p0 = d0 ? v->text_sentpos_end[d0 - 1] : 0, p1 = d1 ? v->text_sentpos_end[d1 - 1] : 0;
  size_t n = 0;
  // tokens of the range's texts (+ those emitted by trailing SentenceEnd events)
  const uint64_t t_hi = (tail && sentences && v->n_sentences > s1) ? v->sent_tok[v->n_sentences - 1] : t1;
  if (tokens || tpos) {
    TokenCursor cur(v);
    cur.seek_text(d0);
    for (uint64_t k = t0; k < (t_hi > t1 ? t_hi : t1); k++) {
      cur.next();
      if (tokens) n += (size_t)(cur.hi - cur.lo) + 1;
      if (tpos && k < t1) n += digits10(cur.ps) + digits10(cur.pe) + 2;  // each number is followed by ' ' or '\n'
    }
  }
  if (sentences) n += (tail ? v->n_sentences : s1) - s0;
  if (tpos || spos) {
    for (uint64_t d = d0; d < d1; d++) {
      if (tpos && v->text_tok_end[d] == (d ? v->text_tok_end[d - 1] : 0)) n += 1;  // empty list: just the newline
      if (spos) {
        const uint64_t a = d ? v->text_sentpos_end[d - 1] : 0, b = v->text_sentpos_end[d];
        if (a == b) n += 1;
      }
    }
    if (spos) for (uint64_t k = p0; k < p1; k++) n += digits10(v->sent_pos[k]) + 1;
  } else {
    n += d1 - d0;
  }
  return n;
}

}  // namespace

extern "C" {

size_t datok_format(const datok_result* r, const uint8_t* in, size_t n, uint32_t flags, uint8_t* dst, size_t cap) {
  const datok_view* v = datok_result_view(r);
  if (!v) return (size_t)-1;
  const bool tokens = flags & DATOK_TOKENS, sentences = flags & DATOK_SENTENCES;
  const bool tpos = flags & DATOK_TOKEN_POS, spos = flags & DATOK_SENTENCE_POS;
  const bool have_delta = v->tok_delta || v->tok_delta8;
  if ((tokens && !v->tok_bytes && !have_delta) || (sentences && !v->sent_tok && v->n_sentences) ||
      (tpos && !v->tok_pos && !have_delta) || (spos && !v->sent_pos))
    return (size_t)-1;  // the array was not requested at transduce time
  // ---- ranges of texts, one per worker thread ----
  unsigned workers = std::thread::hardware_concurrency();
  if (const char* e = std::getenv("DATOK_FORMAT_THREADS")) workers = (unsigned)std::atoi(e);
  if (workers > 64) workers = 64;
  if (v->n_texts < 64 || workers < 2) workers = 1;
  if (workers == 1) {
    Sink s{dst, dst ? cap : 0};
    format_range(v, in, n, flags, 0, v->n_texts, true, s);
    return s.n;
  }
  std::vector<uint64_t> lo(workers + 1);
  for (unsigned i = 0; i <= workers; i++) lo[i] = v->n_texts * i / workers;
  // pass 1: sizes
  std::vector<size_t> size(workers, 0);
  {
    std::vector<std::thread> th;
    for (unsigned i = 0; i < workers; i++)
      th.emplace_back([&, i] {
        if (!v->has_invalid_utf8) { size[i] = count_range(v, flags, lo[i], lo[i + 1], i + 1 == workers); return; }
        Sink s{nullptr, 0};
        format_range(v, in, n, flags, lo[i], lo[i + 1], i + 1 == workers, s);
        size[i] = s.n;
      });
    for (auto& t : th) t.join();
  }
  size_t total = 0;
  std::vector<size_t> off(workers);
  for (unsigned i = 0; i < workers; i++) { off[i] = total; total += size[i]; }
  if (!dst) return total;
  if (cap < total) {  // truncated output: the plain sequential writer handles the cut
    Sink s{dst, cap};
    format_range(v, in, n, flags, 0, v->n_texts, true, s);
    return s.n;
  }
  // pass 2: every range writes at its offset
  std::vector<size_t> wrote(workers, 0);
  {
    std::vector<std::thread> th;
    for (unsigned i = 0; i < workers; i++)
      th.emplace_back([&, i] {
        Sink s{dst + off[i], size[i]};
        format_range(v, in, n, flags, lo[i], lo[i + 1], i + 1 == workers, s);
        wrote[i] = s.n;
      });
    for (auto& t : th) t.join();
  }
  for (unsigned i = 0; i < workers; i++)
    if (wrote[i] != size[i]) {  // the two passes disagree: never hand out a torn buffer
      Sink s{dst, cap};
      format_range(v, in, n, flags, 0, v->n_texts, true, s);
      return s.n;
    }
  return total;
}

int datok_replay(const datok_result* r, const uint8_t* in, size_t n, const datok_callbacks* cb) {
  const datok_view* v = datok_result_view(r);
  if (!v || !cb) return DATOK_ERR_INVALID_ARG;
  if ((v->n_tokens && !v->tok_bytes && !v->tok_delta && !v->tok_delta8) || (v->n_sentences && !v->sent_tok)) return DATOK_ERR_INVALID_ARG;
  (void)n;
  uint64_t tok = 0, sen = 0;
  size_t bufstart = 0;  // the reference's buffer[0]: the last rewind point (matrix.go:608-622)
  TokenCursor cur(v);
  auto emit_tokens = [&](uint64_t upto) {
    for (; tok < upto; tok++) {
      cur.next();
      const size_t lo = cur.lo, hi = cur.hi;
      int32_t runes = 0;
      for (size_t p = bufstart; p < lo;) { int w; datok::decode_rune(in + p, lo - p, &w); p += (size_t)w; runes++; }
      if (cb->token) cb->token(cb->user, in + bufstart, hi - bufstart, lo - bufstart, runes);
      bufstart = hi;
    }
  };
  auto emit_sentences = [&](uint64_t upto) {
    for (; sen < upto; sen++) {
      emit_tokens(v->sent_tok[sen]);
      if (cb->sentence_end) cb->sentence_end(cb->user);
    }
  };
  for (uint64_t d = 0; d < v->n_texts; d++) {
    emit_sentences(v->text_sent_end[d]);
    emit_tokens(v->text_tok_end[d]);
    if (cb->text_end) cb->text_end(cb->user);
    bufstart = v->text_byte_end[d];
  }
  emit_sentences(v->n_sentences);
  return DATOK_OK;
}

}  // extern "C"
