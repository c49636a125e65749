// compact_core.cuh -- boundary bits -> offset arrays (K3), per-word bodies.
//
// Replaces the *data* half of the reference's TokenWriter (token_writer.go:36-175):
// the closures' captured state (posC, pos, sent, sentB, init) and the walk's
// sentenceEnd/textEnd flags (matrix.go:360-363,595-605,683-695) become an
// associative summary (Agg) that is scanned over 32-position words of the
// boundary bitmaps written by the walk (walk_core.cuh, fast_core.cuh).
//
// Stream order of events: by byte position, and at one position
//     END (token ends here)  <  SENT (SentenceEnd)  <  TEND (EOT at this byte fired TextEnd).
//
// Everything about a word is computed with word-parallel logic: which TextEnds force a
// SentenceEnd (matrix.go:595-598) and which tokens open a sentence (token_writer.go:76) depend on
// the kind of the preceding event, a fill-forward over the word done with one addition
// (word_masks); the output index of every event is then a popcount.  Per-text state (posC reset,
// NEWLINE_AFTER_EOT shift, `init`) lives in a table of DocRec written by a pass over the TextEnds
// (emit_texts) before tokens and sentences are emitted (emit_tokens, emit_sentences).
#pragma once
#include "walk_core.cuh"

namespace datok {

#define DATOK_RARE(x) __builtin_expect(!!(x), 0)

// event kinds (last_kind / first_kind)
constexpr uint32_t EV_NONE = 0;    // empty span
constexpr uint32_t EV_END = 1;     // Token              -> sentenceEnd=false, textEnd=false (matrix.go:571-572)
constexpr uint32_t EV_SENT = 2;    // SentenceEnd        -> sentenceEnd=true, sentB=true
constexpr uint32_t EV_TEND = 3;    // TextEnd            -> sentenceEnd=true, textEnd=true, sentB=true
constexpr uint32_t EV_START0 = 4;  // stream start with sentenceEnd=false (matrix.go:360), sentB=true

// TokenWriter Bits (token_writer.go:17-25)
constexpr uint32_t F_TOKENS = 1, F_SENTENCES = 2, F_TOKEN_POS = 4, F_SENTENCE_POS = 8, F_NL_AFTER_EOT = 16,
                   F_WRITER_USED = 256;

// Summary of a contiguous span of positions.
struct Agg {
  uint32_t n_rune;     // rune starts in the span
  uint32_t n_tok;      // Token events
  uint32_t n_sent;     // SentenceEnd events (regular + forced ones already decidable)
  uint32_t n_text;     // TextEnd events
  uint32_t n_sentpos;  // entries appended to TokenWriter.sent (token_writer.go:76-79,108)
  uint32_t kinds;      // kind of the first event | kind of the last event << 8 (EV_NONE: no event)
  uint32_t last_end_pos;   // byte position of the last END (K_NOPOS = none)
  uint32_t doc_start;      // byte after the last TEND (K_NOPOS = no TEND in span)
};
constexpr int AGG_WORDS = 8;

// State of the TokenWriter at the start of a text (token_writer.go:130-167 resets posC, pos, sent).
struct DocRec {
  uint32_t start;      // first byte of the text
  uint32_t rank;       // runes before `start`
  uint32_t adj;        // NEWLINE_AFTER_EOT shift of the text (token_writer.go:66-68)
  uint32_t tok;        // tokens before `start`
};

DATOK_HD Agg agg_zero() {
  Agg a;
  a.n_rune = a.n_tok = a.n_sent = a.n_text = a.n_sentpos = 0;
  a.kinds = 0;
  a.last_end_pos = K_NOPOS;
  a.doc_start = K_NOPOS;
  return a;
}
DATOK_HD uint32_t agg_first(const Agg& a) { return a.kinds & 0xFFu; }
DATOK_HD uint32_t agg_last(const Agg& a) { return a.kinds >> 8; }

// does a TextEnd following an event of kind `k` force a SentenceEnd? (matrix.go:595-598)
DATOK_HD bool forces_sentence(uint32_t k) { return k == EV_END || k == EV_START0; }
// is a Token following an event of kind `k` the first of a sentence (sentB)? (token_writer.go:76)
DATOK_HD bool opens_sentence(uint32_t k) { return k == EV_SENT || k == EV_TEND || k == EV_START0; }

// R = A followed by B
DATOK_HD Agg agg_combine(const Agg& A, const Agg& B) {
  Agg R;
  const uint32_t al = agg_last(A), bf = agg_first(B);
  uint32_t extra_sent = 0, extra_open = 0;
  if (al != EV_NONE) {
    if (bf == EV_TEND && forces_sentence(al)) extra_sent = 1;
    if (bf == EV_END && opens_sentence(al)) extra_open = 1;
  }
  R.n_rune = A.n_rune + B.n_rune;
  R.n_tok = A.n_tok + B.n_tok;
  R.n_sent = A.n_sent + B.n_sent + extra_sent;
  R.n_text = A.n_text + B.n_text;
  R.n_sentpos = A.n_sentpos + B.n_sentpos + extra_sent + extra_open;
  R.kinds = (agg_first(A) != EV_NONE ? agg_first(A) : bf) | ((agg_last(B) != EV_NONE ? agg_last(B) : al) << 8);
  R.last_end_pos = B.last_end_pos != K_NOPOS ? B.last_end_pos : A.last_end_pos;
  R.doc_start = B.doc_start != K_NOPOS ? B.doc_start : A.doc_start;
  return R;
}

// What the reduce pass leaves per warp unit (32 * COMPACT_WPT words) for the texts and emit passes: the summary of the
// units BEFORE it in its block, with WAGG_HAS_TEXT set in `kinds` (whose kinds use 16 bits) when the unit's own words
// hold a TextEnd.
constexpr uint32_t WAGG_HAS_TEXT = 1u << 31;
DATOK_HD Agg load_warp_prefix(const Agg* warp_agg, size_t unit) {
  Agg a = warp_agg[unit];
  a.kinds &= ~WAGG_HAS_TEXT;
  return a;
}

struct CompactCtx {
  const uint8_t* in;       // input bytes (for the NEWLINE_AFTER_EOT test, token_writer.go:66)
  uint32_t N;
  uint32_t n_words;        // bitmap words
  const uint32_t* rstart;
  const uint32_t* b_end;
  const uint32_t* b_skip;
  const uint32_t* b_sent;
  const uint32_t* b_tend;
  uint32_t flags;          // TokenWriter Bits (+ F_WRITER_USED)
  uint32_t eot_rewind;     // 0: double-array walk -- the buffer is not rewound at an EOT (datok.go:1019-1030): the first
                           // Token call of a text carries the runes since the last token of the text before
  // this input is a piece of a longer stream: what the pieces before it produced (added to every
  // index / byte offset written to the outputs; DocRec and Agg stay piece-relative)
  uint32_t base_tok, base_sent, base_sentpos, base_byte;
  // outputs (device); any of the first four may be null
  uint32_t* tok_bytes;     // 2 per token
  int32_t* tok_pos;        // 2 per token
  uint16_t* tok_delta;     // DATOK_COMPACT: 4 per token, instead of the two above
  uint8_t* tok_delta8;     // DATOK_COMPACT8: 4 bytes per token; values >= 255 go to the escape list
  uint32_t* esc;           // escape list: pairs {token index, field << 16 | value}, unordered on the device
  uint32_t* esc_count;
  uint32_t esc_cap;        // pairs
  int32_t* sent_pos;
  uint32_t* sent_tok;
  uint32_t* text_tok_end;
  uint32_t* text_sent_end;
  uint32_t* text_sentpos_end;
  uint32_t* text_byte_end;
  DocRec* docs;            // docs[d]: the text that TextEnd d closes; docs[0] = stream start
  unsigned long long* err_key;  // min over (position << 8 | code), ~0 = none
};

// first position >= p whose bit is set in `w` (bitmap of n_words words), or K_NOPOS
DATOK_HD uint32_t next_set(const uint32_t* w, uint32_t n_words, uint32_t p) {
  uint32_t i = p >> 5;
  if (i >= n_words) return K_NOPOS;
  uint32_t m = w[i] & mask_from(p & 31);
  while (m == 0) {
    if (++i >= n_words) return K_NOPOS;
    m = w[i];
  }
  return (i << 5) + ctz32(m);
}
// first position >= p whose bit is CLEAR
DATOK_HD uint32_t next_clear(const uint32_t* w, uint32_t n_words, uint32_t p) {
  uint32_t i = p >> 5;
  if (i >= n_words) return p;
  uint32_t m = ~w[i] & mask_from(p & 31);
  while (m == 0) {
    if (++i >= n_words) return i << 5;
    m = ~w[i];
  }
  return (i << 5) + ctz32(m);
}

// NEWLINE_AFTER_EOT shift of the text whose first buffer starts at byte D
// (token_writer.go:66-68): posC-- whenever a Token call arrives with posC == 0 and
// buf[0] == '\n'.  posC is 0 at the first Token of a text, and again after a
// chunk that was exactly "\n".  (The `init` exemption is applied by the caller.)
// double-array walk: the text's first buffer starts at D, the end of the last token before the EOT at p (0: none
// yet); buf[0] is that rune.  One Token call at posC == 0 at most: its buffer spans the EOT, it is longer than "\n".
DATOK_HD uint32_t newline_adjust_norewind(const CompactCtx& c, uint32_t D, uint32_t p) {
  if (!(c.flags & F_NL_AFTER_EOT)) return 0;
  if (D >= c.N || c.in[D] != '\n') return 0;
  const uint32_t e = next_set(c.b_end, c.n_words, p + 1);   // the first token of the text behind the EOT ...
  if (e == K_NOPOS) return 0;
  const uint32_t q = next_set(c.b_tend, c.n_words, p + 1);  // ... if the text has one
  return (q != K_NOPOS && q < e) ? 0u : 1u;
}
DATOK_HD uint32_t newline_adjust(const CompactCtx& c, uint32_t D) {
  if (!(c.flags & F_NL_AFTER_EOT)) return 0;
  uint32_t adj = 0, b = D;
  for (;;) {
    if (b >= c.N || c.in[b] != '\n') break;
    uint32_t e = next_set(c.b_end, c.n_words, b + 1);   // end of the token whose buffer starts at b
    if (e == K_NOPOS) break;
    uint32_t q = next_set(c.b_tend, c.n_words, b);      // a TextEnd before that token ends the text
    if (q != K_NOPOS && q < e) break;
    adj++;
    if (e != b + 1) break;  // the buffer was longer than "\n": posC > 0 from here on
    b = e;
  }
  return adj;
}

DATOK_HD void report_error(const CompactCtx& c, uint32_t pos, uint32_t code) {
  unsigned long long key = ((unsigned long long)pos << 8) | code;
#if defined(__CUDA_ARCH__)
  atomicMin(c.err_key, key);
#else
  if (key < *c.err_key) *c.err_key = key;
#endif
}

// the five bitmap words of positions [32w, 32w+32)
struct WordBits {
  uint32_t rs, e, k, s, t;  // rune starts, END, SKIP, SENT, TEND
};
DATOK_HD WordBits word_load(const CompactCtx& c, uint32_t w) {
  WordBits b;
  b.rs = c.rstart[w]; b.e = c.b_end[w]; b.k = c.b_skip[w]; b.s = c.b_sent[w]; b.t = c.b_tend[w];
  return b;
}

// "was the last event before position p a Token?" for every p of the word, given the answer
// `cin` for the word's first position: positions whose last event is END generate, SENT/TEND kill,
// the others propagate -- the carries of one addition.
DATOK_HD uint32_t prev_is_end(const WordBits& b, uint32_t cin) {
  const uint32_t g = b.e & ~(b.s | b.t);
  const uint32_t x = g | ~(b.e | b.s | b.t);
  return (x + g + (cin & 1u)) ^ x ^ g;
}

struct WordMasks {
  uint32_t forced;  // TextEnds that force a SentenceEnd first (matrix.go:595-598)
  uint32_t opener;  // Tokens that open a sentence span (token_writer.go:76-79)
  uint32_t se;      // all SentenceEnd events: SENT | forced
};
// last_kind: kind of the last event before the word (EV_NONE only for a relative summary, where the
// decisions about the word's first event are left to agg_combine)
DATOK_HD WordMasks word_masks(const WordBits& b, uint32_t last_kind) {
  WordMasks m;
  const uint32_t pf = prev_is_end(b, forces_sentence(last_kind) ? 1u : 0u);
  // START0 both forces and opens; the two chains only differ up to the first event
  const uint32_t po = last_kind == EV_START0 ? prev_is_end(b, 0u) : pf;
  m.forced = b.t & ~b.s & (b.e | pf);
  m.opener = b.e & ~po;
  if (last_kind == EV_NONE) {
    const uint32_t any = b.e | b.s | b.t;
    if (any) m.opener &= ~(any & (0u - any));  // a first-event Token is decided by the predecessor
  }
  m.se = b.s | m.forced;
  return m;
}

// relative summary of word w
DATOK_HD Agg word_agg(uint32_t w, const WordBits& b) {
  Agg a = agg_zero();
  a.n_rune = popc32(b.rs);
  const uint32_t any = b.e | b.s | b.t;
  if (any == 0) return a;
  const WordMasks m = word_masks(b, EV_NONE);
  const uint32_t fb = any & (0u - any), lb = 0x80000000u >> clz32(any);
  const uint32_t first = (b.e & fb) ? EV_END : (b.s & fb) ? EV_SENT : EV_TEND;
  const uint32_t last = (b.t & lb) ? EV_TEND : (b.s & lb) ? EV_SENT : EV_END;
  a.n_tok = popc32(b.e);
  a.n_text = popc32(b.t);
  a.n_sent = popc32(m.se);
  a.n_sentpos = a.n_sent + popc32(m.opener);
  a.kinds = first | (last << 8);
  if (b.e) a.last_end_pos = (w << 5) + 31u - clz32(b.e);
  if (b.t) a.doc_start = (w << 5) + 32u - clz32(b.t);
  return a;
}

DATOK_HD int32_t doc_shift(const CompactCtx& c, const DocRec& d) {
  // `init` (token_writer.go:42,66,70): the text holding the stream's first token is never shifted
  return (!(c.flags & F_WRITER_USED) && d.tok == 0) ? 0 : (int32_t)d.adj;
}

// Texts pass: the TextEnd events of word w (token_writer.go:130-167).  A: absolute summary of
// everything before the word (including the stream-start pseudo event).
DATOK_HD void emit_texts(const CompactCtx& c, uint32_t w, const WordBits& b, const Agg& A) {
  if (b.t == 0) return;
  const WordMasks m = word_masks(b, agg_last(A));
  uint32_t t = b.t;
  uint32_t doc_start = A.doc_start == K_NOPOS ? 0u : A.doc_start;
  while (t) {
    const uint32_t bp = ctz32(t);
    t &= t - 1;
    const uint32_t le = mask_below(bp + 1), p = (w << 5) + bp;
    const uint32_t text = A.n_text + popc32(b.t & mask_below(bp));
    const uint32_t tok = A.n_tok + popc32(b.e & le);
    const uint32_t sent = A.n_sent + popc32(m.se & le);
    const uint32_t sentpos = A.n_sentpos + popc32(m.opener & le) + popc32(m.se & le);
    // token-less text (token_writer.go:135,145): no END since the text started
    uint32_t last_end = A.last_end_pos;
    if (b.e & le) last_end = (w << 5) + 31u - clz32(b.e & le);
    // (a token that ends exactly at doc_start holds the EOT before it: possible in a double-array walk only, where the
    // EOT can stay in the buffer -- it was handed out after that TextEnd and belongs to this text)
    if (last_end == K_NOPOS || last_end < doc_start || (last_end == doc_start && (c.eot_rewind || doc_start == 0))) {
      if (c.flags & F_TOKEN_POS) report_error(c, p, E_TEXT_NO_TOKEN);
      else if (c.flags & F_SENTENCE_POS) report_error(c, p, E_TEXT_NO_SENT);
    }
    c.text_tok_end[text] = c.base_tok + tok;
    c.text_sent_end[text] = c.base_sent + sent;
    c.text_sentpos_end[text] = c.base_sentpos + sentpos;
    c.text_byte_end[text] = c.base_byte + p + 1;
    DocRec d;
    if (c.eot_rewind) {
      d.start = p + 1;
      d.rank = A.n_rune + popc32(b.rs & le);
      d.adj = newline_adjust(c, p + 1);
    } else {
      // no rewind at the EOT: the next text's first buffer starts where the last token ended (stream start: 0)
      const uint32_t any_end = (b.e & le) ? (w << 5) + 31u - clz32(b.e & le) : A.last_end_pos;
      if (any_end == K_NOPOS) { d.start = 0; d.rank = 0; }
      else {
        d.start = any_end;
        d.rank = (b.e & le) ? A.n_rune + popc32(b.rs & mask_below(31u - clz32(b.e & le)))
                            : A.n_rune - count_range(c.rstart, any_end, w << 5);
      }
      d.adj = newline_adjust_norewind(c, d.start, p);
    }
    d.tok = tok;
    c.docs[text + 1] = d;
    doc_start = p + 1;
  }
}

// The SentenceEnd events of word w (token_writer.go:103-127), regular and forced.
DATOK_HD void emit_sentences(const CompactCtx& c, uint32_t w, const WordBits& b, const WordMasks& m, const Agg& A) {
  uint32_t se = m.se;
  while (se) {
    const uint32_t bp = ctz32(se);
    se &= se - 1;
    const uint32_t lt = mask_below(bp), le = mask_below(bp + 1), p = (w << 5) + bp;
    const uint32_t sent = A.n_sent + popc32(m.se & lt);
    const uint32_t tok = A.n_tok + popc32(b.e & le);
    const uint32_t sentpos = A.n_sentpos + popc32(m.opener & le) + popc32(m.se & lt);
    if (c.sent_tok) c.sent_tok[sent] = c.base_tok + tok;
    if (!(c.flags & F_SENTENCE_POS) && !c.sent_pos) continue;
    const DocRec d = c.docs[A.n_text + popc32(b.t & lt)];
    if (tok == d.tok) {  // no token in this text yet (token_writer.go:108)
      if (c.flags & F_SENTENCE_POS) report_error(c, p, E_SENT_NO_TOKEN);
      continue;
    }
    if (!c.sent_pos) continue;
    // end of the last token: TokenWriter.pos[len(pos)-1]
    uint32_t rank;
    if (b.e & le) rank = A.n_rune + popc32(b.rs & mask_below(31u - clz32(b.e & le)));
    else rank = A.n_rune - count_range(c.rstart, A.last_end_pos, w << 5);
    c.sent_pos[sentpos] = (int32_t)(rank - d.rank) - doc_shift(c, d);
  }
}

constexpr uint32_t E_COMPACT_RANGE = 24;  // DATOK_ERR_COMPACT_RANGE

// The Token events of word w (token_writer.go:59-95).  Token k goes to slot k - tok_base of
// tok_bytes/tok_pos (2 values each, ABS) or of tok_delta (4 values, !ABS; the kernel stages a block's
// tokens in shared memory); sentence openers go to c.sent_pos.
DATOK_HD void emit_escape(const CompactCtx& c, uint32_t tok, uint32_t field, uint32_t value, uint32_t pos) {
#if defined(__CUDA_ARCH__)
  const uint32_t slot = atomicAdd(c.esc_count, 1u);
#else
  const uint32_t slot = (*c.esc_count)++;
#endif
  // beyond the capacity the pair is only counted: the host then runs the pass again with a list of that size
  if (slot < c.esc_cap) { c.esc[2 * slot] = c.base_tok + tok; c.esc[2 * slot + 1] = (field << 16) | value; }
  (void)pos;
}

// FORM: 0 absolute pairs (tok_bytes / tok_pos), 1 four u16 deltas per token (tok_delta), 2 four u8 deltas
// per token (tok_delta is then a byte array) with escapes
template <int FORM>
DATOK_HD void emit_tokens(const CompactCtx& c, uint32_t w, const WordBits& b, const WordMasks& m, const Agg& A,
                          uint32_t* tok_bytes, int32_t* tok_pos, uint16_t* tok_delta, uint32_t tok_base) {
  constexpr bool ABS = FORM == 0;
  uint32_t e = b.e;
  if (e == 0) return;
  const uint32_t w0 = w << 5;
  uint32_t tok = A.n_tok - tok_base;
  uint32_t prev_end = A.last_end_pos;
  uint32_t doc_id = A.n_text;
  DocRec d = c.docs[doc_id];
  int32_t shift = doc_shift(c, d);
  while (e) {
    const uint32_t bp = ctz32(e);
    e &= e - 1;
    const uint32_t lt = mask_below(bp), p = w0 + bp;
    if (DATOK_RARE(b.t & lt)) {  // a TextEnd earlier in this word
      const uint32_t id = A.n_text + popc32(b.t & lt);
      if (id != doc_id) { doc_id = id; d = c.docs[id]; shift = doc_shift(c, d); }
    }
    // the Token call's buffer starts at the previous rewind point; offset = leading non-token runes
    uint32_t bufstart = d.start;
    bool first = true;  // first token of its text
    if (prev_end != K_NOPOS && prev_end > bufstart) { bufstart = prev_end; first = false; }
    uint32_t s, runes, skipped = 0;  // token start, runes in [s, p), runes in [bufstart, s)
    const uint32_t from = mask_from(bufstart >= w0 ? bufstart - w0 : 32u);
    const uint32_t cl = ~b.k & lt & from;
    if (bufstart >= w0 && cl) {  // within the word
      const uint32_t sb = ctz32(cl);
      s = w0 + sb;
      runes = popc32(b.rs & lt & mask_from(sb));
      if (!ABS) skipped = popc32(b.rs & from & mask_below(sb));
    } else {
      s = next_clear(c.b_skip, c.n_words, bufstart);
      runes = count_range(c.rstart, s, p);
      if (!ABS) skipped = count_range(c.rstart, bufstart, s);
    }
    if (ABS) {
      const uint32_t rank_e = A.n_rune + popc32(b.rs & lt);
      const int32_t pe = (int32_t)(rank_e - d.rank) - shift, ps = pe - (int32_t)runes;
      if (tok_bytes) { tok_bytes[2 * tok] = c.base_byte + s; tok_bytes[2 * tok + 1] = c.base_byte + p; }
      if (tok_pos) { tok_pos[2 * tok] = ps; tok_pos[2 * tok + 1] = pe; }
    } else {
      const int32_t rskip = (int32_t)skipped - (first ? shift : 0);
      const uint32_t bskip = s - bufstart, blen = p - s;
      if (DATOK_RARE((bskip | blen | runes | (uint32_t)rskip) > 0xFFFFu)) report_error(c, p, E_COMPACT_RANGE);
      if (FORM == 1) {
        uint32_t* o = reinterpret_cast<uint32_t*>(tok_delta) + 2 * tok;  // {skip bytes, bytes}, {skip runes, runes}
        o[0] = bskip | (blen << 16);
        o[1] = ((uint32_t)rskip & 0xFFFFu) | (runes << 16);
      } else {
        // (no array here: dynamic indexing would put it into local memory for every token)
        const uint32_t v0 = bskip, v1 = blen, v2 = (uint32_t)rskip & 0xFFFFu, v3 = runes;
        if (DATOK_RARE((v0 | v1 | v2 | v3) >= 255u)) {
          if (v0 >= 255u) emit_escape(c, tok + tok_base, 0, v0, p);
          if (v1 >= 255u) emit_escape(c, tok + tok_base, 1, v1, p);
          if (v2 >= 255u) emit_escape(c, tok + tok_base, 2, v2, p);
          if (v3 >= 255u) emit_escape(c, tok + tok_base, 3, v3, p);
        }
        uint32_t packed;
#if defined(__CUDA_ARCH__)
        {  // four values saturated to 255 and packed, two instructions
          uint32_t hi2;
          asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi2) : "r"(v3), "r"(v2), "r"(0u));
          asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(packed) : "r"(v1), "r"(v0), "r"(hi2));
        }
#else
        {
          const uint32_t b0 = v0 < 255u ? v0 : 255u, b1 = v1 < 255u ? v1 : 255u, b2 = v2 < 255u ? v2 : 255u,
                         b3 = v3 < 255u ? v3 : 255u;
          packed = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
#endif
        reinterpret_cast<uint32_t*>(tok_delta)[tok] = packed;
      }
    }
    if (DATOK_RARE((m.opener >> bp) & 1u)) {
      if (c.sent_pos) {
        const uint32_t rank_e = A.n_rune + popc32(b.rs & lt);
        const int32_t ps = (int32_t)(rank_e - d.rank) - shift - (int32_t)runes;
        c.sent_pos[A.n_sentpos + popc32(m.opener & lt) + popc32(m.se & lt)] = ps;
      }
    }
    tok++;
    prev_end = p;
  }
}

// The pseudo span that precedes position 0: the walk's initial sentenceEnd flag.
DATOK_HD Agg agg_stream_start(bool sentence_end) {
  Agg a = agg_zero();
  const uint32_t k = sentence_end ? EV_SENT : EV_START0;
  a.kinds = k | (k << 8);
  return a;
}
// docs[0]: the text the stream starts in; for a reused TokenWriter, its shift
DATOK_HD DocRec doc_stream_start(const CompactCtx& c) {
  DocRec d;
  d.start = 0; d.rank = 0; d.tok = 0;
  d.adj = (c.flags & F_WRITER_USED) ? newline_adjust(c, 0) : 0;
  return d;
}

// what the host needs to know about the whole stream
struct StreamTotals {
  uint32_t n_rune, n_tok, n_sent, n_text, n_sentpos;
  uint32_t last_kind;
  uint32_t tokless;    // no Token since the last TextEnd (or since the stream start)
  uint32_t reserved;
};

// End of input (matrix.go:680-695): the final SentenceEnd / TextEnd.  `tot` is the absolute
// summary of the whole stream.
DATOK_HD StreamTotals finalize_stream(const CompactCtx& c, const Agg& tot, bool text_end_in, bool final_input) {
  StreamTotals r;
  r.n_rune = tot.n_rune; r.n_tok = tot.n_tok; r.n_sent = tot.n_sent; r.n_text = tot.n_text; r.n_sentpos = tot.n_sentpos;
  r.last_kind = agg_last(tot);
  r.reserved = 0;
  const uint32_t doc_start = tot.doc_start == K_NOPOS ? 0u : tot.doc_start;
  const bool have_tok = tot.last_end_pos != K_NOPOS &&
                        (tot.last_end_pos > doc_start || (tot.last_end_pos == doc_start && !c.eot_rewind && doc_start != 0));
  r.tokless = have_tok ? 0u : 1u;
  if (!final_input) return r;
  const DocRec d = c.docs[tot.n_text];
  if (forces_sentence(r.last_kind)) {  // :683 if !sentenceEnd
    if (c.sent_tok) c.sent_tok[r.n_sent] = c.base_tok + r.n_tok;
    if (!have_tok) { if (c.flags & F_SENTENCE_POS) report_error(c, c.N, E_SENT_NO_TOKEN); }
    else if (c.sent_pos) {
      const uint32_t rank = tot.n_rune - count_range(c.rstart, tot.last_end_pos, c.n_words << 5);
      c.sent_pos[r.n_sentpos] = (int32_t)(rank - d.rank) - doc_shift(c, d);
    }
    r.n_sent++;
    r.n_sentpos++;
  }
  // textEnd (:363): true after a TextEnd with no Token since
  bool text_end;
  if (have_tok) text_end = false;
  else if (tot.n_text > 0) text_end = true;
  else text_end = text_end_in;
  if (!text_end) {  // :690
    if (!have_tok) {
      if (c.flags & F_TOKEN_POS) report_error(c, c.N, E_TEXT_NO_TOKEN);
      else if (c.flags & F_SENTENCE_POS) report_error(c, c.N, E_TEXT_NO_SENT);
    }
    c.text_tok_end[r.n_text] = c.base_tok + r.n_tok;
    c.text_sent_end[r.n_text] = c.base_sent + r.n_sent;
    c.text_sentpos_end[r.n_text] = c.base_sentpos + r.n_sentpos;
    c.text_byte_end[r.n_text] = c.base_byte + c.N;
    r.n_text++;
  }
  return r;
}

}  // namespace datok
