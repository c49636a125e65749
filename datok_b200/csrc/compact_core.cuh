// compact_core.cuh -- boundary bits -> offset arrays (K3), per-word bodies.
//
// Replaces the *data* half of the reference's TokenWriter (token_writer.go:36-175):
// the closures' captured state (posC, pos, sent, sentB, init) and the walk's
// sentenceEnd/textEnd flags (matrix.go:360-363,595-605,683-695) become an
// associative summary (Agg) that is scanned over 32-position words of the
// boundary bitmaps written by walk_run() (walk_core.cuh).
//
// Stream order of events: by byte position, and at one position
//     END (token ends here)  <  SENT (SentenceEnd)  <  TEND (EOT at this byte fired TextEnd).
#pragma once
#include "walk_core.cuh"

namespace datok {

// event kinds (last_kind / first_kind)
constexpr uint32_t EV_NONE = 0;    // empty span
constexpr uint32_t EV_END = 1;     // Token              -> sentenceEnd=false, textEnd=false (matrix.go:571-572)
constexpr uint32_t EV_SENT = 2;    // SentenceEnd        -> sentenceEnd=true, sentB=true
constexpr uint32_t EV_TEND = 3;    // TextEnd            -> sentenceEnd=true, textEnd=true, sentB=true
constexpr uint32_t EV_START0 = 4;  // stream start with sentenceEnd=false (matrix.go:360), sentB=true

// TokenWriter Bits (token_writer.go:17-25)
constexpr uint32_t F_TOKENS = 1, F_SENTENCES = 2, F_TOKEN_POS = 4, F_SENTENCE_POS = 8, F_NL_AFTER_EOT = 16,
                   F_WRITER_USED = 256;

// Summary of a contiguous span of positions.  Ranks are rune counts relative to
// the span start; positions are absolute bytes.
struct Agg {
  uint32_t n_rune;     // rune starts in the span
  uint32_t n_tok;      // Token events
  uint32_t n_sent;     // SentenceEnd events (regular + forced ones already decidable)
  uint32_t n_text;     // TextEnd events
  uint32_t n_sentpos;  // entries appended to TokenWriter.sent (token_writer.go:76-79,108)
  uint32_t first_kind; // kind of the first event (its forced/opener decision needs the predecessor)
  uint32_t last_kind;
  uint32_t last_end_pos;   // byte position of the last END (K_NOPOS = none)
  uint32_t last_end_rank;  // runes from span start to that position
  uint32_t doc_start;      // byte after the last TEND (K_NOPOS = no TEND in span)
  uint32_t doc_rank;       // runes from span start to doc_start
  uint32_t doc_adj;        // NEWLINE_AFTER_EOT shift of the text starting at doc_start
  uint32_t doc_tok;        // tokens in the span before doc_start
};
constexpr int AGG_WORDS = 13;

DATOK_HD Agg agg_zero() {
  Agg a;
  a.n_rune = a.n_tok = a.n_sent = a.n_text = a.n_sentpos = 0;
  a.first_kind = a.last_kind = EV_NONE;
  a.last_end_pos = K_NOPOS; a.last_end_rank = 0;
  a.doc_start = K_NOPOS; a.doc_rank = 0; a.doc_adj = 0; a.doc_tok = 0;
  return a;
}

// does a TextEnd following an event of kind `k` force a SentenceEnd? (matrix.go:595-598)
DATOK_HD bool forces_sentence(uint32_t k) { return k == EV_END || k == EV_START0; }
// is a Token following an event of kind `k` the first of a sentence (sentB)? (token_writer.go:76)
DATOK_HD bool opens_sentence(uint32_t k) { return k == EV_SENT || k == EV_TEND || k == EV_START0; }

// R = A followed by B
DATOK_HD Agg agg_combine(const Agg& A, const Agg& B) {
  Agg R;
  uint32_t extra_sent = 0, extra_open = 0;
  if (A.last_kind != EV_NONE) {
    if (B.first_kind == EV_TEND && forces_sentence(A.last_kind)) extra_sent = 1;
    if (B.first_kind == EV_END && opens_sentence(A.last_kind)) extra_open = 1;
  }
  R.n_rune = A.n_rune + B.n_rune;
  R.n_tok = A.n_tok + B.n_tok;
  R.n_sent = A.n_sent + B.n_sent + extra_sent;
  R.n_text = A.n_text + B.n_text;
  R.n_sentpos = A.n_sentpos + B.n_sentpos + extra_sent + extra_open;
  R.first_kind = A.first_kind != EV_NONE ? A.first_kind : B.first_kind;
  R.last_kind = B.last_kind != EV_NONE ? B.last_kind : A.last_kind;
  if (B.last_end_pos != K_NOPOS) { R.last_end_pos = B.last_end_pos; R.last_end_rank = A.n_rune + B.last_end_rank; }
  else { R.last_end_pos = A.last_end_pos; R.last_end_rank = A.last_end_rank; }
  if (B.doc_start != K_NOPOS) {
    R.doc_start = B.doc_start; R.doc_rank = A.n_rune + B.doc_rank; R.doc_adj = B.doc_adj; R.doc_tok = A.n_tok + B.doc_tok;
  } else {
    R.doc_start = A.doc_start; R.doc_rank = A.doc_rank; R.doc_adj = A.doc_adj; R.doc_tok = A.doc_tok;
  }
  return R;
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
  // outputs (device); any may be null
  uint32_t* tok_bytes;     // 2 per token
  int32_t* tok_pos;        // 2 per token
  int32_t* sent_pos;
  uint32_t* sent_tok;
  uint32_t* text_tok_end;
  uint32_t* text_sent_end;
  uint32_t* text_sentpos_end;
  uint32_t* text_byte_end;
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

// Walks the events of bitmap word `w` in stream order.
//   EMIT == false: returns the word's Agg (relative ranks), `carry` unused.
//   EMIT == true : `carry` is the absolute summary of everything before the word
//                  (including the stream-start pseudo event); writes the outputs and
//                  returns the absolute summary including the word.
template <bool EMIT>
DATOK_HD Agg process_word(const CompactCtx& c, uint32_t w, const Agg& carry) {
  const uint32_t rs = c.rstart[w], we = c.b_end[w], ws = c.b_sent[w], wt = c.b_tend[w];
  Agg a = agg_zero();
  a.n_rune = popc32(rs);
  uint32_t m = we | ws | wt;
  if (m == 0) {
    if (EMIT) { a = carry; a.n_rune = carry.n_rune + popc32(rs); }
    return a;
  }
  // running absolute state (EMIT) / relative state (!EMIT)
  uint32_t lk = EMIT ? carry.last_kind : EV_NONE;
  uint32_t tok = EMIT ? carry.n_tok : 0, sent = EMIT ? carry.n_sent : 0, text = EMIT ? carry.n_text : 0;
  uint32_t sentpos = EMIT ? carry.n_sentpos : 0;
  const uint32_t rank0 = EMIT ? carry.n_rune : 0;  // rank of the word's first position
  uint32_t last_end_pos = EMIT ? carry.last_end_pos : K_NOPOS, last_end_rank = EMIT ? carry.last_end_rank : 0;
  uint32_t doc_start = EMIT ? carry.doc_start : K_NOPOS, doc_rank = EMIT ? carry.doc_rank : 0;
  uint32_t doc_adj = EMIT ? carry.doc_adj : 0, doc_tok = EMIT ? carry.doc_tok : 0;
  const bool writer_used = (c.flags & F_WRITER_USED) != 0;

  while (m) {
    const uint32_t b = ctz32(m);
    m &= m - 1;
    const uint32_t p = (w << 5) + b;
    if ((we >> b) & 1u) {  // ---- Token (token_writer.go:59-88) ----
      const uint32_t rank_e = rank0 + popc32(rs & mask_below(b));
      if (a.first_kind == EV_NONE) a.first_kind = EV_END;
      const bool opener = (lk != EV_NONE) && opens_sentence(lk);
      if (EMIT) {
        const uint32_t D = doc_start == K_NOPOS ? 0u : doc_start;
        uint32_t bufstart = D;
        if (last_end_pos != K_NOPOS && last_end_pos > bufstart) bufstart = last_end_pos;
        const uint32_t s = next_clear(c.b_skip, c.n_words, bufstart);  // offset = leading non-token runes
        const uint32_t rank_s = rank_e - count_range(c.rstart, s, p);
        // `init` (token_writer.go:42,66,70): the text holding the stream's first token is never shifted
        const int32_t adj = (!writer_used && doc_tok == 0) ? 0 : (int32_t)doc_adj;
        const int32_t ps = (int32_t)(rank_s - doc_rank) - adj, pe = (int32_t)(rank_e - doc_rank) - adj;
        if (c.tok_bytes) { c.tok_bytes[2 * (size_t)tok] = s; c.tok_bytes[2 * (size_t)tok + 1] = p; }
        if (c.tok_pos) { c.tok_pos[2 * (size_t)tok] = ps; c.tok_pos[2 * (size_t)tok + 1] = pe; }
        if (opener && c.sent_pos) c.sent_pos[sentpos] = ps;
      }
      if (opener) sentpos++;
      tok++;
      lk = EV_END;
      last_end_pos = p;
      last_end_rank = rank_e;
    }
    if ((ws >> b) & 1u) {  // ---- SentenceEnd (token_writer.go:103-127) ----
      if (a.first_kind == EV_NONE) a.first_kind = EV_SENT;
      if (EMIT) {
        const int32_t adj = (!writer_used && doc_tok == 0) ? 0 : (int32_t)doc_adj;
        if (c.sent_tok) c.sent_tok[sent] = tok;
        if (tok == doc_tok) { if (c.flags & F_SENTENCE_POS) report_error(c, p, E_SENT_NO_TOKEN); }
        else if (c.sent_pos) c.sent_pos[sentpos] = (int32_t)(last_end_rank - doc_rank) - adj;
      }
      sent++;
      sentpos++;
      lk = EV_SENT;
    }
    if ((wt >> b) & 1u) {  // ---- EOT: forced SentenceEnd + TextEnd (matrix.go:593-605) ----
      if (a.first_kind == EV_NONE) a.first_kind = EV_TEND;
      if (lk != EV_NONE && forces_sentence(lk)) {
        if (EMIT) {
          const int32_t adj = (!writer_used && doc_tok == 0) ? 0 : (int32_t)doc_adj;
          if (c.sent_tok) c.sent_tok[sent] = tok;
          if (tok == doc_tok) { if (c.flags & F_SENTENCE_POS) report_error(c, p, E_SENT_NO_TOKEN); }
          else if (c.sent_pos) c.sent_pos[sentpos] = (int32_t)(last_end_rank - doc_rank) - adj;
        }
        sent++;
        sentpos++;
      }
      if (EMIT) {
        if (tok == doc_tok) {  // token-less text (token_writer.go:135,145)
          if (c.flags & F_TOKEN_POS) report_error(c, p, E_TEXT_NO_TOKEN);
          else if (c.flags & F_SENTENCE_POS) report_error(c, p, E_TEXT_NO_SENT);
        }
        c.text_tok_end[text] = tok;
        c.text_sent_end[text] = sent;
        c.text_sentpos_end[text] = sentpos;
        c.text_byte_end[text] = p + 1;
      }
      text++;
      lk = EV_TEND;
      doc_start = p + 1;
      doc_rank = rank0 + popc32(rs & mask_below(b + 1));
      doc_adj = newline_adjust(c, p + 1);
      doc_tok = tok;
    }
  }
  if (EMIT) { a.n_rune = rank0 + popc32(rs); a.first_kind = carry.first_kind; }
  a.n_tok = tok; a.n_sent = sent; a.n_text = text; a.n_sentpos = sentpos;
  a.last_kind = lk;
  a.last_end_pos = last_end_pos; a.last_end_rank = last_end_rank;
  a.doc_start = doc_start; a.doc_rank = doc_rank; a.doc_adj = doc_adj; a.doc_tok = doc_tok;
  return a;
}

// The pseudo span that precedes position 0: the walk's initial sentenceEnd flag
// and, for a reused TokenWriter, the shift of the first text.
DATOK_HD Agg agg_stream_start(const CompactCtx& c, bool sentence_end) {
  Agg a = agg_zero();
  a.first_kind = a.last_kind = sentence_end ? EV_SENT : EV_START0;
  if (c.flags & F_WRITER_USED) { a.doc_start = 0; a.doc_adj = newline_adjust(c, 0); }
  return a;
}

// End of input (matrix.go:680-695): the final SentenceEnd / TextEnd.  `tot` is the
// absolute summary of the whole stream.  Returns the final counts in `tot`.
DATOK_HD void finalize_stream(const CompactCtx& c, Agg& tot, bool text_end_in) {
  const bool writer_used = (c.flags & F_WRITER_USED) != 0;
  const int32_t adj = (!writer_used && tot.doc_tok == 0) ? 0 : (int32_t)tot.doc_adj;
  const bool have_tok = tot.n_tok != tot.doc_tok;
  if (forces_sentence(tot.last_kind)) {  // :683 if !sentenceEnd
    if (c.sent_tok) c.sent_tok[tot.n_sent] = tot.n_tok;
    if (!have_tok) { if (c.flags & F_SENTENCE_POS) report_error(c, c.N, E_SENT_NO_TOKEN); }
    else if (c.sent_pos) c.sent_pos[tot.n_sentpos] = (int32_t)(tot.last_end_rank - tot.doc_rank) - adj;
    tot.n_sent++;
    tot.n_sentpos++;
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
    c.text_tok_end[tot.n_text] = tot.n_tok;
    c.text_sent_end[tot.n_text] = tot.n_sent;
    c.text_sentpos_end[tot.n_text] = tot.n_sentpos;
    c.text_byte_end[tot.n_text] = c.N;
    tot.n_text++;
  }
}

}  // namespace datok
