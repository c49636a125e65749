// walk_core.cuh -- per-thread bodies of the classify (K1) and walk (K2) kernels.
//
// These are the B200 re-design of the reference's per-rune loop
// (MatrixTokenizer.TransduceTokenWriter, matrix.go:348-698):
//   * rune reader + sigma map (matrix.go:388-435)   -> classify_pos()
//   * greedy walk, single epsilon backtrack, hard fail, EOT, EOF tail
//     (matrix.go:437-695)                            -> walk_run()
// The walk works on absolute BYTE positions of the input instead of the
// reference's 1024-rune sliding buffer, and emits boundary BITS instead of
// calling TokenWriter closures:
//     END[p]   a token ends (exclusive) at byte p          (w.Token, :528,569,675)
//     SKIP[p]  byte p belongs to a skipped leading non-token rune (bufft++, :584-588)
//     SENT[p]  a SentenceEnd event fired at byte p         (w.SentenceEnd, :575)
//     TEND[q]  the EOT rune at byte q fired TextEnd        (w.TextEnd, :600)
// Offsets, sentence spans and the forced SentenceEnd at EOT (:595-598) are
// derived from the bits by the compaction kernels (compact_core.cuh).
//
// Everything here is __host__ __device__ so that tests/emul can run the very same
// bodies sequentially on the CPU for unit tests; the product only launches them
// from kernels.cu.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DATOK_HD __host__ __device__ __forceinline__
#define DATOK_HD_SLOW inline __host__ __device__ __noinline__  // rare paths: keep them out of the hot loop's registers
#else
#define DATOK_HD inline
#define DATOK_HD_SLOW inline
#endif

namespace datok {

#ifndef DATOK_MODEL_CONSTS
#define DATOK_MODEL_CONSTS
constexpr uint32_t K_CLS_EPS = 0, K_CLS_CONT = 1, K_CLS_EOT = 2;
constexpr uint32_t K_NT = 0x8000u;
#endif
constexpr uint32_t K_WINDOW = 1024;  // matrix.go:365 buffer := make([]rune, 1024)
constexpr uint32_t K_NOPOS = 0xFFFFFFFFu;

// error codes (DATOK_ERR_* of include/datok_b200.h)
constexpr uint32_t E_OVERFLOW = 1, E_SENT_NO_TOKEN = 2, E_TEXT_NO_TOKEN = 3, E_TEXT_NO_SENT = 4, E_DEGENERATE = 5;

DATOK_HD uint32_t popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}
DATOK_HD uint32_t ctz32(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
  return (uint32_t)(__ffs((int)x) - 1);
#else
  return (uint32_t)__builtin_ctz(x);
#endif
}
DATOK_HD uint32_t clz32(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
  return (uint32_t)__clz((int)x);
#else
  return (uint32_t)__builtin_clz(x);
#endif
}
// mask of bits [0, hi)
DATOK_HD uint32_t mask_below(uint32_t hi) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(0u), "r"(hi));  // one instruction, widths >= 32 clamp to all ones
  return r;
#else
  return hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
#endif
}
// mask of bits [lo, 32)
DATOK_HD uint32_t mask_from(uint32_t lo) { return ~mask_below(lo); }

// number of set bits of bitmap `w` at positions [lo, hi)
DATOK_HD uint32_t count_range(const uint32_t* w, uint32_t lo, uint32_t hi) {
  if (hi <= lo) return 0;
  uint32_t a = lo >> 5, b = (hi - 1) >> 5;
  if (a == b) return popc32(w[a] & mask_from(lo & 31) & mask_below(((hi - 1) & 31) + 1));
  uint32_t n = popc32(w[a] & mask_from(lo & 31));
  for (uint32_t i = a + 1; i < b; i++) n += popc32(w[i]);
  return n + popc32(w[b] & mask_below(((hi - 1) & 31) + 1));
}
// clear bits [lo, hi) of a bitmap the caller owns exclusively
DATOK_HD void clear_range(uint32_t* w, uint32_t lo, uint32_t hi) {
  if (hi <= lo) return;
  uint32_t a = lo >> 5, b = (hi - 1) >> 5;
  uint32_t ma = mask_from(lo & 31), mb = mask_below(((hi - 1) & 31) + 1);
  if (a == b) { w[a] &= ~(ma & mb); return; }
  w[a] &= ~ma;
  for (uint32_t i = a + 1; i < b; i++) w[i] = 0;
  w[b] &= ~mb;
}
DATOK_HD void set_bit(uint32_t* w, uint32_t p) { w[p >> 5] |= 1u << (p & 31); }
DATOK_HD bool get_bit(const uint32_t* w, uint32_t p) { return (w[p >> 5] >> (p & 31)) & 1u; }

// ---------------------------------------------------------------- classify (K1)

struct ClsTables {
  const uint8_t* ascii_cls;   // [128] class of runes 0x00..0x7F (0x04 -> K_CLS_EOT)
  const uint8_t* latin1_cls;  // [128] class of runes 0x80..0xFF (sigmaASCII, matrix.go:421-425)
  const uint32_t* rune_key;   // sorted runes >= 0x100 of sigma (the Go map, matrix.go:427)
  const uint8_t* rune_cls;
  uint32_t n_rune;
  uint32_t identity_cls;      // runes not in sigma -> identity (matrix.go:430-434)
  const ClsTables* self;      // device: a copy of this struct in shared memory, for the out-of-line classify_pos
};

DATOK_HD uint32_t class_of_rune(const ClsTables& T, uint32_t r) {
  if (r < 0x80) return T.ascii_cls[r];
  if (r < 0x100) return T.latin1_cls[r - 0x80];
  uint32_t lo = 0, hi = T.n_rune;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    uint32_t k = T.rune_key[mid];
    if (k == r) return T.rune_cls[mid];
    if (k < r) lo = mid + 1; else hi = mid;
  }
  return T.identity_cls;
}

// Width (2..4) of the well-formed UTF-8 sequence starting with the lead byte at q,
// or 0 if Go's DecodeRune would yield (RuneError, 1) there.  *rune gets the code point.
DATOK_HD uint32_t utf8_seq(const uint8_t* in, uint32_t N, uint32_t q, uint32_t* rune) {
  uint32_t b0 = in[q];
  uint32_t need, lo = 0x80, hi = 0xBF, cp;
  if (b0 >= 0xC2 && b0 <= 0xDF) { need = 2; cp = b0 & 0x1F; }
  else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 3; cp = b0 & 0x0F; if (b0 == 0xE0) lo = 0xA0; if (b0 == 0xED) hi = 0x9F; }
  else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 4; cp = b0 & 0x07; if (b0 == 0xF0) lo = 0x90; if (b0 == 0xF4) hi = 0x8F; }
  else return 0;
  if (N - q < need) return 0;
  for (uint32_t i = 1; i < need; i++) {
    uint32_t b = in[q + i];
    if (b < lo || b > hi) return 0;
    cp = (cp << 6) | (b & 0x3F);
    lo = 0x80; hi = 0xBF;
  }
  *rune = cp;
  return need;
}

// Class of the byte at p and whether a rune starts there.  UTF-8 is
// self-synchronising and Go consumes exactly one byte per malformed byte, so rune
// starts are decidable from a 3-byte neighbourhood: every non-continuation byte
// starts a rune; a continuation byte starts one (as U+FFFD) unless a well-formed
// sequence beginning within the 3 bytes before it covers it.
DATOK_HD uint32_t classify_pos_inl(const uint8_t* in, uint32_t N, uint32_t p, const ClsTables& T, bool* is_start, bool* invalid);
#if defined(__CUDA_ARCH__) && defined(DATOK_OUTLINE_CLASSIFY)
// Measured alternative (-DDATOK_OUTLINE_CLASSIFY): the body once, out of line -- it is inlined in many rarely
// taken places (every walker instance, the look-ahead, the rare classes of the fast path: ~3700 of the walk
// kernel's ~10000 instructions).  The kernel shrinks by a third, but the calls cost more than the instruction
// cache gains: walk 3.67 -> 3.87 ms, stitch 0.30 -> 0.37 ms per GiB.  The tables come through a copy of the
// struct in shared memory (T.self): a reference to a kernel-local struct would force that struct into local memory.
// Result: class | is_start << 8 | invalid << 9.
static __device__ __noinline__ uint32_t classify_pos_call(const uint8_t* in, uint32_t N, uint32_t p, const ClsTables* T) {
  bool st, inv;
  const uint32_t cl = classify_pos_inl(in, N, p, *T, &st, &inv);
  return cl | (st ? 256u : 0u) | (inv ? 512u : 0u);
}
DATOK_HD uint32_t classify_pos(const uint8_t* in, uint32_t N, uint32_t p, const ClsTables& T, bool* is_start, bool* invalid) {
  const uint32_t r = classify_pos_call(in, N, p, T.self);
  *is_start = (r >> 8) & 1u;
  *invalid = (r >> 9) & 1u;
  return r & 0xFFu;
}
#else
DATOK_HD uint32_t classify_pos(const uint8_t* in, uint32_t N, uint32_t p, const ClsTables& T, bool* is_start, bool* invalid) {
  return classify_pos_inl(in, N, p, T, is_start, invalid);
}
#endif
DATOK_HD
uint32_t classify_pos_inl(const uint8_t* in, uint32_t N, uint32_t p, const ClsTables& T, bool* is_start, bool* invalid) {
  uint32_t b = in[p];
  *invalid = false;
  *is_start = true;
  if (b < 0x80) return T.ascii_cls[b];
  uint32_t rune = 0xFFFD;
  if ((b & 0xC0) == 0x80) {
    for (uint32_t k = 1; k <= 3 && k <= p; k++) {
      uint32_t c = in[p - k];
      if ((c & 0xC0) == 0x80) continue;  // another continuation byte: look further back
      if (c >= 0xC0) {
        uint32_t r2;
        if (utf8_seq(in, N, p - k, &r2) > k) { *is_start = false; return K_CLS_CONT; }
      }
      break;  // nearest non-continuation byte decides
    }
    *invalid = true;
    return class_of_rune(T, 0xFFFD);
  }
  if (utf8_seq(in, N, p, &rune) == 0) { *invalid = true; rune = 0xFFFD; }
  return class_of_rune(T, rune);
}

// ------------------------------------------------------------------- walk (K2)

// Machine state at the top of the reference's loop with newchar == true
// (matrix.go:384-386), in absolute byte positions.
// (32 bytes, 16-byte aligned: one DRAM sector per chunk record, moved with two 128-bit accesses)
struct alignas(16) WState {
  uint32_t pos;        // base + buffc
  uint32_t tstart;     // base + bufft
  uint32_t eps_pos;    // base + epsilonOffset
  uint32_t base;       // position of buffer[0]: the last rewind point (matrix.go:608-622)
  uint32_t hw;         // furthest byte read since `base` (buffer fill, for the 1024 check)
  uint16_t t;          // current state, GPU numbering
  uint16_t eps_state;  // epsilonState (0 = none)
  uint32_t flags;      // WS_* bits
  uint32_t reserved;
};
constexpr uint32_t WS_PEND = 1;     // a hard-fail token ends exactly at `pos`; its END bit is still to be set
constexpr uint32_t WS_DONE = 2;     // EOF tail finished (matrix.go:650-678)
constexpr uint32_t WS_INVALID = 4;  // no usable state (void chunk not yet walked, or walker stopped on error)
constexpr uint32_t WS_EPS_OVER_EOT = 8;  // double-array walk only: an EOT was consumed behind the pending epsilon point
                                         // (a backtrack to it would read the EOT, and fire TextEnd, a second time)
constexpr uint32_t WS_ERR_SHIFT = 8;  // error code that stopped the walker

DATOK_HD bool wstate_equal(const WState& a, const WState& b) {
  return a.pos == b.pos && a.tstart == b.tstart && a.eps_pos == b.eps_pos && a.base == b.base &&
         a.hw == b.hw && a.t == b.t && a.eps_state == b.eps_state && a.flags == b.flags;
}

struct WalkCtx {
  const uint16_t* table;  // exact table: table[t << row_shift | cls], bit 15 = non-token, 0 = no transition
  uint32_t row_shift;
  uint32_t start;         // GPU id of the reference's state 1
  const uint8_t* in;      // raw input bytes; classes are derived on demand (classify_pos)
  uint32_t N;             // input bytes
  ClsTables cls;
  uint32_t* b_end;
  uint32_t* b_skip;
  uint32_t* b_sent;
  uint32_t* b_tend;
  uint32_t* hist;         // optional: visits per state (calibration of the hot-row order)
  uint32_t* hist_cls;     // with hist: occurrences per class (calibration of the class order)
  uint32_t final_input;   // 0: the stream continues in a later call: no end-of-input processing (matrix.go:650-695)
  uint32_t eot_rewind;    // 1: matrix walk -- an EOT rewinds the buffer (matrix.go:601-603); 0: double-array walk -- it
                          // does not (datok.go:1019-1030): window, token start and epsilon point live on across the EOT
};

// class of the byte at pos (continuation bytes of a well-formed rune: K_CLS_CONT)
DATOK_HD uint32_t class_at(const uint8_t* in, uint32_t N, uint32_t pos, const ClsTables& T) {
  const uint32_t b = in[pos];
  if (b < 0x80) return T.ascii_cls[b];
  bool st, inv;
  return classify_pos(in, N, pos, T, &st, &inv);
}
DATOK_HD uint32_t cls_at(const WalkCtx& c, uint32_t pos) { return class_at(c.in, c.N, pos, c.cls); }

// True iff more than 1024 runes would have been buffered: runes in [base, hw].
// Only ever evaluated when the byte distance alone allows it, i.e. practically never.
DATOK_HD bool window_overflow(const WalkCtx& c, uint32_t base, uint32_t hw) {
  if (hw < base || hw - base < K_WINDOW) return false;  // bytes >= runes
  uint32_t runes = 0;
  for (uint32_t p = base; p <= hw && p < c.N; p++) {
    bool st, inv;
    classify_pos(c.in, c.N, p, c.cls, &st, &inv);
    runes += st ? 1u : 0u;
  }
  return runes > K_WINDOW;
}

// Info about the first buffer window of a speculative walk, whose true `base` is
// only known once the predecessor chunk has been stitched.
struct SpecInfo {
  uint32_t first_hw;    // hw when the first rewind happened (or at exit if none)
  uint32_t had_rewind;
};

// Runs the reference loop from `st`.
//
// PROBE == true: until the FINAL arrival at the loop top with pos >= stop, or until
// EOF processing is complete (WS_DONE).  Hand-off rule: when the walk first arrives
// at `stop` it may still hold a pending epsilon point before `stop`
// (matrix.go:448-449) to which a later failure would backtrack (matrix.go:487-497).
// The walk therefore continues in PROBE mode -- reading on, writing nothing at
// positions >= stop -- until either that point is dead (it was consumed, or the
// next state has its own epsilon transition, which replaces it), or a failure
// backtracks below `stop`, in which case normal walking resumes and `stop` will be
// reached again.  The state handed to the successor is the snapshot taken at the
// last arrival, with the dead epsilon point cleared, so
//   * every event at a position < stop is written by this walker only, and
//   * a successor never has to backtrack below its own start.
// The lookahead is bounded by the reference's own 1024-rune buffer.
//
// PROBE == false: plain segment of the same lane; returns at the first arrival at
// pos >= stop with the epsilon point kept.
//
// SPEC: the walk started from a guessed clean state, so the first window's
// overflow check is deferred to the stitch (SpecInfo).  STOP_REWIND: return at the
// loop top that follows the first rewind (used to hand a fresh lane to the fast path).
//
// Returns an error code (0 = none); on error the state is unusable.
// resume_at (PROBE == false only): additionally return at the first loop top with
// pos >= resume_at once at least one iteration has been executed -- the fast path
// uses this to take the lane back as soon as the rare case is dealt with.
// walk_run_inl: the body, inlined into its caller (the stitch kernel: context and state stay in registers
// instead of a stack frame per thread); walk_run: the same, as a call (the fused walk kernel, where the
// rare paths must stay out of the hot loop's registers).
template <bool SPEC, bool PROBE, bool STOP_REWIND>
DATOK_HD uint32_t walk_run_inl(const WalkCtx& c, WState& st, uint32_t stop, SpecInfo* spec,
                               uint32_t resume_at = 0xFFFFFFFFu) {
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
  g_exact_calls++;
#endif
  uint32_t pos = st.pos, tstart = st.tstart, eps_pos = st.eps_pos, base = st.base, hw = st.hw;
  uint32_t t = st.t, eps_state = st.eps_state, flags = st.flags;
  const uint32_t N = c.N;
  bool first_window = SPEC;
  bool probing = false;
  bool stepped = false;
  uint32_t err = 0;
  // snapshot taken at the (latest) arrival at `stop`
  uint32_t s_pos = 0, s_tstart = 0, s_base = 0, s_hw = 0, s_t = 0, s_flags = 0;
  bool s_first_window = false;

  if ((flags & WS_PEND) && pos < stop) { set_bit(c.b_end, pos); flags &= ~WS_PEND; }

// window close: the reference rewinds its buffer to the current position (matrix.go:608-622)
#define DATOK_REWIND(newbase)                                                              \
  do {                                                                                     \
    if (first_window) { spec->first_hw = hw; spec->had_rewind = 1; first_window = false; } \
    else if (window_overflow(c, base, hw)) { err = E_OVERFLOW; goto out; }                 \
    base = (newbase); hw = base; eps_state = 0;                                            \
  } while (0)

  for (;;) {
    if (pos >= stop) {
      if (!PROBE) break;
      if (!probing) {
        probing = true;
        s_pos = pos; s_tstart = tstart; s_base = base; s_t = t; s_flags = flags;
        s_hw = (pos > base && hw < pos - 1) ? pos - 1 : hw;  // everything below pos has been read
        s_first_window = first_window;
      }
      if (eps_state == 0 || eps_pos >= stop) break;  // final arrival
      if (pos < N && c.table[(t << c.row_shift) | K_CLS_EPS] != 0) break;  // this state replaces the point
    }
    if (!PROBE && stepped && pos >= resume_at) break;
    stepped = true;
    if (STOP_REWIND && SPEC && !first_window) break;
    if (pos >= N) {
      if (!c.final_input) { flags |= WS_DONE; break; }  // shard of a longer stream: stop at the loop top
      // ---- EOF tail (matrix.go:650-678) ----
      if (N > 0 && hw < N - 1) hw = N - 1;
      uint32_t e = c.table[(t << c.row_shift) | K_CLS_EPS];
      if (e == 0) {
        if (eps_state == 0) {
          if (pos > tstart) {  // flush the last token (:671-678)
            if (pos < stop) set_bit(c.b_end, pos); else flags |= WS_PEND;
            DATOK_REWIND(pos);
            tstart = pos;
          } else if (!first_window && window_overflow(c, base, hw)) { err = E_OVERFLOW; goto out; }
          flags |= WS_DONE;
          break;
        }
        const uint32_t t0 = eps_state;  // :660-667
        eps_state = 0;
        pos = eps_pos;
        if (pos < stop) probing = false;
        e = c.table[(t0 << c.row_shift) | K_CLS_EPS];
      }
      // epsilon transition taken at `pos` (matrix.go:563-576)
      if (pos > tstart) {
        if (pos < stop) set_bit(c.b_end, pos);
        DATOK_REWIND(pos);
        tstart = pos;
      } else if (pos < stop) {
        if (get_bit(c.b_sent, pos)) { err = E_DEGENERATE; goto out; }
        set_bit(c.b_sent, pos);
      }
      t = e & 0x7FFFu;
      continue;
    }
    {
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
      g_exact_steps++;
#endif
      const uint32_t cl = cls_at(c, pos);
      const uint16_t* row = c.table + ((size_t)t << c.row_shift);
      if (c.hist) {
#if defined(__CUDA_ARCH__)
        atomicAdd(&c.hist[t], 1u);
        atomicAdd(&c.hist_cls[cl], 1u);
#else
        c.hist[t]++;
        c.hist_cls[cl]++;
#endif
      }
      if (row[K_CLS_EPS] != 0) { eps_state = t; eps_pos = pos; flags &= ~WS_EPS_OVER_EOT; }  // :442-454
      const uint32_t nt = row[cl];                                // :463
      if (nt != 0) {
        // ---- transition consumes the byte (matrix.go:579-605) ----
        const uint32_t before = pos;
        pos = before + 1;
        if (tstart == before && (nt & K_NT)) {  // leading non-token rune (:584-588)
          if (before < stop) set_bit(c.b_skip, before);
          tstart = pos;
        }
        if (cl == K_CLS_EOT) {  // :593-605
          if (c.eot_rewind) {
            if (probing) { eps_state = 0; continue; }  // the rewind kills the pending epsilon point
            set_bit(c.b_tend, before);
            if (hw < before) hw = before;
            if (tstart > pos) {  // stale bufft is reset by the rewind (:622)
              clear_range(c.b_skip, pos, tstart < stop ? tstart : stop);
            }
            DATOK_REWIND(pos);
            tstart = pos;
          } else {  // datok.go:1019-1030: SentenceEnd / TextEnd, the buffer stays as it is
            if (!probing) {
              set_bit(c.b_tend, before);
              if (hw < before) hw = before;
            }
            if (eps_state != 0 && eps_pos <= before) flags |= WS_EPS_OVER_EOT;
          }
        }
        t = nt & 0x7FFFu;
        continue;
      }
    }
    // ---- no transition (matrix.go:472-557); the unknown retry (:478-485) cannot succeed ----
    if (hw < pos) hw = pos;
    if (eps_state != 0) {
      // backtrack to the last state that had an epsilon transition (:487-497)
      if (flags & WS_EPS_OVER_EOT) { err = E_DEGENERATE; goto out; }  // (the reference would fire the TextEnd twice: not representable)
      const uint32_t t0 = eps_state;
      eps_state = 0;
      pos = eps_pos;
      if (pos < stop) probing = false;
      const uint32_t e = c.table[(t0 << c.row_shift) | K_CLS_EPS];
      if (pos > tstart) {  // :565-572
        if (pos < stop) set_bit(c.b_end, pos);
        DATOK_REWIND(pos);
        tstart = pos;
      } else if (pos < stop) {  // :573-576
        if (get_bit(c.b_sent, pos)) { err = E_DEGENERATE; goto out; }
        set_bit(c.b_sent, pos);
      }
      t = e & 0x7FFFu;
      continue;
    }
    // hard fail: drop at least one rune as a token and restart at the root (:499-552).
    // (cannot happen while probing: probing implies a pending epsilon point)
    if (pos <= tstart) {  // buffc-bufft <= 0 -> buffc++ (one RUNE)
      pos++;
      while (pos < N && cls_at(c, pos) == K_CLS_CONT) pos++;
      if (hw < pos - 1) hw = pos - 1;
    }
    if (tstart >= pos) { err = E_DEGENERATE; goto out; }  // empty or negative slice (token_writer.go:85)
    if (pos < stop) set_bit(c.b_end, pos); else flags |= WS_PEND;
    DATOK_REWIND(pos);
    tstart = pos;
    t = c.start;  // :548 t = uint32(1)
  }
out:
#undef DATOK_REWIND
  if (err) {
    st.flags = WS_INVALID | (err << WS_ERR_SHIFT);
    return err;
  }
  if (probing) {  // hand over the snapshot of the final arrival; its epsilon point is dead
    pos = s_pos; tstart = s_tstart; base = s_base; hw = s_hw; t = s_t; flags = s_flags;
    first_window = s_first_window;
    eps_state = 0;
  }
  if (SPEC) { if (first_window) { spec->first_hw = hw; spec->had_rewind = 0; } }
  if (!eps_state) flags &= ~WS_EPS_OVER_EOT;
  st.pos = pos; st.tstart = tstart; st.eps_pos = eps_state ? eps_pos : 0; st.base = base; st.hw = hw;
  st.t = (uint16_t)t; st.eps_state = (uint16_t)eps_state; st.flags = flags;
  return 0;
}

template <bool SPEC, bool PROBE, bool STOP_REWIND>
DATOK_HD_SLOW uint32_t walk_run(const WalkCtx& c, WState& st, uint32_t stop, SpecInfo* spec,
                                uint32_t resume_at = 0xFFFFFFFFu) {
  return walk_run_inl<SPEC, PROBE, STOP_REWIND>(c, st, stop, spec, resume_at);
}

DATOK_HD bool sync_class(const uint32_t* mask, uint32_t cl) { return (mask[cl >> 5] >> (cl & 31)) & 1u; }

}  // namespace datok
