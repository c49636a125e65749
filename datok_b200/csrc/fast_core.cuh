// fast_core.cuh -- the fast path of the speculative chunk walk (K2a).
//
// One lane walks one chunk in 32-byte SEGMENTS.  Per segment the lane
//   1. classifies its 32 raw bytes (UTF-8 decode + sigma map, matrix.go:388-435)
//      into a small class buffer (shared memory in the kernel) and produces the
//      rune-start word,
//   2. steps through the classes with ONE fused-table lookup per byte
//      (model.hpp: T2 folds "fail -> epsilon at this position -> re-read",
//      matrix.go:472-497,563-576), accumulating the boundary bits of the segment
//      in registers,
//   3. stores the four boundary words of the segment.
// Real backtracks (to an epsilon point recorded at an earlier byte, matrix.go:487-497)
// are handled in place when the point lies in the current segment; anything else
// that is rare -- far backtracks, hard fails, stale buffer offsets, the window-limit
// vicinity, chunk hand-off, EOF -- drops to the exact walker walk_run() for the rest
// of the segment.  Both walkers are step-equivalent, so a lane can switch between
// them at any loop top.
#pragma once
#include <stdint.h>
#include "walk_core.cuh"

namespace datok {

constexpr uint32_t SEG = 32;                 // bytes per segment = bits per bitmap word
constexpr uint32_t T2K_SHIFT = 16;
constexpr uint32_t T2_EPS = 1u << 18;
constexpr uint32_t T2_SLOWMARK = 1u << 31;
constexpr uint32_t FAST_WINDOW_GUARD = 960;  // stay exact when the buffer window could reach 1024 runes

struct FastTables {
  const uint32_t* hot;     // fused rows of states 0..n_hot-1 (shared memory in the kernel)
  const uint32_t* cold;    // the full fused table (global memory)
  uint32_t n_hot;
  uint32_t stride;         // entries per row (odd)
};

DATOK_HD uint32_t t2_lookup(const FastTables& T, uint32_t t, uint32_t cl) {
  const uint32_t idx = t * T.stride + cl;
  return t < T.n_hot ? T.hot[idx] : T.cold[idx];
}

// eps_b: [14:0] state at the loop top where the point was recorded, [17:16] epsilon
// steps taken there before the byte was consumed, [18] a token was pending, [19] valid
constexpr uint32_t EB_PENDING = 1u << 18, EB_VALID = 1u << 19;

#if defined(__CUDACC__)
#define DATOK_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define DATOK_UNLIKELY(x) __builtin_expect(!!(x), 0)
#endif

struct FastLane {
  uint32_t pos, tstart;
  uint32_t base;                 // last rewind point known at the START of the current segment;
                                 // rewinds inside the segment are read off the END/TEND bits (lane_base)
  uint32_t t;
  uint32_t eps_pos, eps_b;
  uint32_t hw_med;               // furthest failing position seen by an in-place backtrack (never reset:
                                 // an over-estimate from an older window cannot fake an overflow, see to_exact)
  uint32_t first_hw;             // SpecInfo of a guessed start: hw at the first rewind
  uint32_t first_window;         // 1 until the first rewind of a guessed start
};

struct SegBits {
  uint32_t end, skip, sent, tend;
};

enum { FAST_OK = 0, FAST_SLOW = 1 };

// last rewind point, taking the boundary bits of the current segment into account
DATOK_HD uint32_t lane_base(const FastLane& L, const SegBits& B, uint32_t seg_start) {
  uint32_t base = L.base;
  if (B.end) { const uint32_t p = seg_start + 31u - clz32(B.end); if (p > base) base = p; }
  if (B.tend) { const uint32_t p = seg_start + 32u - clz32(B.tend); if (p > base) base = p; }
  return base;
}
// a guessed start's first window closes with the first END/TEND bit of the lane
DATOK_HD void lane_note_first_rewind(FastLane& L, const SegBits& B, uint32_t seg_start) {
  const uint32_t m = B.end | B.tend;
  if (L.first_window && m) { L.first_hw = seg_start + ctz32(m); L.first_window = 0; }
}

// exact state -> fast lane.  Requires can_go_fast(st).
DATOK_HD bool can_go_fast(const WState& st) {
  return (st.flags & ~WS_PEND) == 0 && st.tstart <= st.pos;
}
DATOK_HD void to_fast(const WState& st, FastLane& L) {
  L.pos = st.pos; L.tstart = st.tstart; L.base = st.base; L.t = st.t;
  L.eps_pos = st.eps_pos;
  L.eps_b = st.eps_state ? (EB_VALID | st.eps_state | (st.eps_pos > st.tstart ? EB_PENDING : 0)) : 0;
  L.hw_med = st.hw;
}
// fast lane -> exact state (resolves the lazily stored epsilon point and window base)
DATOK_HD void to_exact(const FastLane& L, const SegBits& B, uint32_t seg_start, const FastTables& T, WState& st) {
  const uint32_t base = lane_base(L, B, seg_start);
  st.pos = L.pos; st.tstart = L.tstart; st.base = base; st.t = (uint16_t)L.t;
  st.flags = 0;
  uint32_t es = 0;
  if (L.eps_b & EB_VALID) {
    es = L.eps_b & 0x7FFFu;
    for (uint32_t k = (L.eps_b >> T2K_SHIFT) & 3u; k; k--) es = t2_lookup(T, es, K_CLS_EPS) & 0x7FFFu;
  }
  st.eps_state = (uint16_t)es;
  st.eps_pos = es ? L.eps_pos : 0;
  // hw: furthest byte read in the current buffer window.  hw_med may stem from an older window;
  // it was a real read then, [base, hw_med] is contained in that older window, so counting its
  // runes can never exceed what the reference itself had buffered.
  uint32_t hw = base;
  if (L.hw_med > hw) hw = L.hw_med;
  if (L.pos > base && hw < L.pos - 1) hw = L.pos - 1;
  st.hw = hw;
}

// In-place backtrack to the epsilon point recorded in this segment (matrix.go:487-497),
// or FAST_SLOW (nothing changed) when the exact walker has to take over.
DATOK_HD int fast_backtrack(FastLane& L, const FastTables& T, uint32_t seg_start, SegBits& B) {
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
  if (!(L.eps_b & EB_VALID)) g_hard++; else if (L.eps_pos < seg_start) g_far++;
#endif
  if (!(L.eps_b & EB_VALID) || L.eps_pos < seg_start) return FAST_SLOW;  // hard fail / far backtrack
  const uint32_t bbit = 1u << (L.eps_pos - seg_start);
  const bool pending = (L.eps_b & EB_PENDING) != 0;
  if (!pending && (B.sent & bbit)) return FAST_SLOW;  // second SentenceEnd at one position
  uint32_t cur = L.eps_b & 0x7FFFu;
  for (uint32_t k = (L.eps_b >> T2K_SHIFT) & 3u; k; k--) cur = t2_lookup(T, cur, K_CLS_EPS) & 0x7FFFu;
  const uint32_t tgt = t2_lookup(T, cur, K_CLS_EPS) & 0x7FFFu;
  if (L.hw_med < L.pos) L.hw_med = L.pos;
  if (pending) {  // Token + rewind (:565-572)
    if (L.first_window) {
      const uint32_t m = (B.end | B.tend) & (bbit - 1u);
      L.first_hw = m ? seg_start + ctz32(m) : L.hw_med;
      L.first_window = 0;
    }
    B.end |= bbit;
    L.tstart = L.eps_pos;
  } else {
    B.sent |= bbit;  // SentenceEnd (:573-576)
  }
  L.pos = L.eps_pos;
  L.eps_b = 0;
  L.t = tgt;
  return FAST_OK;
}

// One loop-top iteration of the reference at L.pos, which must lie in the segment
// [seg_start, seg_start+32) whose classes are seg_cls[0..31].  On FAST_SLOW nothing
// has been changed and the exact walker must take over at L.pos.
DATOK_HD int fast_step(FastLane& L, const FastTables& T, const uint8_t* seg_cls, uint32_t seg_start, SegBits& B) {
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
  g_fast++;
#endif
  const uint32_t off = L.pos - seg_start;
  const uint32_t cl = seg_cls[off];
  const uint32_t e = t2_lookup(T, L.t, cl);
  const uint32_t bit = 1u << off;
  if (DATOK_UNLIKELY((int32_t)e <= 0)) {  // 0: failure without epsilon transition; bit 31: leave to walk_run
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
    g_bt++;
#endif
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
    if (e != 0) g_mark++;
#endif
    if (e != 0) return FAST_SLOW;
    return fast_backtrack(L, T, seg_start, B);
  }
  const uint32_t k = (e >> T2K_SHIFT) & 3u;
  const bool is_eps = k != 0;
  const bool pending = L.pos > L.tstart;
  const bool two = k == 2;
  // k epsilon steps before the byte is consumed: Token if something is pending, else SentenceEnd
  {
    // a second SentenceEnd at one position is left to walk_run (it reports the degenerate case)
    const uint32_t sent_here = (B.sent & bit) != 0;
    const uint32_t bad = (uint32_t)is_eps & ((sent_here & (uint32_t)!(pending && !two)) | (uint32_t)(two && !pending));
    if (DATOK_UNLIKELY(bad)) return FAST_SLOW;
  }
  B.end |= (is_eps && pending) ? bit : 0u;
  B.sent |= ((is_eps && !pending) || two) ? bit : 0u;
  if (is_eps) { L.tstart = L.pos; L.eps_b = 0; }
  if (e & T2_EPS) {
    L.eps_pos = L.pos;
    L.eps_b = EB_VALID | L.t | (e & (3u << T2K_SHIFT)) | ((!is_eps && pending) ? EB_PENDING : 0u);
  }
  const uint32_t next = L.pos + 1;
  const bool skip = (L.tstart == L.pos) && (e & K_NT);  // :584-588
  B.skip |= skip ? bit : 0u;
  L.tstart = skip ? next : L.tstart;
  if (DATOK_UNLIKELY(cl == K_CLS_EOT)) {  // :593-605 (the forced SentenceEnd is derived by the compaction)
    B.tend |= bit;
    L.tstart = next; L.eps_b = 0;
  }
  L.t = e & 0x7FFFu;
  L.pos = next;
  return FAST_OK;
}

// loads the 32 raw bytes at `p` (zero padded beyond N) as 8 little-endian words
DATOK_HD void load_segment_words(const uint8_t* in, uint32_t N, uint32_t seg_start, uint32_t* words) {
  const uint8_t* p = in + seg_start;
  if (seg_start + SEG <= N && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
#if defined(__CUDA_ARCH__)
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const uint4 c = *reinterpret_cast<const uint4*>(p + 16);
    words[0] = a.x; words[1] = a.y; words[2] = a.z; words[3] = a.w;
    words[4] = c.x; words[5] = c.y; words[6] = c.z; words[7] = c.w;
#else
    for (int k = 0; k < 8; k++) {
      words[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) |
                 ((uint32_t)p[4 * k + 3] << 24);
    }
#endif
    return;
  }
  for (int k = 0; k < 8; k++) {
    uint32_t v = 0;
    for (int j = 0; j < 4; j++) {
      const uint32_t q = seg_start + 4 * k + j;
      if (q < N) v |= (uint32_t)in[q] << (8 * j);
    }
    words[k] = v;
  }
}

// Classes and rune starts of the 32 bytes at seg_start (bytes >= N: no rune start,
// class unspecified).  seg_cls must be 4-byte aligned.  ASCII bytes go through the
// LUT four at a time; the few other bytes are decoded afterwards, one rune each.
DATOK_HD void classify_segment(const uint8_t* in, uint32_t N, uint32_t seg_start, const ClsTables& T,
                               uint8_t* seg_cls, uint32_t* rstart_word, bool* any_invalid) {
  uint32_t words[8];
  load_segment_words(in, N, seg_start, words);
  uint32_t* out = reinterpret_cast<uint32_t*>(seg_cls);
  uint32_t nonascii = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 8; k++) {
    const uint32_t v = words[k];
    const uint32_t c0 = T.ascii_cls[v & 0x7Fu], c1 = T.ascii_cls[(v >> 8) & 0x7Fu];
    const uint32_t c2 = T.ascii_cls[(v >> 16) & 0x7Fu], c3 = T.ascii_cls[(v >> 24) & 0x7Fu];
    out[k] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
    const uint32_t h = v & 0x80808080u;
    nonascii |= (((h >> 7) | (h >> 14) | (h >> 21) | (h >> 28)) & 0xFu) << (4 * k);
  }
  const uint32_t valid = (seg_start + SEG <= N) ? 0xFFFFFFFFu : mask_below(N > seg_start ? N - seg_start : 0);
  uint32_t rs = ~nonascii & valid;
  uint32_t m = nonascii & valid;
  while (m) {
    const uint32_t j = ctz32(m);
    m &= m - 1;
    const uint32_t p = seg_start + j;
    bool st, inv;
    const uint32_t cl = classify_pos(in, N, p, T, &st, &inv);
    seg_cls[j] = (uint8_t)cl;
    if (st) rs |= 1u << j;
    if (inv) *any_invalid = true;
    if (st && !inv) {
      // a well-formed multi-byte rune: its continuation bytes need no decoding of their own
      const uint32_t b0 = in[p];
      const uint32_t wdt = b0 >= 0xF0 ? 4u : b0 >= 0xE0 ? 3u : 2u;
      for (uint32_t q = 1; q < wdt && j + q < SEG; q++) {
        seg_cls[j + q] = (uint8_t)K_CLS_CONT;
        m &= ~(1u << (j + q));
      }
    }
  }
  *rstart_word = rs;
}

// first sync point in (from, from+32] given the raw bytes: a position whose
// preceding byte is an ASCII byte the root state skips.  K_NOPOS if none.
DATOK_HD uint32_t find_sync(const uint8_t* in, uint32_t N, const uint32_t* sync_ascii, uint32_t lo, uint32_t hi) {
  // positions p in [lo, hi) with p >= 1, p < N, in[p-1] a sync byte
  for (uint32_t p = lo ? lo : 1; p < hi && p < N; p++) {
    const uint32_t b = in[p - 1];
    if (b < 0x80 && ((sync_ascii[b >> 5] >> (b & 31)) & 1u)) return p;
  }
  return K_NOPOS;
}

}  // namespace datok
