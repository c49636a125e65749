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

struct FastLane {
  uint32_t pos, tstart, base;
  uint32_t t;
  uint32_t eps_pos, eps_b;
  uint32_t hw_med, hw_med_base;  // furthest failing position seen by an in-place backtrack since `hw_med_base`
};

struct SegBits {
  uint32_t end, skip, sent, tend;
};

enum { FAST_OK = 0, FAST_SLOW = 1 };

// exact state -> fast lane.  Requires can_go_fast(st).
DATOK_HD bool can_go_fast(const WState& st) {
  return (st.flags & ~WS_PEND) == 0 && st.tstart <= st.pos;
}
DATOK_HD void to_fast(const WState& st, FastLane& L) {
  L.pos = st.pos; L.tstart = st.tstart; L.base = st.base; L.t = st.t;
  L.eps_pos = st.eps_pos;
  L.eps_b = st.eps_state ? (EB_VALID | st.eps_state | (st.eps_pos > st.tstart ? EB_PENDING : 0)) : 0;
  L.hw_med = st.hw; L.hw_med_base = st.base;
}
// fast lane -> exact state (resolves the lazily stored epsilon point)
DATOK_HD void to_exact(const FastLane& L, const FastTables& T, WState& st) {
  st.pos = L.pos; st.tstart = L.tstart; st.base = L.base; st.t = (uint16_t)L.t;
  st.flags = 0;
  uint32_t es = 0;
  if (L.eps_b & EB_VALID) {
    es = L.eps_b & 0x7FFFu;
    for (uint32_t k = (L.eps_b >> T2K_SHIFT) & 3u; k; k--) es = t2_lookup(T, es, K_CLS_EPS) & 0x7FFFu;
  }
  st.eps_state = (uint16_t)es;
  st.eps_pos = es ? L.eps_pos : 0;
  uint32_t hw = L.base;
  if (L.hw_med_base == L.base && L.hw_med > hw) hw = L.hw_med;
  if (L.pos > L.base && hw < L.pos - 1) hw = L.pos - 1;
  st.hw = hw;
}

// One loop-top iteration of the reference at L.pos, which must lie in the segment
// [seg_start, seg_start+32) whose classes are seg_cls[0..31].  On FAST_SLOW nothing
// has been changed and the exact walker must take over at L.pos.
DATOK_HD int fast_step(FastLane& L, const FastTables& T, const uint8_t* seg_cls, uint32_t seg_start, SegBits& B) {
  const uint32_t off = L.pos - seg_start;
  const uint32_t cl = seg_cls[off];
  const uint32_t e = t2_lookup(T, L.t, cl);
  const uint32_t bit = 1u << off;
  if (e == 0) {
    // failure in a state without epsilon transition: backtrack to the recorded point
    // (matrix.go:487-497) if it lies in this segment
    if (!(L.eps_b & EB_VALID) || L.eps_pos < seg_start) return FAST_SLOW;
    const uint32_t bbit = 1u << (L.eps_pos - seg_start);
    const bool pending = (L.eps_b & EB_PENDING) != 0;
    if (!pending && (B.sent & bbit)) return FAST_SLOW;  // second SentenceEnd at one position
    uint32_t cur = L.eps_b & 0x7FFFu;
    for (uint32_t k = (L.eps_b >> T2K_SHIFT) & 3u; k; k--) cur = t2_lookup(T, cur, K_CLS_EPS) & 0x7FFFu;
    const uint32_t tgt = t2_lookup(T, cur, K_CLS_EPS) & 0x7FFFu;
    if (L.hw_med_base != L.base) { L.hw_med = 0; L.hw_med_base = L.base; }
    if (L.hw_med < L.pos) L.hw_med = L.pos;
    L.pos = L.eps_pos;
    if (pending) { B.end |= bbit; L.base = L.pos; L.tstart = L.pos; }  // Token + rewind (:565-572)
    else B.sent |= bbit;                                                // SentenceEnd (:573-576)
    L.eps_b = 0;
    L.t = tgt;
    return FAST_OK;
  }
  if (e & T2_SLOWMARK) return FAST_SLOW;
  const uint32_t k = (e >> T2K_SHIFT) & 3u;
  if (k) {
    const bool pending = L.pos > L.tstart;
    if ((B.sent & bit) && !(pending && k == 1)) return FAST_SLOW;  // would repeat a SentenceEnd here
    if (k == 2 && !pending) return FAST_SLOW;
    if (pending) { B.end |= bit; L.base = L.pos; } else B.sent |= bit;
    if (k == 2) B.sent |= bit;
    L.tstart = L.pos;
    L.eps_b = 0;
  }
  if (e & T2_EPS) {
    L.eps_pos = L.pos;
    L.eps_b = EB_VALID | L.t | (k << T2K_SHIFT) | ((k == 0 && L.pos > L.tstart) ? EB_PENDING : 0);
  }
  const uint32_t next = L.pos + 1;
  if (L.tstart == L.pos && (e & K_NT)) { B.skip |= bit; L.tstart = next; }  // :584-588
  if (cl == K_CLS_EOT) {  // :593-605 (the forced SentenceEnd is derived by the compaction)
    B.tend |= bit;
    L.base = next; L.tstart = next; L.eps_b = 0;
  }
  L.t = e & 0x7FFFu;
  L.pos = next;
  return FAST_OK;
}

// Classes and rune starts of the 32 bytes at seg_start (bytes >= N: class 0, no start).
DATOK_HD void classify_segment(const uint8_t* in, uint32_t N, uint32_t seg_start, const ClsTables& T,
                               uint8_t* seg_cls, uint32_t* rstart_word, bool* any_invalid) {
  uint32_t rs = 0;
  for (uint32_t j = 0; j < SEG; j++) {
    const uint32_t p = seg_start + j;
    uint32_t cl = 0;
    if (p < N) {
      const uint32_t b = in[p];
      if (b < 0x80) { cl = T.ascii_cls[b]; rs |= 1u << j; }
      else {
        bool st, inv;
        cl = classify_pos(in, N, p, T, &st, &inv);
        if (st) rs |= 1u << j;
        if (inv) *any_invalid = true;
      }
    }
    seg_cls[j] = (uint8_t)cl;
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
