// chunk_core.cuh -- per-chunk bodies of the speculative chunked walk (K2a-K2d).
//
// The reference walks one stream sequentially (matrix.go:384-635).  Here the
// input is cut into fixed chunks of `chunk` bytes, one lane per chunk:
//   spec    every lane but lane 0 starts at the first SYNC point of its chunk
//           (a byte following a rune that the root state skips as whitespace)
//           with the guessed state (root, nothing pending) and walks to the end
//           of its chunk;
//   stitch  lane i then walks the head of its chunk [i*chunk, sync) from the exit
//           state of lane i-1 and compares what it arrives with against the guess;
//   rewalk  on a mismatch (or if the chunk has no sync point) lane i re-walks the
//           rest of its chunk from the true state, replacing the speculative trace;
//   commit  if a lane's exit state changed, its successor is queued for the next
//           round.  Rounds repeat until no exit state changes; at that fixpoint
//           every chunk's bits equal those of the sequential walk.
// A lane writes boundary bits only inside its own chunk (walk_run's hand-off rule),
// so no atomics are needed.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "compact_core.cuh"
#include "fast_core.cuh"
#include "walk_core.cuh"

namespace datok {

constexpr uint32_t CF_HAD_REWIND = 1;   // the speculative walk closed at least one buffer window
constexpr uint32_t CF_OVERWRITTEN = 2;  // the speculative trace has been replaced by a re-walk

struct DeviceModel {
  const uint16_t* table;   // exact table (walk_run)
  const uint32_t* table2;  // fused table T3 (fast_run), stride2 entries per row
  const uint16_t* hot16;   // compact rows of the first hot16_rows states, stride16 entries per row
  uint32_t row_shift, start, n_classes, stride2, stride16, hot16_rows, hot_cols;
  uint32_t eot_rewind;     // 0: a double-array model (datok.go): an EOT does not rewind the buffer
  ClsTables cls;           // pointers into device memory
  uint32_t sync_ascii[4];  // ASCII bytes the root state skips: a chunk may start right after one
  uint32_t sync_cls[8];    // the same set as classes (HostModel.sync_mask)
};

struct WalkBuffers {
  const uint8_t* in;
  uint32_t N;
  uint32_t chunk;          // bytes per chunk (multiple of 32)
  uint32_t n_chunks;       // N / chunk + 1
  uint32_t n_words;        // n_chunks * chunk / 32
  uint32_t *rstart, *b_end, *b_skip, *b_sent, *b_tend;
  WState *E, *exitA, *Enew, *Ytmp;
  uint32_t *sync, *first_hw, *cflags;
  uint32_t *list_cur, *list_next, *list_rewalk;
  uint32_t* counters;      // [0] next list size, [1] rewalk list size, [2] invalid-utf8 flag
  unsigned long long* err_key;
  uint32_t final_input;    // see WalkCtx
};

DATOK_HD WalkCtx make_walk_ctx(const DeviceModel& m, const WalkBuffers& b) {
  WalkCtx c;
  c.table = m.table; c.row_shift = m.row_shift; c.start = m.start;
  c.in = b.in; c.N = b.N; c.cls = m.cls;
  c.b_end = b.b_end; c.b_skip = b.b_skip; c.b_sent = b.b_sent; c.b_tend = b.b_tend;
  c.hist = nullptr; c.hist_cls = nullptr;
  c.eot_rewind = m.eot_rewind;
  c.final_input = b.final_input;
  return c;
}

DATOK_HD WState wstate_invalid(uint32_t err) {
  WState s;
  s.pos = s.tstart = s.eps_pos = s.base = s.hw = 0;
  s.t = 0; s.eps_state = 0;
  s.flags = WS_INVALID | (err << WS_ERR_SHIFT);
  return s;
}

DATOK_HD void clear_chunk_bits(const WalkBuffers& b, uint32_t lo, uint32_t hi) {
  clear_range(b.b_end, lo, hi);
  clear_range(b.b_skip, lo, hi);
  clear_range(b.b_sent, lo, hi);
  clear_range(b.b_tend, lo, hi);
}

// K2a: speculative walk of chunk i.  start_state: GPU id of the stream's first state.
DATOK_HD void chunk_spec(const DeviceModel& m, const WalkBuffers& b, uint32_t i, uint32_t start_state) {
  const WalkCtx c = make_walk_ctx(m, b);
  const uint32_t lo = i * b.chunk, hi = lo + b.chunk;
  WState st;
  st.eps_pos = 0; st.eps_state = 0; st.flags = 0;
  SpecInfo si;
  si.first_hw = 0; si.had_rewind = 0;
  b.cflags[i] = 0;
  if (i == 0) {  // the stream start is not a guess
    st.pos = st.tstart = st.base = st.hw = 0;
    st.t = (uint16_t)start_state;
    b.sync[0] = 0;
    walk_run<false, true, false>(c, st, hi, &si);
    b.exitA[0] = st;
    b.E[0] = st;
    b.first_hw[0] = 0;
    b.cflags[0] = CF_HAD_REWIND;
    return;
  }
  const uint32_t s = find_sync(b.in, b.N, m.sync_ascii, lo, hi);
  b.sync[i] = s;
  if (s == K_NOPOS) {  // no sync point: the predecessor's state has to be walked through
    st = wstate_invalid(0);
  } else {
    st.pos = st.tstart = st.base = st.hw = s;
    st.t = (uint16_t)m.start;
    walk_run<true, true, false>(c, st, hi, &si);
  }
  b.exitA[i] = st;
  b.E[i] = st;
  b.first_hw[i] = si.first_hw;
  b.cflags[i] = si.had_rewind ? CF_HAD_REWIND : 0;
}

DATOK_HD void note_invalid_utf8(const WalkBuffers& b) {
#if defined(__CUDA_ARCH__)
  atomicOr(&b.counters[2], 1u);
#else
  b.counters[2] |= 1u;
#endif
}

// A lane fills a 32-byte sector of each bitmap with 8 word stores spread over 8 segments; the partly
// written sectors are asked to stay in L2 until then (evict_last) instead of going to DRAM piecemeal.
DATOK_HD void store_word_keep(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__) && !defined(DATOK_BITS_PLAIN_STORE)
  asm volatile(
      "{\n\t.reg .b64 pol;\n\t"
      "createpolicy.fractional.L2::evict_last.b64 pol, 1.0;\n\t"
      "st.global.L2::cache_hint.u32 [%0], %1, pol;\n\t}"
      :: "l"(p), "r"(v) : "memory");
#else
  *p = v;
#endif
}
#if !defined(__CUDA_ARCH__)
// host builds (tests/emul): the word a lane is about to store is checked against its chunk -- a store into
// a neighbour's words is invisible to a sequential emulation but a race on the GPU
static thread_local uint32_t g_own_lo = 0, g_own_hi = 0xFFFFFFFFu;
#endif
DATOK_HD void store_seg_bits(const WalkBuffers& b, uint32_t w, const SegBits& B) {
#if !defined(__CUDA_ARCH__)
  if (w < g_own_lo || w >= g_own_hi) { fprintf(stderr, "store_seg_bits: word %u outside [%u, %u)\n", w, g_own_lo, g_own_hi); abort(); }
#endif
  store_word_keep(b.b_end + w, B.end); store_word_keep(b.b_skip + w, B.skip);
  store_word_keep(b.b_sent + w, B.sent); store_word_keep(b.b_tend + w, B.tend);
}
DATOK_HD void load_seg_bits(const WalkBuffers& b, uint32_t w, SegBits& B) {
  B.end = b.b_end[w]; B.skip = b.b_skip[w]; B.sent = b.b_sent[w]; B.tend = b.b_tend[w];
}

// Short look-ahead of the hand-off: what becomes of the epsilon point at eps_pos < hi that the lane
// (state t at `hi`) still holds?
//   LOOK_DEAD   a later state has an epsilon transition of its own and replaces it, or a boundary / a
//               consumed EOT kills it (walk_run, PROBE): the arrival state stands
//   LOOK_EXACT  left to the exact probe: a lookup fails within a few bytes of the point (the exact walker
//               re-walks that short distance at less cost than the lanes of the warp waiting for this one
//               to repeat whole segments), or a rare case, the window limit, the end of the input
//   LOOK_PROBE  still alive after LOOK bytes, or a failure far from the point (long tokens: compounds,
//               markup): the lane reads on through the fast path
enum { LOOK_DEAD = 0, LOOK_EXACT = 1, LOOK_PROBE = 2 };
#ifndef DATOK_NEAR_BACKTRACK
#define DATOK_NEAR_BACKTRACK 8
#endif
constexpr uint32_t NEAR_BACKTRACK = DATOK_NEAR_BACKTRACK;  // bytes: up to this distance the exact walker re-walks
#if defined(DATOK_NI_LOOKAHEAD)
DATOK_HD_SLOW
#else
DATOK_HD
#endif
int look_ahead(const FastTables& FT, const WalkCtx& c, uint32_t t, uint32_t hi, uint32_t base, uint32_t eps_pos) {
  constexpr uint32_t LOOK = 16;
  if (eps_target(FT, t) != 0) return LOOK_DEAD;
  if (hi + LOOK > c.N || hi + LOOK - base >= FAST_WINDOW_GUARD) return LOOK_EXACT;
  for (uint32_t k = 0; k < LOOK; k++) {
    const uint32_t cl = cls_at(c, hi + k);
    const uint32_t e3 = t3_load(FT, t, cl);
    if (e3 & F3_SLOWMARK) return LOOK_EXACT;
    if ((e3 & F3_TGT) == 0) return hi + k - eps_pos <= NEAR_BACKTRACK ? LOOK_EXACT : LOOK_PROBE;
    if ((e3 & (F3_KANY | F3_EA)) || (cl == K_CLS_EOT && FT.eot_rewind)) return LOOK_DEAD;
    t = e3 & F3_TGT;
  }
  return LOOK_PROBE;
}

// K1+K2a fused: classification and speculative walk of chunk i, segment by segment
// (fast_core.cuh).  Produces exactly what chunk_spec() produces, plus the rune-start
// words of the chunk.  seg_cls: 32 bytes of lane-private scratch (shared memory in
// the kernel).  Every boundary word of the chunk is written here: no clearing pass is needed.
//
// from == nullptr: speculative walk (K2a).  from != nullptr: re-walk (K2c) of the chunk from the
// known state *from (its position lies in the chunk); the result goes to Enew[i] instead and the
// rune-start words are left alone.
DATOK_HD void chunk_spec_fast(const DeviceModel& m, const WalkBuffers& b, const FastTables& FT, uint32_t i,
                              uint32_t start_state, uint8_t* seg_cls, const WState* from = nullptr, uint32_t stage_saddr = 0) {
  const WalkCtx c = make_walk_ctx(m, b);
  const uint32_t lo = i * b.chunk, hi = lo + b.chunk, N = b.N;
  const bool rewalk = from != nullptr;
  FastCtx FX;
  FX.in = b.in; FX.N = N; FX.cls = &m.cls;
#if defined(DATOK_STAGE_ASYNC)
  SegStage stage;
  stage_init(stage, stage_saddr);
#else
  (void)stage_saddr;
#endif
#if !defined(__CUDA_ARCH__)
  g_own_lo = lo >> 5; g_own_hi = hi >> 5;
#endif
  WState st;
  st.pos = st.tstart = st.base = st.hw = 0;
  st.eps_pos = 0; st.eps_state = 0; st.flags = 0; st.t = (uint16_t)start_state;
  FastLane L;
  L.pos = L.tstart = L.base = L.eps_p = L.eps_rec = L.hw_med = L.first_hw = 0;
  L.stale_end = L.raw_from = 0;
  L.u_in = 1;
  L.t = start_state;
  L.first_window = (i != 0 && !rewalk);  // a guessed start: the first window's overflow check is deferred (SpecInfo)
  RawBits R;
  raw_clear(R);
  bool started = (i == 0) || rewalk, fast = true, halted = false, sync_carry = false;
  uint32_t sync = (i == 0) ? 0u : K_NOPOS;
  uint32_t err = 0;
  SegBits B;
  B.end = B.skip = B.sent = B.tend = 0;
  uint32_t first_seg = lo;
  uint32_t walk_from = lo;  // first position of this walk: an epsilon point below it is not the lane's to go back to
  if (rewalk) {
    st = *from;
    walk_from = st.pos;
    fast = false;  // the fast path takes over as soon as the state allows it (see below)
    first_seg = st.pos & ~(SEG - 1);
    if (first_seg < lo) first_seg = lo;
  }

  // A backtrack to an epsilon point in an EARLIER segment (matrix.go:487-497) is taken by the exact walker
  // (one iteration); the lane then goes back to that segment -- first_seg again, its boundary words are in
  // memory like those of a re-walk's first segment -- and continues from there through the fast path.
  //
  // Hand-off through the fast path (the rule of walk_run's PROBE mode): at the chunk end an epsilon point
  // below `hi` may still be pending.  The lane reads on in the same loop -- `probing`, writing nothing --
  // until that point is dead: replaced by a newer one, or killed by a boundary / EOT.  If a lookup fails
  // while the point is alive, the exact walker takes that one backtrack and the lane walks on from the
  // point, inside its own chunk again.  Anything else (a rare case, the window limit, the end of the
  // input) is left to the exact probe after the loop, from the arrival state.
  // (One loop for everything, no jumps back into it: the lanes of a warp stay converged.)
  // (as few values as possible live across the hot loop: it is very sensitive to register pressure)
  enum { PH_WALK = 0, PH_PROBE = 1, PH_HANDED_OFF = 2, PH_GAVE_UP = 3 };
  uint32_t phase = PH_WALK;
  bool in_regs = !rewalk;  // the segment's boundary words live in B (else in memory: first segment of a re-walk)
  for (uint32_t seg_start = first_seg;; seg_start += SEG) {
    if (seg_start >= hi) {
      if (phase == PH_WALK) {
        // arrival at the chunk end: st is what the successor starts from, if it stands
        if (fast && !halted) { B.end = B.skip = B.sent = B.tend = 0; to_exact(L, B, hi, FT, st); }
        if (!fast || halted || L.first_window || hi >= N) break;  // the exact probe decides
        const int look = st.eps_state == 0 ? LOOK_DEAD : look_ahead(FT, c, L.t, hi, L.base, st.eps_pos);
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
        g_probe[2]++;
        if (st.eps_state == 0 || eps_target(FT, L.t) != 0) g_probe[3]++;
        if (look == LOOK_PROBE) g_probe[0]++;
#endif
        if (look == LOOK_DEAD) phase = PH_HANDED_OFF;
        if (look != LOOK_PROBE) break;
        phase = PH_PROBE;
      }
      if (seg_start - hi >= 1024u || seg_start + SEG > N || seg_start + SEG - L.base >= FAST_WINDOW_GUARD) break;
    }
    const uint32_t seg_end = seg_start + SEG, w = seg_start >> 5;
    uint32_t rs, eotm, nonascii;
    bool inv = false;
#if defined(DATOK_STAGE_ASYNC)
    classify_segment(b.in, N, seg_start, m.cls, FT.ascii_cls2, FT.stop_cl2, seg_cls, &rs, &eotm, &inv, &nonascii, &stage);
#else
    classify_segment(b.in, N, seg_start, m.cls, FT.ascii_cls2, FT.stop_cl2, seg_cls, &rs, &eotm, &inv, &nonascii);
#endif
    const uint32_t limit = seg_end < N ? seg_end : N;
    const uint32_t eotk = FT.eot_rewind ? eotm : 0u;  // the EOTs that rewind the buffer (none in the double-array walk)
#if defined(__CUDA_ARCH__)
    // the next segment on its way while this one is walked (no registers held): as a prefetch into L1, or
    // (build option DATOK_STAGE_ASYNC) into the lane's staging slot through the async copy unit
#if defined(DATOK_STAGE_ASYNC)
    if (seg_end < hi && stage.slot_saddr) stage_segment(stage, b.in, N, seg_end);
#else
    if (seg_end < hi && seg_end + SEG <= N) asm volatile("prefetch.global.L1 [%0];" :: "l"(b.in + seg_end));
#endif
#endif
    if (!rewalk && phase == PH_WALK) {
      store_word_keep(b.rstart + w, rs);
      if (inv) note_invalid_utf8(b);
    }
    // (the speculative walk writes every boundary word of its chunk, so the bitmaps need no clearing pass)
    B.end = B.skip = B.sent = B.tend = 0;
    if (halted) { if (!rewalk) store_seg_bits(b, w, B); continue; }
    if (!started) {
      // (from the class buffer; a sync point at the segment's first position is found through the last byte of
      // the segment before -- except in the chunk's first segment, where the next point does as well)
      sync = sync_carry ? seg_start : find_sync_cls(seg_cls, nonascii, FT.sync_cls, FT.stop_cl2, seg_start, 0, 31);
      if (sync != K_NOPOS && (sync >= hi || sync >= N)) sync = K_NOPOS;
      {
        const uint32_t c2 = seg_cls[31], c = c2 >> 1;
        sync_carry = !((nonascii >> 31) & 1u) && c2 != FT.stop_cl2 && ((FT.sync_cls[c >> 5] >> (c & 31)) & 1u);
      }
      if (sync == K_NOPOS) { store_seg_bits(b, w, B); continue; }
      started = true;
      walk_from = sync;
      L.pos = L.tstart = L.base = L.hw_med = L.raw_from = sync;
      L.u_in = 1;
      L.t = m.start;
    }
    B.end = B.skip = B.sent = B.tend = 0;
    if (fast && seg_end - L.base >= FAST_WINDOW_GUARD) {  // too close to the 1024-rune buffer limit
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
      g_guard++;
#endif
      to_exact(L, B, seg_start, FT, st);
      fast = false;
    }
    bool must_walk_exact = false;  // the fast path just gave up at st.pos: the exact walker has to move first
    bool reenter = false;          // a far backtrack through the fast path: back to the segment of L.pos
    if (fast && !in_regs) { load_seg_bits(b, w, B); in_regs = true; }  // (that segment: its words are in memory)
    for (;;) {
      if (fast) {
        const int rc = fast_run(L, R, FT, FX, seg_cls, seg_start, limit, eotk, B);
        if (!fast_flush(L, R, B, eotk, eotm, seg_start)) {  // two SentenceEnds at one position
          if (phase == PH_PROBE) { phase = PH_GAVE_UP; break; }
          err = E_DEGENERATE;
          st = wstate_invalid(E_DEGENERATE);
          halted = true;
          break;
        }
        if (rc == FAST_OK && L.pos >= seg_end) break;  // segment done, stay fast
        if (phase == PH_PROBE) {
          phase = PH_GAVE_UP;
          if (rc == FAST_SLOW_FAIL && L.eps_rec && L.eps_p < hi) {
            // (the flush has just confirmed that the point is alive: the failing iteration is this backtrack)
            to_exact(L, B, seg_start, FT, st);
            SpecInfo dummy;
            err = walk_run<false, false, false>(c, st, st.pos + 1, &dummy, 0);
            phase = PH_WALK;
            fast = false;
            // (stale bufft, matrix.go:573-576: the exact walker takes the lane to the chunk end; as above)
            if (!err && !can_go_fast(st)) err = walk_run<false, false, false>(c, st, hi, &dummy);
            if (err) halted = true;  // the stream itself is in error here
          }
          break;
        }
#if !defined(DATOK_NO_FAST_FAR)
        // A backtrack (matrix.go:487-497) to an epsilon point that lies before the raw range, in this or an
        // earlier segment of the lane's own walk, while a token is pending there -- the common far backtrack
        // (a word, then a few bytes that turn out not to belong to it, across a segment boundary): the
        // Token boundary goes into the stored words, and the lane walks on from the point through the fast
        // path, starting with the point's segment again (its words are in memory, like those of a re-walk's
        // first segment).  Nothing the first pass wrote behind the point has to be taken back: the point
        // being alive, there was no boundary and no EOT behind it, and a pending token means no skipped rune.
        if (rc == FAST_SLOW_FAIL && L.eps_rec && ((L.eps_rec >> 16) & 3u) == 0 && !L.first_window &&
            L.eps_p >= walk_from && L.tstart < L.eps_p && L.stale_end <= L.eps_p) {
          const uint32_t q = L.eps_p, tgt = eps_target(FT, L.eps_rec & F3_TGT);
          if (tgt != 0) {
            DATOK_STAT(g_bt_far_fast);
            if (q >= seg_start && in_regs) store_seg_bits(b, w, B);  // (the point's segment is this one)
            set_bit(b.b_end, q);
            if (L.hw_med < L.pos) L.hw_med = L.pos;
            L.pos = L.tstart = L.base = L.raw_from = q;
            L.u_in = 1;
            L.t = tgt;
            L.eps_rec = 0;
            reenter = true;
            break;
          }
        }
#endif
        lane_note_first_rewind(L, B, seg_start, FT.eot_rewind);
        to_exact(L, B, seg_start, FT, st);              // rare case, or end of input
        fast = false;
        must_walk_exact = true;
        if (st.pos >= seg_end) break;
      }
      // exact walker for the rest of the segment (or until the state allows the fast path again)
      if (!must_walk_exact && st.pos >= seg_start && st.pos < seg_end && st.pos < N && can_go_fast(st) &&
          seg_end - st.base < FAST_WINDOW_GUARD) {
        if (!in_regs) { load_seg_bits(b, w, B); in_regs = true; }
        if (st.flags & WS_PEND) { B.end |= 1u << (st.pos - seg_start); st.flags &= ~WS_PEND; }
        to_fast(st, L, R);
        fast = true;
        continue;
      }
      if (st.pos >= seg_end) break;
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
      if (must_walk_exact) g_nf[0]++;
      else if (st.pos < seg_start || st.pos >= N) g_nf[1]++;
      else if (st.flags & ~WS_PEND) g_nf[2]++;
      else if (st.tstart > st.pos) g_nf[3]++;
      else g_nf[4]++;
#endif
      if (in_regs) { store_seg_bits(b, w, B); in_regs = false; }
      SpecInfo si;
      si.first_hw = 0; si.had_rewind = 0;
      // (resume_at 0: back after one iteration, wherever it led)
      if (L.first_window) {
        err = walk_run<true, false, true>(c, st, seg_end, &si, 0);
        if (si.had_rewind) { L.first_hw = si.first_hw; L.first_window = 0; }
      } else {
        err = walk_run<false, false, false>(c, st, seg_end, &si, 0);
      }
      if (err || (st.flags & WS_DONE)) { halted = true; break; }
      must_walk_exact = false;
      // far backtrack (this segment's words are in memory).  A lane with a stale bufft (matrix.go:573-576)
      // stays with the exact walker instead: the skipped runes ahead of it keep the SKIP bits they have there
      // (a near one is re-walked by the exact walker, iteration by iteration: cheaper than the other lanes of
      // the warp waiting for this one to repeat whole segments)
      if (seg_start - st.pos > NEAR_BACKTRACK && st.pos < seg_start && can_go_fast(st)) break;
    }
    if (phase == PH_GAVE_UP) break;
    if (reenter) {
      seg_start = (L.pos & ~(SEG - 1)) - SEG;
      in_regs = false;
      continue;
    }
    if (!fast && !halted && st.pos < seg_start && seg_start - st.pos > NEAR_BACKTRACK) {  // back to the segment of the epsilon point: its words are in memory
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
      g_nf[2]++;
#endif
      seg_start = (st.pos & ~(SEG - 1)) - SEG;
      in_regs = false;
      continue;
    }
    if (seg_start >= hi) {  // beyond the chunk end: these words belong to the successor, nothing is stored
      // (not probing any more: the probe's backtrack led the exact walker to the chunk end again, or into an
      // error: the loop top deals with this arrival)
      if (phase != PH_PROBE) continue;
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
      g_probe[1]++;
#endif
      if (!L.eps_rec || L.eps_p >= hi) { phase = PH_HANDED_OFF; break; }
      L.base = lane_base(L, B, seg_start, FT.eot_rewind);
      continue;
    }
    if (fast && !halted) {
      lane_note_first_rewind(L, B, seg_start, FT.eot_rewind);
      L.base = lane_base(L, B, seg_start, FT.eot_rewind);
    }
    if (in_regs) store_seg_bits(b, w, B);
    in_regs = true;
  }

#if defined(__CUDA_ARCH__) && defined(DATOK_STAGE_ASYNC)
  if (stage.staged_for != K_NOPOS) stage_wait(stage);  // (the slot is reused by the lane's next chunk)
#if defined(DATOK_STAGE_TMA)
  if (stage.slot_saddr) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(stage.slot_saddr + 32u) : "memory");
#endif
#endif
  SpecInfo si;
  si.first_hw = L.first_hw; si.had_rewind = L.first_window ? 0u : 1u;
  if (!started) st = wstate_invalid(0);
  else if (!err && !(halted && (st.flags & WS_DONE))) {
    if (phase == PH_HANDED_OFF) { st.eps_state = 0; st.eps_pos = 0; }
    else {
      // hand-off at the chunk end (probe, walk_run) from the arrival state
      if (L.first_window) err = walk_run<true, true, false>(c, st, hi, &si);
      else { SpecInfo dummy; err = walk_run<false, true, false>(c, st, hi, &dummy); }
    }
  } else if (!err && L.first_window) {
    si.first_hw = st.hw; si.had_rewind = 0;
  }
  if (!rewalk) b.sync[i] = sync;
  if (rewalk) {
    b.Enew[i] = st;
    return;
  }
  b.exitA[i] = st;
  b.E[i] = st;
  b.first_hw[i] = si.first_hw;
  b.cflags[i] = (i == 0 || si.had_rewind) ? CF_HAD_REWIND : 0;
}

DATOK_HD void chunk_rewalk(const DeviceModel& m, const WalkBuffers& b, uint32_t i);

// K2c through the fast path: clears the chunk's bits from the re-walk position on, then walks.
DATOK_HD void chunk_rewalk_fast(const DeviceModel& m, const WalkBuffers& b, const FastTables& FT, uint32_t i,
                                uint8_t* seg_cls, uint32_t stage_saddr = 0) {
  const uint32_t lo = i * b.chunk, hi = lo + b.chunk;
  const WState Y = b.Ytmp[i];
  const uint32_t from = (b.sync[i] == K_NOPOS) ? lo : Y.pos;
  clear_chunk_bits(b, from, hi);
  if (Y.pos >= hi || (Y.flags & (WS_INVALID | WS_DONE))) {  // nothing left to walk in this chunk
    chunk_rewalk(m, b, i);
    return;
  }
  chunk_spec_fast(m, b, FT, i, 0, seg_cls, &Y, stage_saddr);
  b.cflags[i] |= CF_OVERWRITTEN;
}

// K2b: returns true if chunk i must be re-walked (state in Ytmp[i]); otherwise Enew[i] is set.
// first_round: the head [lo, sync) has not been walked yet -- the speculative walk left its bits zero,
// so there is nothing to clear.
DATOK_HD bool chunk_stitch(const DeviceModel& m, const WalkBuffers& b, uint32_t i, bool first_round = false) {
  const WState X = b.E[i - 1];
  if (X.flags & (WS_INVALID | WS_DONE)) {  // predecessor not (yet) usable: leave the chunk as it is
    b.Enew[i] = b.E[i];
    return false;
  }
  const uint32_t lo = i * b.chunk;
  const uint32_t s = b.sync[i];
  if (s == K_NOPOS) { b.Ytmp[i] = X; return true; }
  const WalkCtx c = make_walk_ctx(m, b);
  if (!first_round) clear_chunk_bits(b, lo, s);
  WState Y = X;
  SpecInfo si;
  const uint32_t err = walk_run_inl<false, true, false>(c, Y, s, &si);
  if (err) { b.Enew[i] = Y; return false; }
  const WState A = b.exitA[i];
  // (a pending hard-fail END bit at s does not disturb the guess: it is set below)
  const bool match = Y.pos == s && Y.t == m.start && Y.tstart == s && (Y.flags & ~WS_PEND) == 0 &&
                     !(b.cflags[i] & CF_OVERWRITTEN);
  if (!match) { b.Ytmp[i] = Y; return true; }  // the re-walk clears from Y.pos on, then sets a pending END
  if (Y.flags & WS_PEND) set_bit(b.b_end, s);
  // the guess was right: the speculative trace stands.  Only the buffer-window
  // accounting of its first window has to be redone with the true window base.
  WState R = A;
  if (!(A.flags & WS_INVALID)) {
    const uint32_t fh = b.first_hw[i] > Y.hw ? b.first_hw[i] : Y.hw;
    if (b.cflags[i] & CF_HAD_REWIND) {
      if (window_overflow(c, Y.base, fh)) R = wstate_invalid(E_OVERFLOW);
    } else {
      R.base = Y.base;
      R.hw = fh;
      if ((A.flags & WS_DONE) && window_overflow(c, R.base, R.hw)) R = wstate_invalid(E_OVERFLOW);
    }
  }
  b.Enew[i] = R;
  return false;
}

// K2c: re-walk chunk i from Ytmp[i] to the end of the chunk.
DATOK_HD void chunk_rewalk(const DeviceModel& m, const WalkBuffers& b, uint32_t i) {
  const WalkCtx c = make_walk_ctx(m, b);
  const uint32_t lo = i * b.chunk, hi = lo + b.chunk;
  WState Y = b.Ytmp[i];
  const uint32_t from = (b.sync[i] == K_NOPOS) ? lo : Y.pos;
  clear_chunk_bits(b, from, hi);
  SpecInfo si;
  walk_run<false, true, false>(c, Y, hi, &si);
  b.Enew[i] = Y;
  b.cflags[i] |= CF_OVERWRITTEN;
}

// K2d: returns true if the exit state of chunk i changed (successor must be redone).
DATOK_HD bool chunk_commit(const WalkBuffers& b, uint32_t i) {
  const WState n = b.Enew[i];
  if (wstate_equal(n, b.E[i])) return false;
  b.E[i] = n;
  return true;
}

}  // namespace datok
