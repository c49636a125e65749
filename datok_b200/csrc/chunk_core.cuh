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
#include "compact_core.cuh"
#include "walk_core.cuh"

namespace datok {

constexpr uint32_t CF_HAD_REWIND = 1;   // the speculative walk closed at least one buffer window
constexpr uint32_t CF_OVERWRITTEN = 2;  // the speculative trace has been replaced by a re-walk

struct DeviceModel {
  const uint16_t* table;
  uint32_t row_shift, start, eps_lo, n_classes;
  ClsTables cls;           // pointers into device memory
  uint32_t sync_mask[8];
};

struct WalkBuffers {
  const uint8_t* in;
  uint8_t* cls;            // N + pad
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
};

DATOK_HD WalkCtx make_walk_ctx(const DeviceModel& m, const WalkBuffers& b) {
  WalkCtx c;
  c.table = m.table; c.row_shift = m.row_shift; c.start = m.start; c.eps_lo = m.eps_lo;
  c.cls = b.cls; c.N = b.N; c.rstart = b.rstart;
  c.b_end = b.b_end; c.b_skip = b.b_skip; c.b_sent = b.b_sent; c.b_tend = b.b_tend;
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
    walk_run<false>(c, st, hi, &si);
    b.exitA[0] = st;
    b.E[0] = st;
    b.first_hw[0] = 0;
    b.cflags[0] = CF_HAD_REWIND;
    return;
  }
  uint32_t s = K_NOPOS;
  const uint32_t lim = hi < b.N ? hi : b.N;
  for (uint32_t p = lo; p < lim; p++) {
    if (sync_class(m.sync_mask, b.cls[p - 1])) { s = p; break; }
  }
  b.sync[i] = s;
  if (s == K_NOPOS) {  // no sync point: the predecessor's state has to be walked through
    st = wstate_invalid(0);
  } else {
    st.pos = st.tstart = st.base = st.hw = s;
    st.t = (uint16_t)m.start;
    walk_run<true>(c, st, hi, &si);
  }
  b.exitA[i] = st;
  b.E[i] = st;
  b.first_hw[i] = si.first_hw;
  b.cflags[i] = si.had_rewind ? CF_HAD_REWIND : 0;
}

// K2b: returns true if chunk i must be re-walked (state in Ytmp[i]); otherwise Enew[i] is set.
DATOK_HD bool chunk_stitch(const DeviceModel& m, const WalkBuffers& b, uint32_t i) {
  const WState X = b.E[i - 1];
  if (X.flags & (WS_INVALID | WS_DONE)) {  // predecessor not (yet) usable: leave the chunk as it is
    b.Enew[i] = b.E[i];
    return false;
  }
  const uint32_t lo = i * b.chunk;
  const uint32_t s = b.sync[i];
  if (s == K_NOPOS) { b.Ytmp[i] = X; return true; }
  const WalkCtx c = make_walk_ctx(m, b);
  clear_chunk_bits(b, lo, s);
  WState Y = X;
  SpecInfo si;
  const uint32_t err = walk_run<false>(c, Y, s, &si);
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
      if (window_overflow(b.rstart, Y.base, fh)) R = wstate_invalid(E_OVERFLOW);
    } else {
      R.base = Y.base;
      R.hw = fh;
      if ((A.flags & WS_DONE) && window_overflow(b.rstart, R.base, R.hw)) R = wstate_invalid(E_OVERFLOW);
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
  walk_run<false>(c, Y, hi, &si);
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
