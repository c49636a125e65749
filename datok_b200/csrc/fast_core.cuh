// fast_core.cuh -- the fast path of the speculative chunk walk (K2a).
//
// One lane walks one chunk in 32-byte SEGMENTS.  Per segment the lane
//   1. classifies its 32 raw bytes (UTF-8 decode + sigma map, matrix.go:388-435)
//      into a small class buffer (shared memory in the kernel) and produces the
//      rune-start word,
//   2. steps through the classes with ONE fused-table lookup per byte
//      (model.hpp: T3 folds "fail -> epsilon at this position -> re-read",
//      matrix.go:472-497,563-576).  A step only records what the table entry says
//      about its byte in three RAW masks (boundaries before the byte: >= 1, == 2;
//      consumed by a non-token transition) and remembers the latest epsilon point,
//   3. derives the four boundary words of the segment from the raw masks with
//      word-parallel logic (derive_bits): the reference's bufft/"token pending"
//      bookkeeping (matrix.go:563-591) is a carry chain over the segment, evaluated
//      with one 64-bit addition instead of per-byte branches.
// Real backtracks (to an epsilon point recorded at an earlier byte, matrix.go:487-497)
// are handled in place when the point lies in the raw range of the current segment;
// anything else that is rare -- far backtracks, hard fails, the window-limit
// vicinity, chunk hand-off, EOF -- drops to the exact walker walk_run() for the rest
// of the segment.  Both walkers are step-equivalent, so a lane can switch between
// them at any loop top.
#pragma once
#include <stdint.h>
#include "walk_core.cuh"

namespace datok {

constexpr uint32_t SEG = 32;                 // bytes per segment = bits per bitmap word
// fused table T3 (u32) and the compact hot rows (u16): see model.hpp
constexpr uint32_t F3_TGT = 0x7FFFu, F3_NT = 1u << 15, F3_K1 = 1u << 16, F3_K2 = 1u << 17, F3_KANY = F3_K1 | F3_K2,
                   F3_EA = 1u << 18, F3_SLOWMARK = 1u << 31;
constexpr uint32_t F16_TGT = 0x07FFu, F16_NT = 1u << 11, F16_EA = 1u << 12, F16_KANY = 1u << 13, F16_K2 = 1u << 14,
                   F16_FAIL = F16_NT;  // target 0 + this flag: the full table says 0 (failure without epsilon transition)
// number of epsilon steps <-> the two flag bits (KANY, K2) of a compact entry: 0 -> 00, 1 -> 01, 2 -> 11
DATOK_HD uint32_t k_to_bits(uint32_t k) { return k | (k >> 1); }
DATOK_HD uint32_t bits_to_k(uint32_t x) { return x - (x >> 1); }
// eps_rec: [14:0] state at the loop top where the point was recorded, [17:16] epsilon steps taken
// there before the byte was consumed, [19] valid
constexpr uint32_t ER_VALID = 1u << 19;
constexpr uint32_t FAST_WINDOW_GUARD = 960;  // stay exact when the buffer window could reach 1024 runes

struct FastTables {
  const uint16_t* hot16;   // compact rows of states 0..n_hot-1, targets >= n_hot already zeroed, plus one all-zero
                           // row n_hot (shared memory in the kernel)
  const uint32_t* t3;      // the full fused table (global memory)
  uint32_t n_hot;
  uint32_t row16;          // BYTES per compact row
  uint32_t row16_inv;      // device only: floor(2^32 / row16) + 1 (the row of a shared-memory offset without a division)
  uint32_t stride3;        // entries per T3 row
  uint32_t hot_saddr;      // device only: shared-window address of hot16
  const uint8_t* ascii_cls2;  // [256]: 2 * class of the ASCII bytes, capped at stop_cl2; the other bytes map to themselves,
                              // so that the class buffer still holds them when they are decoded (shared memory in the kernel)
  const uint32_t* sync_cls;   // [8]: bit c set iff class c is a sync class (shared memory in the kernel)
  uint32_t eot_rewind;     // 1: an EOT rewinds the buffer (matrix walk); 0: it does not (double-array walk, datok.go:1019-1030)
  uint32_t stop_cl2;       // 2 * (number of classes that have a column in the compact rows): the all-zero column.
                           // Rarer classes are stored as this value, and so is the sentinel behind a walk range.
};

// compact entry of hot state t (< n_hot) for the doubled class cl2 = 2 * class
DATOK_HD uint32_t h16_load(const FastTables& T, uint32_t t, uint32_t cl2) {
#if defined(__CUDA_ARCH__)
  uint32_t e;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(T.hot_saddr + t * T.row16 + cl2));
  return e;
#else
  return *reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(T.hot16) + t * T.row16 + cl2);
#endif
}
DATOK_HD uint32_t t3_load(const FastTables& T, uint32_t t, uint32_t cl) {
#if defined(__CUDA_ARCH__)
  return __ldg(T.t3 + t * T.stride3 + cl);
#else
  return T.t3[t * T.stride3 + cl];
#endif
}
// target of the epsilon transition of state t (0: none)
DATOK_HD uint32_t eps_target(const FastTables& T, uint32_t t) {
  if (t < T.n_hot) {
    const uint32_t h = h16_load(T, t, 2u * K_CLS_EPS);
    if (h) return h;
  }
  return t3_load(T, t, K_CLS_EPS) & F3_TGT;
}

struct FastLane {
  uint32_t pos;
  uint32_t t;              // current state
  uint32_t tstart;         // token start (base + bufft) as of raw_from
  uint32_t base;           // last rewind point known at the START of the current segment
  uint32_t eps_p, eps_rec; // latest epsilon point: position, ER_* record (0: none).
                           // Checked lazily: a later boundary or EOT kills it (eps_alive)
  uint32_t hw_med;         // furthest failing position seen by an in-place backtrack (never reset:
                           // an over-estimate from an older window cannot fake an overflow, see to_exact)
  uint32_t first_hw;       // SpecInfo of a guessed start: hw at the first rewind
  uint32_t first_window;   // 1 until the first rewind of a guessed start
  uint32_t stale_end;      // bufft is stale (> buffc) below this position (SentenceEnd backtrack over skipped runes)
  uint32_t raw_from;       // the raw masks cover [raw_from, pos)
  uint32_t u_in;           // 1: tstart == raw_from at raw_from (nothing pending there)
};

// what the steps of the current segment recorded, one bit per byte position
struct RawBits {
  uint32_t c1, c2;         // >= 1 / == 2 epsilon steps in the lookup that consumed the byte
  uint32_t cb;             // an in-place backtrack added one boundary before the byte
  uint32_t nt;             // consumed by a non-token transition (or inside a stale-bufft zone)
};

struct SegBits {
  uint32_t end, skip, sent, tend;
};

enum { FAST_OK = 0, FAST_SLOW = 1, FAST_SLOW_FAIL = 2 };  // _FAIL: a lookup failed and the backtrack is not for the fast path

#define DATOK_UNLIKELY(x) __builtin_expect(!!(x), 0)
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
#define DATOK_STAT(x) ((x)++)
#else
#define DATOK_STAT(x) ((void)0)
#endif

// Boundary words of the positions [ro, po) of a segment from the raw masks.  With
//   n1 = c1 | cb   (at least one boundary before the byte),  n2 = c2 | (c1 & cb)  (two),
//   u[p] = "bufft == buffc at the loop top of p" obeys  u[p+1] = ((u[p] | n1[p]) & nt[p]) | eot[p]
// (matrix.go:584-588: a leading non-token rune moves bufft along; :565-572,:601-603: Token and
// EOT rewind), i.e. a carry chain with generate (n1 & nt) | eot and propagate nt.
//   END  = n1 & ~u          first boundary with something pending: Token      (:565-572)
//   SENT = (n1 & u) | n2    boundary with nothing pending: SentenceEnd        (:573-576)
//   SKIP = (u | n1) & nt    leading non-token rune                            (:584-588)
// deg: two SentenceEnds at one position (not representable, walk_run reports it the same way).
struct Derived {
  uint32_t u;              // bits [ro, po)
  uint32_t n1;
  uint32_t end, sent, skip, tend, deg;
  uint32_t u_at_pos;       // u at position po
};
// eotm: the EOTs that rewind the buffer (all of them in the matrix walk, none in the double-array walk, where an
// EOT is an ordinary skipped rune); tendm: all EOTs -- every consumed one fires TextEnd.
DATOK_HD Derived derive_bits(const RawBits& R, uint32_t eotm, uint32_t ro, uint32_t po, uint32_t u_in, uint32_t tendm) {
  Derived D;
  const uint32_t lim = mask_below(po) & mask_from(ro);
  const uint32_t eot = eotm & lim;
  const uint32_t n1 = R.c1 | R.cb, n2 = R.c2 | (R.c1 & R.cb);
  const uint32_t x = R.nt | eot, g = (n1 & R.nt) | eot;
  const unsigned long long s = (unsigned long long)x + g + ((unsigned long long)(u_in & 1u) << ro);
  const uint32_t carries = (uint32_t)s ^ x ^ g;
  D.u_at_pos = po < 32 ? ((carries >> po) & 1u) : (uint32_t)(s >> 32);
  D.u = carries & lim;
  D.n1 = n1;
  D.end = n1 & ~D.u;
  D.sent = (n1 & D.u) | n2;
  D.skip = (D.u | n1) & R.nt;
  D.tend = tendm & lim;
  D.deg = (n2 & D.u) | (R.c2 & R.cb);
  return D;
}

// token start at position po given the derived words (L.tstart if it was not moved in [ro, po))
DATOK_HD uint32_t derived_tstart(const FastLane& L, const RawBits& R, const Derived& D, uint32_t seg_start, uint32_t po) {
  uint32_t ts;
  if (D.u_at_pos) ts = seg_start + po;
  else {
    const uint32_t m = (D.u | D.n1) & ~R.nt;  // bufft was set here and the byte was not skipped
    ts = m ? seg_start + 31u - clz32(m) : L.tstart;
  }
  if (seg_start + po < L.stale_end) ts = L.stale_end;
  return ts;
}

// is the recorded epsilon point still the reference's epsilonState?  A Token/SentenceEnd
// boundary after it (matrix.go:608-627 clears it on rewind; a SentenceEnd consumes it) or a
// consumed EOT at or after it (:601-603) kills it.
DATOK_HD bool eps_alive(const FastLane& L, const RawBits& R, uint32_t eotm, uint32_t seg_start, uint32_t po) {
  if (!L.eps_rec) return false;
  uint32_t above = 0xFFFFFFFFu, from = 0xFFFFFFFFu;
  if (L.eps_p >= seg_start) {
    const uint32_t qo = L.eps_p - seg_start;
    from = mask_from(qo);
    above = mask_from(qo + 1);
  }
  const uint32_t consumed = mask_below(po) & mask_from(L.raw_from > seg_start ? L.raw_from - seg_start : 0);
  return (((R.c1 | R.cb) & above) | (eotm & from & consumed)) == 0;
}

// last rewind point, taking the boundary bits of the current segment into account
DATOK_HD uint32_t lane_base(const FastLane& L, const SegBits& B, uint32_t seg_start, uint32_t eot_rewind) {
  uint32_t base = L.base;
  if (B.end) { const uint32_t p = seg_start + 31u - clz32(B.end); if (p > base) base = p; }
  if (eot_rewind && B.tend) { const uint32_t p = seg_start + 32u - clz32(B.tend); if (p > base) base = p; }
  return base;
}
// a guessed start's first window closes with the first END/TEND bit of the lane
DATOK_HD void lane_note_first_rewind(FastLane& L, const SegBits& B, uint32_t seg_start, uint32_t eot_rewind) {
  const uint32_t m = B.end | (eot_rewind ? B.tend : 0u);
  if (L.first_window && m) { L.first_hw = seg_start + ctz32(m); L.first_window = 0; }
}

// exact state -> fast lane.  Requires can_go_fast(st).
DATOK_HD bool can_go_fast(const WState& st) {
  return (st.flags & ~WS_PEND) == 0 && st.tstart <= st.pos;
}
DATOK_HD void raw_clear(RawBits& R) { R.c1 = R.c2 = R.cb = R.nt = 0; }
DATOK_HD void to_fast(const WState& st, FastLane& L, RawBits& R) {
  L.pos = st.pos; L.tstart = st.tstart; L.base = st.base; L.t = st.t;
  L.eps_p = st.eps_pos;
  L.eps_rec = st.eps_state ? (ER_VALID | st.eps_state) : 0;  // k = 0: the state itself has the transition
  L.hw_med = st.hw;
  L.stale_end = 0;
  L.raw_from = st.pos;
  L.u_in = st.tstart == st.pos ? 1u : 0u;
  raw_clear(R);
}

// Closes the raw range [raw_from, pos) of the segment: merges its boundary words into B and
// moves the lane's token start / unstarted flag to pos.  Returns false on a degenerate event
// sequence.  Afterwards the raw masks are empty and raw_from == pos.
DATOK_HD bool fast_flush(FastLane& L, RawBits& R, SegBits& B, uint32_t eotm, uint32_t tendm, uint32_t seg_start) {
  const uint32_t po = L.pos - seg_start;
  const uint32_t ro = L.raw_from > seg_start ? L.raw_from - seg_start : 0;
  const Derived D = derive_bits(R, eotm, ro, po, L.u_in, tendm);
  const bool alive = eps_alive(L, R, eotm, seg_start, po);
  B.end |= D.end; B.sent |= D.sent; B.skip |= D.skip; B.tend |= D.tend;
  L.tstart = derived_tstart(L, R, D, seg_start, po);
  L.u_in = (L.tstart == L.pos) ? 1u : 0u;
  if (!alive) L.eps_rec = 0;
  L.raw_from = L.pos;
  raw_clear(R);
  return D.deg == 0;
}

// fast lane (flushed: raw range empty) -> exact state
DATOK_HD void to_exact(const FastLane& L, const SegBits& B, uint32_t seg_start, const FastTables& T, WState& st) {
  const uint32_t base = lane_base(L, B, seg_start, T.eot_rewind);
  st.pos = L.pos; st.tstart = L.tstart; st.base = base; st.t = (uint16_t)L.t;
  st.flags = 0;
  uint32_t es = 0;
  if (L.eps_rec) {
    es = L.eps_rec & F3_TGT;
    for (uint32_t k = (L.eps_rec >> 16) & 3u; k; k--) es = eps_target(T, es);
  }
  st.eps_state = (uint16_t)es;
  st.eps_pos = es ? L.eps_p : 0;
  // hw: furthest byte read in the current buffer window.  hw_med may stem from an older window;
  // it was a real read then, [base, hw_med] is contained in that older window, so counting its
  // runes can never exceed what the reference itself had buffered.
  uint32_t hw = base;
  if (L.hw_med > hw) hw = L.hw_med;
  if (L.pos > base && hw < L.pos - 1) hw = L.pos - 1;
  st.hw = hw;
}

// In-place backtrack to the epsilon point recorded in the raw range of this segment
// (matrix.go:487-497), or FAST_SLOW (nothing changed) when the exact walker has to take over.
// Bprev: boundary words of the segment's positions before raw_from.
#if defined(DATOK_NI_BACKTRACK)
DATOK_HD_SLOW
#else
DATOK_HD
#endif
int fast_backtrack(FastLane& L, RawBits& R, const FastTables& T, uint32_t seg_start, uint32_t eotm,
                            const SegBits& Bprev) {
  if (!L.eps_rec) { DATOK_STAT(g_bt_hard); return FAST_SLOW; }                                 // hard fail
  if (L.eps_p < L.raw_from || L.eps_p < seg_start) { DATOK_STAT(g_bt_far); return FAST_SLOW; }  // far backtrack
  const uint32_t po = L.pos - seg_start, qo = L.eps_p - seg_start, qb = 1u << qo;
  if (!eps_alive(L, R, eotm, seg_start, po)) { DATOK_STAT(g_bt_dead); return FAST_SLOW; }  // the point is dead: hard fail
  if ((R.c2 | R.cb) & qb) { DATOK_STAT(g_bt_third); return FAST_SLOW; }                   // would be a third boundary at q
  // epsilon transition(s) from the recorded loop-top state
  uint32_t es = L.eps_rec & F3_TGT;
  for (uint32_t k = (L.eps_rec >> 16) & 3u; k; k--) es = eps_target(T, es);
  const uint32_t tgt = eps_target(T, es);
  if (L.hw_med < L.pos) L.hw_med = L.pos;
  uint32_t zone = 0;
  if (DATOK_UNLIKELY(L.first_window || (R.nt & qb))) {
    const uint32_t ro = L.raw_from > seg_start ? L.raw_from - seg_start : 0;
    const Derived D = derive_bits(R, eotm, ro, po, L.u_in, eotm);
    if (!((D.u | D.n1) & qb) && L.first_window) {  // Token + rewind (:565-572): closes a guessed start's first window
      const uint32_t m = ((D.end | D.tend) & (qb - 1u)) | Bprev.end | Bprev.tend;
      L.first_hw = m ? seg_start + ctz32(m) : L.hw_med;
      L.first_window = 0;
    }
    // bytes from q on that the first pass skipped as leading non-token runes stay skipped: bufft is
    // not reset by a SentenceEnd (:573-576), it is stale until buffc catches up with it
    if (D.skip & qb) {
      const uint32_t s = D.skip + qb;       // the carry runs through the skip run that starts at q
      zone = D.skip & ~s & mask_from(qo);
      L.stale_end = seg_start + qo + popc32(zone);
      DATOK_STAT(g_zone);
    }
  }
  const uint32_t keep = (qb << 1) - 1u, below = qb - 1u;   // positions <= q, < q
  // the byte at q is read again: its own lookup bits are cleared, the boundary taken here is kept in cb.
  // (c1 at q set: that lookup's epsilon step came first; the chain bound of model.cpp rules out more.)
  R.c2 &= below;
  R.c1 &= keep;
  R.cb = (R.cb & below) | qb;
  R.nt = (R.nt & below) | zone;
  L.pos = L.eps_p;
  L.eps_rec = 0;
  L.t = tgt;
  DATOK_STAT(g_bt_ok);
  return FAST_OK;
}

// Loop-top iterations of the reference from L.pos up to `limit` (<= seg_start + 32) over the
// DOUBLED classes seg_cls[0..31] of the segment.  On FAST_SLOW nothing has been changed by the
// failing iteration and the exact walker must take over at L.pos.
//
// One loop for all lanes of a warp: a lane that needs the rare path (cold state, rare class, failure)
// takes it inside the iteration and rejoins the others at the end of the same iteration (the loop's end
// test is what makes the compiler reconverge the lanes there: with an endless loop it turns the hot path
// into a tight loop of its own that a lane leaves at its first rare event and re-enters only when every
// other lane has left it too -- twice the iterations per segment).
// In the loop the latest epsilon point is (eps_bit, eps_a | eps_rec): eps_bit = the position's bit in this
// segment (0: the point is older than the segment, L.eps_p holds it); if the lookup that recorded it went
// through the compact rows, eps_a is the shared-memory address of that entry (state and epsilon steps are
// recovered from it when they are needed: one select in the hot path), else eps_a is 0 and eps_rec holds
// the ER_* record.
struct FastCtx {  // what the rare path needs beyond the tables: the raw input, for the class of a rare rune
  const uint8_t* in;
  uint32_t N;
  const ClsTables* cls;
};
// ER_* record of the epsilon point whose lookup was the compact entry at (row t, doubled class cl2)
DATOK_HD uint32_t er_from_entry(const FastTables& T, uint32_t t, uint32_t cl2) {
  const uint32_t e = h16_load(T, t, cl2);
  return ER_VALID | t | (bits_to_k((e >> 13) & 3u) << 16);
}
// offset of the position whose bit is `bit` (0: the bit has been shifted out behind position 31)
DATOK_HD uint32_t off_of_bit(uint32_t bit) { return bit ? 31u - clz32(bit) : 32u; }
DATOK_HD int fast_run(FastLane& L, RawBits& R, const FastTables& T, const FastCtx& X, const uint8_t* seg_cls,
                      uint32_t seg_start, uint32_t limit, uint32_t eotm, const SegBits& Bprev) {
  if (L.pos >= limit) return FAST_OK;
  const uint32_t lim_off = limit - seg_start;
  uint32_t end_bit = lim_off < 32 ? 1u << lim_off : 0u;
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(end_bit));  // opaque: otherwise recomputed in every iteration of the loop below
#endif
  uint32_t off = L.pos - seg_start;
  uint32_t bit = 1u << off;
  uint32_t t = L.t;
  uint32_t eps_bit = (L.eps_rec && L.eps_p >= seg_start) ? 1u << (L.eps_p - seg_start) : 0u;
  uint32_t eps_rec = L.eps_rec;  // ER_* form; only read while eps_a == 0
  uint32_t eps_a = 0;            // device: address of the recording entry; host: 1 + (row * 65536 + doubled class)
  uint32_t c1 = R.c1, c2 = R.c2, nt = R.nt;
#if defined(__CUDA_ARCH__)
  // On the device the position is carried as (p, bit) only: p = shared-window address of the position's class byte.
  // The offset is recovered from `bit` where the rare paths need it -- the address of the lane's class buffer would
  // have to be re-derived from the thread index each time (the hot loop leaves no register for it).
  uint32_t p = (uint32_t)__cvta_generic_to_shared(seg_cls) + off;
#define DATOK_CUR_OFF() (31u - clz32(bit))  /* inside the loop the bit is never shifted out */
#define DATOK_END_OFF() off_of_bit(bit)
#define DATOK_CUR_CLS(dst) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(dst) : "r"(p))
#define DATOK_MOVE_TO(new_off) do { const uint32_t no_ = (new_off); p += no_ - DATOK_CUR_OFF(); bit = 1u << no_; } while (0)
#define DATOK_MOVE_TO_BIT(nb) do { const uint32_t nb_ = (nb); p += (31u - clz32(nb_)) - DATOK_CUR_OFF(); bit = nb_; } while (0)
#define DATOK_STEP() do { p++; bit += bit; } while (0)
  // hc: shared-window address of the row's column for the byte at p, i.e. T.hot_saddr + its doubled class.  The class
  // byte of the NEXT position is fetched while the current lookup is in flight (it does not depend on the state), so
  // the per-byte dependency chain is one multiply-add, the row lookup and one AND -- not two loads back to back.
  uint32_t hc;
#if defined(DATOK_NO_CLASS_PREFETCH)
#define DATOK_LOAD_HC() ((void)0)
#else
#define DATOK_LOAD_HC() do { uint32_t c_; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c_) : "r"(p)); hc = T.hot_saddr + c_; } while (0)
#endif
  DATOK_LOAD_HC();
  uint32_t row16 = T.row16;
  asm volatile("" : "+r"(row16));  // opaque, like end_bit
#define DATOK_ROW_OF(addr) __umulhi((addr) - T.hot_saddr, T.row16_inv)
#define DATOK_EPS_REC() (eps_a ? er_from_entry(T, DATOK_ROW_OF(eps_a), (eps_a - T.hot_saddr) - DATOK_ROW_OF(eps_a) * T.row16) : eps_rec)
#else
#define DATOK_EPS_REC() (eps_a ? er_from_entry(T, (eps_a - 1u) >> 16, (eps_a - 1u) & 0xFFFFu) : eps_rec)
#define DATOK_CUR_OFF() off
#define DATOK_END_OFF() off
#define DATOK_CUR_CLS(dst) ((dst) = seg_cls[off])
#define DATOK_MOVE_TO(new_off) do { off = (new_off); bit = 1u << off; } while (0)
#define DATOK_MOVE_TO_BIT(nb) do { bit = (nb); off = ctz32(bit); } while (0)
#define DATOK_STEP() do { off++; bit += bit; } while (0)
#endif
  // row of the lookup: the state's own, or the all-zero row n_hot ("see the full table") for a cold state,
  // which the loop top only sees on entry: the rare path below steps until the state is hot again
  // (the hot path only carries tl: while the state is hot, t == tl; a cold state is kept in t_cold)
  uint32_t tl = t < T.n_hot ? t : T.n_hot, t_cold = t;
  do {
    // ---- one compact-row lookup per byte ----
    uint32_t e, a;
#if defined(__CUDA_ARCH__)
#if defined(DATOK_NO_CLASS_PREFETCH)
    {
      uint32_t cl2;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(cl2) : "r"(p));
      a = T.hot_saddr + tl * row16 + cl2;
      asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(a));
    }
#else
    uint32_t cl2n;
    {
      a = tl * row16 + hc;
      asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(a));
      asm volatile("ld.shared.u8 %0, [%1 + 1];" : "=r"(cl2n) : "r"(p));  // (behind the range: a scratch byte, never used)
    }
#endif
#else
    a = 1u + ((tl << 16) | seg_cls[off]);
    e = h16_load(T, tl, seg_cls[off]);
#endif
    if (DATOK_UNLIKELY((e & F16_TGT) == 0)) {
      // ---- rare: a failure (F16_FAIL), or not in the compact rows (0): cold state, rare class, target
      // outside the hot rows, marked entry.  Steps through the full table until the state is hot again ----
      bool failed = e == F16_FAIL;
      {  // the state of this lookup, from the entry's address (tl itself is not kept across the lookup: one move less per byte)
#if defined(__CUDA_ARCH__)
        const uint32_t row = DATOK_ROW_OF(a);
#else
        const uint32_t row = (a - 1u) >> 16;
#endif
        t = row == T.n_hot ? t_cold : row;
      }
      if (failed && eps_a != 0) {
        // The common backtrack, inline: the point was recorded in this raw range by a lookup without
        // epsilon steps of its own, both states are hot, q holds no boundary or skipped rune yet and
        // nothing killed the point.  Everything else goes through fast_backtrack() below.
        // (a recording lookup with epsilon steps of its own left a boundary at q: the `dead` test sees it)
#if defined(__CUDA_ARCH__)
        const uint32_t es = DATOK_ROW_OF(eps_a);
#else
        const uint32_t es = (eps_a - 1u) >> 16;
#endif
        const uint32_t qb = eps_bit;
        const uint32_t ro = L.raw_from > seg_start ? L.raw_from - seg_start : 0;
        if (qb >= (1u << ro) && !L.first_window) {
          const uint32_t below = qb - 1u, keep = below | qb;
          const uint32_t tgt = h16_load(T, es, 2u * K_CLS_EPS);
          const uint32_t cb = R.cb;
          const uint32_t dead = ((c1 | cb) & ~keep) | (eotm & ~below & (bit - 1u)) | ((c1 | c2 | cb | nt) & qb);
          if (tgt != 0 && dead == 0) {
            DATOK_STAT(g_bt_ok); DATOK_STAT(g_bt_inline);
            const uint32_t pos = seg_start + DATOK_CUR_OFF();
            if (L.hw_med < pos) L.hw_med = pos;
            c1 &= below; c2 &= below; nt &= below;
            R.cb = (cb & below) | qb;
            DATOK_MOVE_TO_BIT(qb);
            tl = tgt; eps_rec = 0; eps_a = 0; eps_bit = 0;  // (tgt is a hot state: the compact row says so)
#if defined(__CUDA_ARCH__)
            DATOK_LOAD_HC();
#endif
            continue;
          }
        }
      }
      for (;;) {
        uint32_t e3 = 0;
        if (!failed) {
          DATOK_STAT(g_cold);
          uint32_t cl2;
          DATOK_CUR_CLS(cl2);
          if (cl2 == T.stop_cl2) cl2 = 2u * class_at(X.in, X.N, seg_start + DATOK_CUR_OFF(), *X.cls);  // a rare class: from the raw bytes
          e3 = t3_load(T, t, cl2 >> 1);
        }
        if ((e3 & F3_TGT) == 0) {  // 0: failure without epsilon transition; marked: leave to walk_run
          L.pos = seg_start + DATOK_CUR_OFF(); L.t = t;
          if (eps_bit) L.eps_p = seg_start + ctz32(eps_bit);
          L.eps_rec = DATOK_EPS_REC();
          R.c1 = c1; R.c2 = c2; R.nt = nt;
          if (e3 != 0) { DATOK_STAT(g_mark); return FAST_SLOW; }
          if (fast_backtrack(L, R, T, seg_start, eotm, Bprev) != FAST_OK) return FAST_SLOW_FAIL;
          DATOK_MOVE_TO(L.pos - seg_start);
          t = L.t; eps_rec = 0; eps_a = 0; eps_bit = 0;
          c1 = R.c1; c2 = R.c2; nt = R.nt;
          failed = false;
        } else {
          DATOK_STAT(g_fast);
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
          g_hist[t]++;
#endif
          if (e3 & F3_KANY) c1 |= bit;
          if (e3 & F3_K2) c2 |= bit;
          if (e3 & F3_NT) nt |= bit;
          if (e3 & F3_EA) { eps_bit = bit; eps_a = 0; eps_rec = ER_VALID | t | (e3 & F3_KANY); }
          t = e3 & F3_TGT;
          DATOK_STEP();
        }
        if (t < T.n_hot || bit == end_bit) break;
      }
      tl = t < T.n_hot ? t : T.n_hot;
      t_cold = t;
#if defined(__CUDA_ARCH__)
      DATOK_LOAD_HC();
#endif
      continue;
    }
    DATOK_STAT(g_fast);
#if defined(DATOK_COUNT) && !defined(__CUDA_ARCH__)
    g_hist[tl]++;
#endif
#if defined(__CUDA_ARCH__)
    // (one flag per bit of the entry's high byte: the four tests become one R2P)
    asm("{\n\t.reg .pred pk, p2, pn, pe;\n\t.reg .b32 x;\n\t"
        "and.b32 x, %5, 0x2000;\n\tsetp.ne.u32 pk, x, 0;\n\t"
        "and.b32 x, %5, 0x4000;\n\tsetp.ne.u32 p2, x, 0;\n\t"
        "and.b32 x, %5, 0x0800;\n\tsetp.ne.u32 pn, x, 0;\n\t"
        "and.b32 x, %5, 0x1000;\n\tsetp.ne.u32 pe, x, 0;\n\t"
        "@pk or.b32 %0, %0, %6;\n\t"
        "@p2 or.b32 %1, %1, %6;\n\t"
        "@pn or.b32 %2, %2, %6;\n\t"
        "@pe mov.b32 %3, %6;\n\t"
        "@pe mov.b32 %4, %7;\n\t}"
        : "+r"(c1), "+r"(c2), "+r"(nt), "+r"(eps_bit), "+r"(eps_a)
        : "r"(e), "r"(bit), "r"(a));
    p++;
#if !defined(DATOK_NO_CLASS_PREFETCH)
    hc = T.hot_saddr + cl2n;
#endif
#else
    if (e & F16_KANY) c1 |= bit;
    if (e & F16_K2) c2 |= bit;
    if (e & F16_NT) nt |= bit;
    if (e & F16_EA) { eps_bit = bit; eps_a = a; }
    off++;
#endif
    tl = e & F16_TGT;
    bit += bit;
  } while (bit != end_bit);
  t = tl == T.n_hot ? t_cold : tl;
  L.pos = seg_start + DATOK_END_OFF(); L.t = t;
  if (eps_bit) L.eps_p = seg_start + ctz32(eps_bit);
  L.eps_rec = DATOK_EPS_REC();
#undef DATOK_EPS_REC
#undef DATOK_ROW_OF
#undef DATOK_LOAD_HC
#undef DATOK_CUR_OFF
#undef DATOK_END_OFF
#undef DATOK_CUR_CLS
#undef DATOK_MOVE_TO
#undef DATOK_MOVE_TO_BIT
#undef DATOK_STEP
  R.c1 = c1; R.c2 = c2; R.nt = nt;
  return FAST_OK;
}

// Asynchronous staging of a lane's next segment (device only, build options): the 32 raw bytes go from global
// memory into the lane's slot in shared memory while the lane walks the current segment; the next classification
// finds them there.
//   DATOK_STAGE_ASYNC  through the async copy unit: cp.async (LDGSTS), 2 x 16 bytes, L1 bypassed
//   DATOK_STAGE_TMA    through the TMA engine: one cp.async.bulk (UBLKCP) of 32 bytes per lane and segment, completion
//                      on the lane's own mbarrier (transaction bytes), waited for with try_wait.parity
// Both are measured slower than plain loads + an L1 prefetch (profiles/r2_stage_ab.txt): the slots displace resident
// table rows, and the latency they hide is already hidden by the 32 warps of the CTA.
struct SegStage {
  uint32_t slot_saddr;   // shared-window address of the lane's 32-byte slot (0: no staging, e.g. on the host);
                         // DATOK_STAGE_TMA: the lane's mbarrier sits 32 bytes behind it
  uint32_t staged_for;   // position of the segment the slot holds or is being filled with (K_NOPOS: none)
  uint32_t parity;       // DATOK_STAGE_TMA: phase of the mbarrier that the next wait completes
};
DATOK_HD void stage_init(SegStage& S, uint32_t slot_saddr) {
  S.slot_saddr = slot_saddr; S.staged_for = K_NOPOS; S.parity = 0;
#if defined(__CUDA_ARCH__) && defined(DATOK_STAGE_TMA)
  if (slot_saddr) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(slot_saddr + 32u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#endif
}
DATOK_HD void stage_wait(SegStage& S) {  // the copy in flight (if any) has landed
#if defined(__CUDA_ARCH__)
#if defined(DATOK_STAGE_TMA)
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(S.slot_saddr + 32u), "r"(S.parity) : "memory");
  }
  S.parity ^= 1u;
#else
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
#endif
  S.staged_for = K_NOPOS;
}
DATOK_HD void stage_segment(SegStage& S, const uint8_t* in, uint32_t N, uint32_t seg_start) {
#if defined(__CUDA_ARCH__)
  const uint8_t* p = in + seg_start;
  if (S.slot_saddr && seg_start + SEG <= N && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
#if defined(DATOK_STAGE_TMA)
    // (the slot was last read with ordinary loads: order them before the async-proxy write)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 32;" :: "r"(S.slot_saddr + 32u) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                 :: "r"(S.slot_saddr), "l"(p), "r"(S.slot_saddr + 32u) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.cg.shared.global [%0 + 16], [%1 + 16], 16;"
                 :: "r"(S.slot_saddr), "l"(p) : "memory");
#endif
    S.staged_for = seg_start;
  }
#else
  (void)S; (void)in; (void)N; (void)seg_start;
#endif
}

// loads the 32 raw bytes at `p` (zero padded beyond N) as 8 little-endian words
DATOK_HD void load_segment_words(const uint8_t* in, uint32_t N, uint32_t seg_start, uint32_t* words, SegStage* S = nullptr) {
  const uint8_t* p = in + seg_start;
#if defined(__CUDA_ARCH__)
  if (S && S->staged_for == seg_start) {  // staged while the previous segment was walked
    stage_wait(*S);
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(words[0]), "=r"(words[1]), "=r"(words[2]), "=r"(words[3]) : "r"(S->slot_saddr) : "memory");
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4 + 16];" : "=r"(words[4]), "=r"(words[5]), "=r"(words[6]), "=r"(words[7]) : "r"(S->slot_saddr) : "memory");
    return;
  }
  if (S && S->staged_for != K_NOPOS) stage_wait(*S);  // a copy for another segment is in flight (the lane went elsewhere): let it land
#else
  (void)S;
#endif
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

// DOUBLED classes (2 * class: the byte offset into a compact table row) and rune starts of the
// 32 bytes at seg_start (bytes >= N: no rune start, class unspecified).  seg_cls must be 4-byte
// aligned and hold 33 bytes (the caller puts the walk's stop value behind the range).  ASCII bytes go through the LUT `ascii_cls2` (2 * class per ASCII byte) four at a
// time; the few other bytes are decoded afterwards, one rune each.
// *eot_word: positions holding the byte 0x04 (matrix.go:13,422).
// Classes without a column in the compact rows are stored as stop_cl2 (the all-zero column).
DATOK_HD uint32_t cap_cl2(uint32_t cl, uint32_t stop_cl2) { const uint32_t c2 = 2u * cl; return c2 < stop_cl2 ? c2 : stop_cl2; }
DATOK_HD void classify_segment(const uint8_t* in, uint32_t N, uint32_t seg_start, const ClsTables& T,
                               const uint8_t* ascii_cls2, uint32_t stop_cl2, uint8_t* seg_cls, uint32_t* rstart_word,
                               uint32_t* eot_word, bool* any_invalid, uint32_t* nonascii_word = nullptr,
                               SegStage* stage = nullptr) {
  uint32_t words[8];
  load_segment_words(in, N, seg_start, words, stage);
  uint32_t* out = reinterpret_cast<uint32_t*>(seg_cls);
  uint32_t nonascii = 0, eot_any = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 8; k++) {
    const uint32_t v = words[k];
    const uint32_t c0 = ascii_cls2[v & 0xFFu], c1 = ascii_cls2[(v >> 8) & 0xFFu];
    const uint32_t c2 = ascii_cls2[(v >> 16) & 0xFFu], c3 = ascii_cls2[v >> 24];
    out[k] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);  // the LUT already holds 2 * class
    const uint32_t h = v & 0x80808080u;
    nonascii |= (((h >> 7) | (h >> 14) | (h >> 21) | (h >> 28)) & 0xFu) << (4 * k);
    const uint32_t x = v ^ 0x04040404u;
    eot_any |= (x - 0x01010101u) & ~x & 0x80808080u;  // some byte of v is 0x04 (exact as an "any" test)
  }
  const uint32_t valid = (seg_start + SEG <= N) ? 0xFFFFFFFFu : mask_below(N > seg_start ? N - seg_start : 0);
  uint32_t eot = 0;
  if (DATOK_UNLIKELY(eot_any != 0)) {
    for (int k = 0; k < 8; k++)
      for (int j = 0; j < 4; j++)
        if (((words[k] >> (8 * j)) & 0xFFu) == 0x04u) eot |= 1u << (4 * k + j);
  }
  *eot_word = eot & valid;
  if (nonascii_word) *nonascii_word = nonascii | ~valid;
  uint32_t rs = ~nonascii & valid;
  uint32_t m = nonascii & valid;
  while (m) {
    const uint32_t j = ctz32(m);
    m &= m - 1;
    const uint32_t p = seg_start + j;
    {
      // the common case first: a well-formed two-byte rune below U+0100 (lead C2/C3: the Latin-1 letters).
      // (The LUT left the bytes >= 0x80 in the class buffer as they are: no second trip to global memory.)
      const uint32_t b0 = seg_cls[j];
      if ((b0 & 0xFEu) == 0xC2u && p + 1 < N) {
        // (an ASCII neighbour has been replaced by its class: it is no continuation byte whatever that value looks like)
        const uint32_t b1 = j + 1 < SEG ? (((nonascii >> (j + 1)) & 1u) ? seg_cls[j + 1] : 0u) : in[p + 1];
        if ((b1 & 0xC0u) == 0x80u) {
          seg_cls[j] = (uint8_t)cap_cl2(T.latin1_cls[((b0 & 1u) << 6) | (b1 & 0x3Fu)], stop_cl2);
          rs |= 1u << j;
          if (j + 1 < SEG) {
            seg_cls[j + 1] = (uint8_t)(2u * K_CLS_CONT);
            m &= ~(1u << (j + 1));
          }
          continue;
        }
      }
    }
    bool st, inv;
    const uint32_t cl = classify_pos(in, N, p, T, &st, &inv);
    seg_cls[j] = (uint8_t)cap_cl2(cl, stop_cl2);
    if (st) rs |= 1u << j;
    if (inv) *any_invalid = true;
    if (st && !inv) {
      // a well-formed multi-byte rune: its continuation bytes need no decoding of their own
      const uint32_t b0 = in[p];
      const uint32_t wdt = b0 >= 0xF0 ? 4u : b0 >= 0xE0 ? 3u : 2u;
      for (uint32_t q = 1; q < wdt && j + q < SEG; q++) {
        seg_cls[j + q] = (uint8_t)(2u * K_CLS_CONT);
        m &= ~(1u << (j + q));
      }
    }
  }
  *rstart_word = rs;
}

// First sync point p in (seg_start + from_off, seg_start + to_off] of a classified segment: the byte before p is an
// ASCII byte of a sync class (one the root state skips).  nonascii: the segment's bytes >= 0x80.  K_NOPOS if none.
// (Classes without a column in the compact rows are stored as stop_cl2 and never count: a later point does as well.)
DATOK_HD uint32_t find_sync_cls(const uint8_t* seg_cls, uint32_t nonascii, const uint32_t* sync_cls, uint32_t stop_cl2,
                                uint32_t seg_start, uint32_t from_off, uint32_t to_off) {
  for (uint32_t j = from_off; j < to_off; j++) {
    const uint32_t c2 = seg_cls[j], c = c2 >> 1;
    if (!((nonascii >> j) & 1u) && c2 != stop_cl2 && ((sync_cls[c >> 5] >> (c & 31)) & 1u)) return seg_start + j + 1;
  }
  return K_NOPOS;
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
