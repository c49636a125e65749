// emul.cpp -- TEST INFRASTRUCTURE: runs the kernel bodies of datok_b200/csrc/*.cuh
// (classify_pos, chunk_spec/stitch/rewalk/commit, process_word, agg_combine,
// finalize_stream) sequentially on the CPU, in the same grid/block/thread
// decomposition the CUDA kernels use.  It lets the CPU test-suite check the
// speculative-chunk algorithm against the oracle without a GPU.  It is NOT part
// of the product library (libdatok_b200.so does not contain it) and is never
// used as a fallback.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../datok_b200/csrc/chunk_core.cuh"
#include "../../datok_b200/csrc/format_core.cuh"
#include "../../datok_b200/csrc/model.hpp"

using namespace datok;

extern "C" {

struct EmulResult {
  int status;
  uint64_t n_tokens, n_sentences, n_texts, n_sent_pos, n_runes;
  uint32_t* tok_bytes;
  int32_t* tok_pos;
  int32_t* sent_pos;
  uint32_t* sent_tok;
  uint32_t *text_tok_end, *text_sent_end, *text_sentpos_end, *text_byte_end;
  uint32_t carry_state, has_invalid;
  uint32_t rounds, n_rewalks, n_stitch_mismatch;
  uint16_t* tok_delta;
  uint8_t* tok_delta8;
  uint32_t* esc;
  uint32_t n_esc;
  uint32_t eot_rewind;  // 0: a double-array model -- the delta-coded forms (cursors restart at every text) do not apply
  uint8_t* text;        // the device formatter's bodies (format_core.cuh) over the arrays above; null with malformed UTF-8
  uint64_t text_len;
  uint32_t delta_range;  // 1: a delta of the compact forms does not fit (DATOK_ERR_COMPACT_RANGE): the absolute arrays stand alone
};

struct EmulModel {
  HostModel hm;
  DeviceModel dm;
};

static uint8_t g_ascii_cls2[256];
// the kernel prologue's job: n_hot compact rows with non-resident targets zeroed
static void make_fast_tables(const HostModel& hm, const DeviceModel& m, uint32_t n_hot, std::vector<uint16_t>& hot,
                             FastTables& FT) {
  if (n_hot > hm.hot16_rows) n_hot = hm.hot16_rows;
  hot.assign(hm.hot16.begin(), hm.hot16.begin() + (size_t)n_hot * hm.stride16);
  for (auto& e : hot) if ((e & F16_TGT) >= n_hot) e = 0;
  hot.resize(hot.size() + hm.stride16, 0);  // the all-zero row n_hot
  for (int i = 0; i < 128; i++) { g_ascii_cls2[i] = (uint8_t)cap_cl2(hm.ascii_cls[i], 2 * hm.hot_cols); g_ascii_cls2[128 + i] = (uint8_t)(128 + i); }
  FT.hot16 = hot.data(); FT.t3 = m.table2; FT.n_hot = n_hot; FT.row16 = hm.stride16 * 2u; FT.stride3 = m.stride2;
  FT.hot_saddr = 0; FT.ascii_cls2 = g_ascii_cls2; FT.stop_cl2 = 2 * hm.hot_cols; FT.sync_cls = hm.sync_mask; FT.eot_rewind = hm.eot_rewind ? 1u : 0u;
}

EmulModel* emul_load(const char* path, int* err) {
  EmulModel* m = new EmulModel();
  std::string why;
  const size_t pl = std::strlen(path);
  const bool foma = pl > 4 && !std::strcmp(path + pl - 4, ".fst");  // the compile path: LoadFomaFile(path).ToMatrix()
  int rc = foma ? load_foma_file(path, m->hm, why) : load_matok_file(path, m->hm, why);
  if (rc) { *err = rc; delete m; return nullptr; }
  HostModel& h = m->hm;
  m->dm.table = h.table.data();
  m->dm.table2 = h.table2.data();
  m->dm.hot16 = h.hot16.data();
  m->dm.row_shift = h.row_shift; m->dm.start = h.start; m->dm.n_classes = h.n_classes; m->dm.stride2 = h.stride2;
  m->dm.stride16 = h.stride16; m->dm.hot16_rows = h.hot16_rows; m->dm.hot_cols = h.hot_cols; m->dm.eot_rewind = h.eot_rewind ? 1u : 0u;
  m->dm.cls.ascii_cls = h.ascii_cls; m->dm.cls.latin1_cls = h.latin1_cls;
  m->dm.cls.rune_key = h.rune_key.data(); m->dm.cls.rune_cls = h.rune_cls.data();
  m->dm.cls.n_rune = (uint32_t)h.rune_key.size(); m->dm.cls.identity_cls = h.identity_cls; m->dm.cls.self = nullptr;
  std::memcpy(m->dm.sync_ascii, h.sync_ascii, sizeof h.sync_ascii);
  std::memcpy(m->dm.sync_cls, h.sync_mask, sizeof h.sync_mask);
  *err = 0;
  return m;
}
// re-lays the model out like the product's one-time calibration (api.cu calibrate_locked): class ids in
// order of their frequency in `data`, compact rows with the frequent classes only (force_cols != 0: that many)
int emul_calibrate(EmulModel* m, const uint8_t* data, uint32_t n, uint32_t force_cols) {
  HostModel& h = m->hm;
  std::vector<uint64_t> cls_hist(256, 0);
  for (uint32_t p = 0; p < n; p++) cls_hist[h.cls_base[class_at(data, n, p, m->dm.cls)]]++;
  h.force_hot_cols = force_cols;
  std::string why;
  int rc = build_layout(h, why, nullptr, n ? cls_hist.data() : nullptr);
  if (rc) return rc;
  m->dm.table = h.table.data();
  m->dm.table2 = h.table2.data();
  m->dm.hot16 = h.hot16.data();
  m->dm.row_shift = h.row_shift; m->dm.start = h.start; m->dm.n_classes = h.n_classes; m->dm.stride2 = h.stride2;
  m->dm.stride16 = h.stride16; m->dm.hot16_rows = h.hot16_rows; m->dm.hot_cols = h.hot_cols; m->dm.eot_rewind = h.eot_rewind ? 1u : 0u;
  m->dm.cls.ascii_cls = h.ascii_cls; m->dm.cls.latin1_cls = h.latin1_cls;
  m->dm.cls.rune_key = h.rune_key.data(); m->dm.cls.rune_cls = h.rune_cls.data();
  m->dm.cls.n_rune = (uint32_t)h.rune_key.size(); m->dm.cls.identity_cls = h.identity_cls; m->dm.cls.self = nullptr;
  std::memcpy(m->dm.sync_ascii, h.sync_ascii, sizeof h.sync_ascii);
  std::memcpy(m->dm.sync_cls, h.sync_mask, sizeof h.sync_mask);
  return 0;
}
uint32_t emul_hot_cols(EmulModel* m) { return m->hm.hot_cols; }
void emul_free(EmulModel* m) { delete m; }
uint32_t emul_n_classes(EmulModel* m) { return m->hm.n_classes; }
uint32_t emul_new_state(EmulModel* m, uint32_t old_state) { return m->hm.new_of_old[old_state]; }
uint32_t emul_old_state(EmulModel* m, uint32_t new_state) { return m->hm.old_of_new[new_state]; }

// order: 0 = ascending thread order, 1 = descending (results must not depend on it)
// mode: 0 = exact walker only (chunk_spec), n > 0 = fused fast path with n hot rows (chunk_spec_fast)
EmulResult* emul_transduce(EmulModel* em, const uint8_t* in, uint32_t N, uint32_t flags, uint32_t chunk,
                           uint32_t carry_state, int sentence_end_in, int text_end_in, int order, int mode) {
  const DeviceModel& m = em->dm;
  EmulResult* R = (EmulResult*)std::calloc(1, sizeof(EmulResult));
  WalkBuffers b;
  std::memset(&b, 0, sizeof b);
  b.in = in; b.N = N; b.chunk = chunk;
  b.final_input = (flags & 512u) ? 0u : 1u;
  b.n_chunks = N / chunk + 1;
  b.n_words = b.n_chunks * (chunk / 32);
  uint32_t counters[8] = {0};
  b.counters = counters;
  std::vector<uint32_t> rstart(b.n_words, 0), bend(b.n_words, 0), bskip(b.n_words, 0), bsent(b.n_words, 0),
      btend(b.n_words, 0);
  std::vector<WState> E(b.n_chunks), exitA(b.n_chunks), Enew(b.n_chunks), Ytmp(b.n_chunks);
  std::vector<uint32_t> sync(b.n_chunks), first_hw(b.n_chunks), cflags(b.n_chunks);
  unsigned long long err_key = ~0ull;
  b.rstart = rstart.data(); b.b_end = bend.data(); b.b_skip = bskip.data();
  b.b_sent = bsent.data(); b.b_tend = btend.data();
  b.E = E.data(); b.exitA = exitA.data(); b.Enew = Enew.data(); b.Ytmp = Ytmp.data();
  b.sync = sync.data(); b.first_hw = first_hw.data(); b.cflags = cflags.data();
  b.err_key = &err_key;

  const uint32_t start_state = carry_state ? em->hm.new_of_old[carry_state] : m.start;
  if (mode == 0) {
    // K1 (rune starts) + K2a exact speculative walk
    for (uint32_t p = 0; p < N; p++) {
      bool st, inv;
      classify_pos(in, N, p, m.cls, &st, &inv);
      if (st) rstart[p >> 5] |= 1u << (p & 31);
      if (inv) counters[2] |= 1;
    }
    for (uint32_t k = 0; k < b.n_chunks; k++) chunk_spec(m, b, order ? b.n_chunks - 1 - k : k, start_state);
  } else {
    // K1+K2a fused fast path (it has to write every boundary word itself: start from garbage)
    for (auto* v : {&bend, &bskip, &bsent, &btend}) std::fill(v->begin(), v->end(), 0xDEADBEEFu);
    FastTables FT;
    std::vector<uint16_t> hot;
    make_fast_tables(em->hm, m, (uint32_t)mode, hot, FT);
    uint8_t seg_cls[36];
    for (uint32_t k = 0; k < b.n_chunks; k++)
      chunk_spec_fast(m, b, FT, order ? b.n_chunks - 1 - k : k, start_state, seg_cls);
  }
  R->has_invalid = counters[2];
  R->eot_rewind = m.eot_rewind;

  FastTables FTr;
  std::vector<uint16_t> hot_r;
  make_fast_tables(em->hm, m, 0, hot_r, FTr);
  uint8_t seg_cls_r[36];
  // K2b-d fix-up rounds
  std::vector<uint32_t> list, next, rew;
  for (uint32_t i = 1; i < b.n_chunks; i++) list.push_back(i);
  while (!list.empty()) {
    R->rounds++;
    rew.clear(); next.clear();
    for (size_t k = 0; k < list.size(); k++) {
      uint32_t i = list[order ? list.size() - 1 - k : k];
      if (chunk_stitch(m, b, i, R->rounds == 1)) rew.push_back(i);
    }
    R->n_stitch_mismatch += (uint32_t)rew.size();
    for (size_t k = 0; k < rew.size(); k++) {
      const uint32_t ci = rew[order ? rew.size() - 1 - k : k];
      if (mode == 0) chunk_rewalk(m, b, ci); else chunk_rewalk_fast(m, b, FTr, ci, seg_cls_r);
      R->n_rewalks++;
    }
    for (size_t k = 0; k < list.size(); k++) {
      uint32_t i = list[k];
      if (chunk_commit(b, i) && i + 1 < b.n_chunks) next.push_back(i + 1);
    }
    list.swap(next);
  }
  // walk errors
  for (uint32_t i = 0; i < b.n_chunks; i++) {
    if (E[i].flags & WS_INVALID) {
      uint32_t code = E[i].flags >> WS_ERR_SHIFT;
      unsigned long long key = ((unsigned long long)(i * chunk) << 8) | (code ? code : 0xFF);
      if (key < err_key) err_key = key;
    }
  }
  if (err_key != ~0ull) { R->status = (int)(err_key & 0xFF); return R; }
  const WState& last = E[b.n_chunks - 1];
  if (!(last.flags & WS_DONE)) { R->status = 0xFE; return R; }
  R->carry_state = em->hm.old_of_new[last.t];

  // K3 compaction: per-thread aggs -> block aggs -> scan -> texts pass -> emit
  CompactCtx c;
  std::memset(&c, 0, sizeof c);
  c.in = in; c.N = N; c.n_words = b.n_words; c.rstart = b.rstart; c.b_end = b.b_end; c.b_skip = b.b_skip;
  c.b_sent = b.b_sent; c.b_tend = b.b_tend; c.flags = flags; c.err_key = &err_key; c.eot_rewind = m.eot_rewind;
  const uint32_t TPB = 256, WPT = 2, WPB = TPB * WPT;
  const uint32_t nblk = (b.n_words + WPB - 1) / WPB;
  // like the reduce kernel: per block its summary, per warp unit (32 threads * WPT words) the summary of the units
  // before it in the block, marked when the unit itself holds a TextEnd
  const uint32_t UPB = TPB / 32, UW = 32 * WPT;  // units per block, words per unit
  std::vector<Agg> block_agg(nblk), block_carry(nblk), unit_prefix((size_t)nblk * UPB);
  for (uint32_t blk = 0; blk < nblk; blk++) {
    Agg acc = agg_zero();
    for (uint32_t u = 0; u < UPB; u++) {
      Agg ua = agg_zero();
      for (uint32_t k = 0; k < UW; k++) {
        uint32_t w = blk * WPB + u * UW + k;
        if (w < b.n_words) ua = agg_combine(ua, word_agg(w, word_load(c, w)));
      }
      Agg ex = acc;
      if (ua.n_text) ex.kinds |= WAGG_HAS_TEXT;
      unit_prefix[(size_t)blk * UPB + u] = ex;
      acc = agg_combine(acc, ua);
    }
    block_agg[blk] = acc;
  }
  Agg run = agg_stream_start(sentence_end_in != 0);
  for (uint32_t blk = 0; blk < nblk; blk++) { block_carry[blk] = run; run = agg_combine(run, block_agg[blk]); }
  const Agg total = run;
  R->n_runes = total.n_rune;
  size_t nt = total.n_tok, ns = total.n_sent + 1, nx = total.n_text + 1, np = total.n_sentpos + 1;
  R->tok_bytes = (uint32_t*)std::calloc(2 * nt + 2, 4);
  R->tok_pos = (int32_t*)std::calloc(2 * nt + 2, 4);
  R->tok_delta = (uint16_t*)std::calloc(4 * nt + 4, 2);
  R->tok_delta8 = (uint8_t*)std::calloc(4 * nt + 8, 1);
  R->esc = (uint32_t*)std::calloc(2 * (4 * nt + 4), 4);
  uint32_t esc_count = 0;
  R->sent_pos = (int32_t*)std::calloc(np + 1, 4);
  R->sent_tok = (uint32_t*)std::calloc(ns + 1, 4);
  R->text_tok_end = (uint32_t*)std::calloc(nx + 1, 4);
  R->text_sent_end = (uint32_t*)std::calloc(nx + 1, 4);
  R->text_sentpos_end = (uint32_t*)std::calloc(nx + 1, 4);
  R->text_byte_end = (uint32_t*)std::calloc(nx + 1, 4);
  std::vector<DocRec> docs(nx + 1);
  c.tok_bytes = R->tok_bytes; c.tok_pos = R->tok_pos; c.tok_delta = R->tok_delta;
  c.tok_delta8 = R->tok_delta8; c.esc = R->esc; c.esc_count = &esc_count; c.esc_cap = (uint32_t)(4 * nt + 4); c.sent_pos = R->sent_pos; c.sent_tok = R->sent_tok;
  c.text_tok_end = R->text_tok_end; c.text_sent_end = R->text_sent_end;
  c.text_sentpos_end = R->text_sentpos_end; c.text_byte_end = R->text_byte_end;
  c.docs = docs.data();
  docs[0] = doc_stream_start(c);
  // the delta-coded forms report into a key of their own: a range error there (DATOK_ERR_COMPACT_RANGE, only raised by
  // a call that asks for DATOK_COMPACT) leaves the absolute form intact
  unsigned long long err_key_delta = ~0ull;
  CompactCtx cd = c;
  cd.err_key = &err_key_delta;
  {  // texts pass: the marked units only, in any order, each seeded with block carry + unit prefix (like the kernel)
    const uint32_t n_units = nblk * UPB;
    for (uint32_t ku = 0; ku < n_units; ku++) {
      const uint32_t u = order ? n_units - 1 - ku : ku;
      if (!(unit_prefix[u].kinds & WAGG_HAS_TEXT)) continue;
      Agg carry = agg_combine(block_carry[u / UPB], load_warp_prefix(unit_prefix.data(), u));
      for (uint32_t k = 0; k < UW; k++) {
        const uint32_t w = u * UW + k;
        if (w >= b.n_words) continue;
        const WordBits wb = word_load(c, w);
        emit_texts(c, w, wb, carry);
        carry = agg_combine(carry, word_agg(w, wb));
      }
    }
  }
  for (int pass = 1; pass < 2; pass++) {  // tokens + sentences
    for (uint32_t kb = 0; kb < nblk; kb++) {
      uint32_t blk = order ? nblk - 1 - kb : kb;
      Agg carry = block_carry[blk];
      // like the kernel: a block's tokens are staged block-relative when they fit
      const uint32_t blk_tok0 = carry.n_tok, blk_ntok = block_agg[blk].n_tok;
      const bool staged = blk_ntok <= 64;
      std::vector<uint32_t> s_tb(2 * 64 + 2);
      std::vector<int32_t> s_tp(2 * 64 + 2);
      std::vector<uint16_t> s_td(4 * 64 + 4), s_t8(2 * 64 + 4);
      for (uint32_t t = 0; t < TPB; t++) {
        // (a warp of the emit kernel seeds its scan with the block's carry and its unit's prefix)
        if (t % 32 == 0) carry = agg_combine(block_carry[blk], load_warp_prefix(unit_prefix.data(), (size_t)blk * UPB + t / 32));
        for (uint32_t k = 0; k < WPT; k++) {
          uint32_t w = blk * WPB + t * WPT + k;
          if (w >= b.n_words) continue;
          const WordBits wb = word_load(c, w);
          if (wb.e | wb.s | wb.t) {
            const WordMasks wm = word_masks(wb, agg_last(carry));
            emit_tokens<0>(c, w, wb, wm, carry, staged ? s_tb.data() : c.tok_bytes, staged ? s_tp.data() : c.tok_pos,
                           nullptr, staged ? blk_tok0 : 0u);
            emit_tokens<1>(cd, w, wb, wm, carry, nullptr, nullptr, staged ? s_td.data() : c.tok_delta,
                           staged ? blk_tok0 : 0u);
            emit_tokens<2>(cd, w, wb, wm, carry, nullptr, nullptr,
                           staged ? s_t8.data() : reinterpret_cast<uint16_t*>(c.tok_delta8), staged ? blk_tok0 : 0u);
            emit_sentences(c, w, wb, wm, carry);
          }
          carry = agg_combine(carry, word_agg(w, wb));
        }
      }
      if (pass == 1 && staged) {
        std::memcpy(c.tok_bytes + 2 * (size_t)blk_tok0, s_tb.data(), 8 * (size_t)blk_ntok);
        std::memcpy(c.tok_pos + 2 * (size_t)blk_tok0, s_tp.data(), 8 * (size_t)blk_ntok);
        std::memcpy(c.tok_delta + 4 * (size_t)blk_tok0, s_td.data(), 8 * (size_t)blk_ntok);
        std::memcpy(c.tok_delta8 + 4 * (size_t)blk_tok0, s_t8.data(), 4 * (size_t)blk_ntok);
      }
    }
  }
  R->n_esc = esc_count;
  const StreamTotals fin = finalize_stream(c, total, text_end_in != 0, b.final_input != 0);
  R->n_tokens = fin.n_tok; R->n_sentences = fin.n_sent; R->n_texts = fin.n_text; R->n_sent_pos = fin.n_sentpos;
  if (err_key != ~0ull) R->status = (int)(err_key & 0xFF);
  if (err_key_delta != ~0ull) {
    if ((err_key_delta & 0xFF) == E_COMPACT_RANGE) R->delta_range = 1;
    else if (R->status == 0) R->status = (int)(err_key_delta & 0xFF);
  }
  if (R->status == 0 && !R->has_invalid) {
    // the device formatter, item by item like its kernels (format_kernels.cu): scans, then the writers
    FmtCtx f;
    std::memset(&f, 0, sizeof f);
    f.in = in; f.tok_bytes = R->tok_bytes; f.tok_pos = R->tok_pos; f.sent_pos = R->sent_pos; f.sent_tok = R->sent_tok;
    f.text_tok_end = R->text_tok_end; f.text_sent_end = R->text_sent_end; f.text_sentpos_end = R->text_sentpos_end;
    f.n_tok = (uint32_t)R->n_tokens; f.n_sent = (uint32_t)R->n_sentences; f.n_sentpos = (uint32_t)R->n_sent_pos;
    f.n_text = (uint32_t)R->n_texts; f.flags = flags & 15u;
    std::vector<uint32_t> loc[4];
    std::vector<unsigned long long> base[4];
    auto scan = [&](int which, uint32_t n, FmtScan& sc) {
      loc[which].assign((size_t)n + 1, 0);
      base[which].assign((size_t)n / FMT_TILE + 2, 0);
      unsigned long long run = 0;
      uint32_t in_tile = 0;
      for (uint32_t i = 0; i <= n; i++) {
        if ((i & (FMT_TILE - 1)) == 0) { base[which][i >> FMT_TILE_SHIFT] = run; in_tile = 0; }
        loc[which][i] = in_tile;
        sc.local = loc[which].data(); sc.base = base[which].data();
        if (i < n) {
          const uint32_t v = which == 0 ? fmt_len_tok(f, i) : which == 1 ? fmt_len_pos(f, i) : which == 2 ? fmt_len_sp(f, i) : fmt_len_text(f, i);
          in_tile += v; run += v;
        }
      }
    };
    scan(0, f.n_tok, f.ptok); scan(1, f.n_tok, f.ppos); scan(2, f.n_sentpos, f.psp); scan(3, f.n_text, f.px);
    R->text_len = fmt_total(f);
    R->text = (uint8_t*)std::malloc(R->text_len + 16);
    std::memset(R->text, 0xEE, R->text_len + 16);
    f.out = R->text;
    for (uint32_t k = 0; k < f.n_tok; k++) fmt_write_token(f, order ? f.n_tok - 1 - k : k);
    for (uint32_t i = 0; i < f.n_sent; i++) fmt_write_sentence(f, i);
    for (uint32_t i = 0; i < f.n_text; i++) fmt_write_text(f, i);
    for (uint32_t i = 0; i < f.n_sentpos; i++) fmt_write_sentpos(f, i);
  }
  return R;
}

void emul_result_free(EmulResult* r) {
  if (!r) return;
  std::free(r->tok_bytes); std::free(r->tok_pos); std::free(r->tok_delta); std::free(r->tok_delta8); std::free(r->esc); std::free(r->sent_pos); std::free(r->sent_tok);
  std::free(r->text); std::free(r->text_tok_end); std::free(r->text_sent_end); std::free(r->text_sentpos_end); std::free(r->text_byte_end);
  std::free(r);
}

}  // extern "C"
