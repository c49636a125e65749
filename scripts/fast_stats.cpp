// fast_stats.cpp -- TEST TOOLING: runs the fused fast path of datok_b200/csrc on the CPU (like tests/emul)
// with event counters, to see how often the rare paths fire on a corpus.
//   g++ -O2 -std=c++17 -DDATOK_COUNT scripts/fast_stats.cpp datok_b200/csrc/model.cpp -lz -o /tmp/fast_stats
//   /tmp/fast_stats model.matok corpus.bin [chunk] [hot_rows]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
static unsigned long long g_exact_calls, g_exact_steps, g_guard, g_fast, g_bt_ok, g_bt_hard, g_bt_far, g_bt_dead,
    g_bt_third, g_zone, g_mark, g_cold, g_bt_inline, g_bt_far_fast;
static unsigned long long g_hist[65536];
static unsigned long long g_probe[4];  // hand-off: probes started, probe segments walked, arrivals, arrivals decided at once
static unsigned long long g_nf[5];  // exact-walker calls by reason: fast path gave up here, position outside the segment, flags, stale bufft, window guard
#include "../datok_b200/csrc/chunk_core.cuh"
#include "../datok_b200/csrc/model.hpp"
using namespace datok;
int run(int argc, char** argv, HostModel& hm, bool report);
int main(int argc, char** argv) {
  HostModel hm; std::string why;
  if (load_matok_file(argv[1], hm, why)) { fprintf(stderr, "%s\n", why.c_str()); return 1; }
  run(argc, argv, hm, false);  // calibration pass: visits per state
  std::vector<uint64_t> hist_old(hm.stateCount + 1, 0);
  for (int t = 1; t <= hm.stateCount; t++) hist_old[hm.old_of_new[t]] = g_hist[t];
  if (build_layout(hm, why, hist_old.data())) return 1;
  g_exact_calls = g_exact_steps = g_guard = g_fast = g_bt_ok = g_bt_hard = g_bt_far = g_bt_dead = g_bt_third = g_zone = g_mark = g_cold = g_bt_inline = g_bt_far_fast = 0;
  memset(g_nf, 0, sizeof g_nf); memset(g_probe, 0, sizeof g_probe);
  return run(argc, argv, hm, true);
}
int run(int argc, char** argv, HostModel& hm, bool report) {
  FILE* f = fopen(argv[2], "rb"); fseek(f, 0, SEEK_END); size_t n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> in(n + 64); if (fread(in.data(), 1, n, f) != n) return 1; fclose(f);
  uint32_t chunk = argc > 3 ? atoi(argv[3]) : 256, hot_rows = argc > 4 ? atoi(argv[4]) : 440;
  DeviceModel m; memset(&m, 0, sizeof m);
  m.table = hm.table.data(); m.table2 = hm.table2.data(); m.row_shift = hm.row_shift; m.start = hm.start;
  m.n_classes = hm.n_classes; m.stride2 = hm.stride2; m.hot16 = hm.hot16.data(); m.stride16 = hm.stride16; m.hot16_rows = hm.hot16_rows; m.hot_cols = hm.hot_cols; m.eot_rewind = hm.eot_rewind ? 1u : 0u;
  m.cls.ascii_cls = hm.ascii_cls; m.cls.latin1_cls = hm.latin1_cls; m.cls.rune_key = hm.rune_key.data();
  m.cls.rune_cls = hm.rune_cls.data(); m.cls.n_rune = hm.rune_key.size(); m.cls.identity_cls = hm.identity_cls;
  memcpy(m.sync_ascii, hm.sync_ascii, sizeof hm.sync_ascii); memcpy(m.sync_cls, hm.sync_mask, sizeof hm.sync_mask);
  WalkBuffers b; memset(&b, 0, sizeof b);
  b.in = in.data(); b.N = n; b.chunk = chunk; b.final_input = 1; b.n_chunks = n / chunk + 1; b.n_words = b.n_chunks * (chunk / 32);
  uint32_t counters[8] = {0}; b.counters = counters;
  std::vector<uint32_t> w[5]; for (auto& x : w) x.assign(b.n_words, 0);
  b.rstart = w[0].data(); b.b_end = w[1].data(); b.b_skip = w[2].data(); b.b_sent = w[3].data(); b.b_tend = w[4].data();
  std::vector<WState> E(b.n_chunks), A(b.n_chunks), En(b.n_chunks), Y(b.n_chunks);
  std::vector<uint32_t> sync(b.n_chunks), fh(b.n_chunks), cf(b.n_chunks);
  b.E = E.data(); b.exitA = A.data(); b.Enew = En.data(); b.Ytmp = Y.data(); b.sync = sync.data(); b.first_hw = fh.data(); b.cflags = cf.data();
  unsigned long long ek = ~0ull; b.err_key = &ek;
  if (hot_rows > hm.hot16_rows) hot_rows = hm.hot16_rows;
  std::vector<uint16_t> hot(hm.hot16.begin(), hm.hot16.begin() + (size_t)hot_rows * hm.stride16);
  for (auto& e : hot) if ((e & F16_TGT) >= hot_rows) e = 0;
  hot.resize(hot.size() + hm.stride16, 0);
  uint8_t lut2[256]; for (int i = 0; i < 128; i++) { lut2[i] = cap_cl2(hm.ascii_cls[i], 2 * hm.hot_cols); lut2[128 + i] = 128 + i; }
  FastTables FT; FT.hot16 = hot.data(); FT.t3 = m.table2; FT.n_hot = hot_rows; FT.row16 = hm.stride16 * 2; FT.stride3 = m.stride2;
  FT.hot_saddr = 0; FT.ascii_cls2 = lut2; FT.stop_cl2 = 2 * hm.hot_cols; FT.sync_cls = hm.sync_mask; FT.eot_rewind = hm.eot_rewind ? 1u : 0u;
  uint8_t cls[36];
  for (uint32_t i = 0; i < b.n_chunks; i++) chunk_spec_fast(m, b, FT, i, m.start, cls);
  if (report) printf("bytes %zu chunks %u\nfast steps %llu (cold %llu = %.3f%%)\nbacktracks in place %llu (%.3f%% of steps), stale zones %llu, inline %llu\n"
         "slow: marks %llu, hard %llu, far %llu (fast %llu), dead %llu, third %llu; guard %llu\nexact calls %llu steps %llu (%.3f%% of bytes)\n",
         n, b.n_chunks, g_fast, g_cold, 100.0 * g_cold / g_fast, g_bt_ok, 100.0 * g_bt_ok / g_fast, g_zone, g_bt_inline, g_mark, g_bt_hard, g_bt_far, g_bt_far_fast,
         g_bt_dead, g_bt_third, g_guard, g_exact_calls, g_exact_steps, 100.0 * g_exact_steps / n);
  if (report) printf("hand-off: %llu arrivals in the fast path, %llu decided at once, %llu by the look-ahead, %llu probes over %llu segments\n",
         g_probe[2], g_probe[3], g_probe[2] - g_probe[3] - g_probe[0], g_probe[0], g_probe[1]);
  if (report) printf("exact calls by reason: gave up here %llu, outside the segment %llu, flags %llu, stale bufft %llu, window guard %llu\n",
         g_nf[0], g_nf[1], g_nf[2], g_nf[3], g_nf[4]);
  return 0;
}
