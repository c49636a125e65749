// api.cu -- C ABI of libdatok_b200.so (include/datok_b200.h).
//
// Host-side orchestration of the kernels in kernels.cu: workspace management,
// the fix-up round loop, output allocation, host<->device copies and timing.
// There is no CPU transduction path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/datok_b200.h"
#include "format_kernels.cuh"
#include "kernels.cuh"
#include "model.hpp"

using namespace datok;

static thread_local std::string g_last_error;

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      g_last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
      return DATOK_ERR_CUDA;                                                                \
    }                                                                                       \
  } while (0)

#define CUDA_TRYF(expr)                                                                     \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      g_last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
      return fail(DATOK_ERR_CUDA);                                                          \
    }                                                                                       \
  } while (0)

namespace {

// every entry point that selects the model's device leaves the caller's current device as it was
struct DeviceGuard {
  int dev = -1;
  DeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } }
  ~DeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
};

struct Block {
  void* p = nullptr;
  size_t bytes = 0;
  bool host = false;
};

enum { T_CLEAR, T_WALK, T_STITCH, T_REWALK, T_COMMIT, T_REDUCE, T_SCAN, T_TEXTS, T_EMIT, T_COUNT };
const char* const kTimerNames[T_COUNT] = {"clear", "walk_fused", "stitch", "rewalk", "commit",
                                          "compact_reduce", "compact_scan", "compact_texts", "compact_emit"};

}  // namespace

struct datok_model {
  HostModel hm;
  int device = 0;
  cudaStream_t stream = nullptr;
  // model tables on the device: [fused table u32 | exact table u16] in one allocation
  uint8_t* d_tables = nullptr;
  size_t d_tables_bytes = 0;
  uint8_t* d_cls_tables = nullptr;  // ascii[128] latin1[128] rune_cls[n]
  uint32_t* d_rune_key = nullptr;
  DeviceModel dm;
  int n_sms = 148;
  size_t smem_optin = 0;
  uint32_t n_hot = 1;
  int fused_threads = 1024;
  std::atomic<bool> calibrated{false};
  bool auto_calibrate = true;
  // workspace (grow only)
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  // cache of result buffers
  std::vector<Block> cache;
  std::mutex mu;
  uint32_t chunk = 640;  // bytes per lane and pass (multiple of 32; DATOK_CHUNK).  Measured on 1 GiB, German / English / long document:
                         // 512: 6.88 / 6.40 / 17.4 ms, 640: 6.72 / 6.25 / 15.4 ms, 768: 6.73 / 6.32 / 15.1 ms, 384: 7.20 ms
  // instrumentation of the last call
  float t_ms[T_COUNT] = {0};
  int launches = 0;
  uint32_t last_rounds = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> tev;
  // pipelined host path: copy streams and double-buffered device staging
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  // (three input buffers: the copy engine runs up to two pieces ahead of the kernels; two result slots)
  cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_free[3] = {nullptr, nullptr, nullptr}, ev_emit[2] = {nullptr, nullptr},
              ev_out[2] = {nullptr, nullptr};
  uint32_t* h_mail = nullptr;    // mapped pinned host memory the kernels report small results through
  uint32_t* d_mail = nullptr;    // its device address
  Block d_in[3];                 // input pieces
  Block d_out[2][7];             // per slot: tok_bytes, tok_pos, sent_pos, sent_tok, text arrays + DocRec, formatter scratch, text
  size_t piece_bytes = (size_t)64 << 20;  // smallest piece; also: inputs >= 2 * piece_bytes are pipelined
  bool piece_fixed = false;               // DATOK_PIECE_MB: every piece has piece_bytes (else sized per call)
  bool pipelined = true;
  // results hold pooled buffers of their model: the model outlives them
  std::atomic<int> live_results{0};   // (counted on the primary for all of its execution contexts)
  std::atomic<bool> freed_by_user{false};
  // Concurrent calls on one model (the reference's model is immutable and shareable, matrix.go:16-26): a call runs on
  // an execution context -- stream, workspace, staging buffers, mailbox -- of its own.  The model handed to the caller
  // is the primary context; further ones are clones that share its device-resident tables and are created when a call
  // finds every context busy (DATOK_MAX_CONCURRENCY, default 4, bounds them).
  datok_model* primary = nullptr;     // clones: the model they belong to
  std::vector<datok_model*> clones;   // primary: its clones (guarded by ctx_mu)
  std::mutex ctx_mu;
  bool busy = false;                  // guarded by the primary's ctx_mu
  int max_ctx = 4;
};
static datok_model* root_of(datok_model* m) { return m->primary ? m->primary : m; }

struct datok_result {
  datok_model* model = nullptr;
  datok_view view;
  std::vector<uint32_t> esc;  // DATOK_COMPACT8 escape pairs, sorted
  std::vector<Block> blocks;
  bool device = false;
};

namespace {

Block acquire(datok_model* m, size_t bytes, bool host, int* rc) {
  bytes = std::max<size_t>(bytes, 256);
  int best = -1;
  for (size_t i = 0; i < m->cache.size(); i++) {
    const Block& b = m->cache[i];
    if (b.host == host && b.bytes >= bytes && (best < 0 || b.bytes < m->cache[best].bytes)) best = (int)i;
  }
  if (best >= 0 && m->cache[best].bytes <= bytes * 2 + (1 << 20)) {
    Block b = m->cache[best];
    m->cache.erase(m->cache.begin() + best);
    return b;
  }
  Block b;
  b.bytes = bytes + bytes / 8;
  b.host = host;
  cudaError_t e = host ? cudaHostAlloc(&b.p, b.bytes, cudaHostAllocDefault) : cudaMalloc(&b.p, b.bytes);
  if (e != cudaSuccess) {
    g_last_error = std::string(host ? "cudaHostAlloc: " : "cudaMalloc: ") + cudaGetErrorString(e);
    *rc = DATOK_ERR_CUDA;
    b.p = nullptr;
  }
  return b;
}

void release(datok_model* m, const Block& b) {
  if (!b.p) return;
  size_t cached = 0;
  for (auto& c : m->cache) cached += c.bytes;
  if (m->cache.size() < 64 && cached < ((size_t)16 << 30)) { m->cache.push_back(b); return; }
  if (b.host) cudaFreeHost(b.p); else cudaFree(b.p);
}

// caller holds m->mu
void free_result_locked(datok_result* r) {
  datok_model* m = r->model;
  for (auto& b : r->blocks) release(m, b);
  root_of(m)->live_results--;
  delete r;
}

void destroy_model(datok_model* m) {
  DeviceGuard guard;
  for (datok_model* c : m->clones) destroy_model(c);
  m->clones.clear();
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (auto& b : m->cache) { if (b.host) cudaFreeHost(b.p); else cudaFree(b.p); }
  if (m->ws) cudaFree(m->ws);
  if (!m->primary) {  // (a clone shares the tables of its primary)
    if (m->d_tables) cudaFree(m->d_tables);
    if (m->d_cls_tables) cudaFree(m->d_cls_tables);
    if (m->d_rune_key) cudaFree(m->d_rune_key);
  }
  for (auto& e : m->ev) if (e) cudaEventDestroy(e);
  for (int i = 0; i < 3; i++) {
    for (cudaEvent_t e : {m->ev_in[i], m->ev_free[i]}) if (e) cudaEventDestroy(e);
    if (m->d_in[i].p) cudaFree(m->d_in[i].p);
  }
  for (int i = 0; i < 2; i++) {
    for (cudaEvent_t e : {m->ev_emit[i], m->ev_out[i]}) if (e) cudaEventDestroy(e);
    for (auto& blk : m->d_out[i]) if (blk.p) cudaFree(blk.p);
  }
  if (m->h_mail) cudaFreeHost(m->h_mail);
  if (m->s_h2d) cudaStreamDestroy(m->s_h2d);
  if (m->s_d2h) cudaStreamDestroy(m->s_d2h);
  for (auto& e : m->tev) cudaEventDestroy(e);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Carver {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// lays the workspace out for an input of N bytes; base == nullptr just measures
size_t carve(uint8_t* base, uint32_t N, uint32_t chunk, bool need_input_copy, WalkBuffers& b, CompactBuffers& cb) {
  Carver c{base};
  b.N = N;
  b.chunk = chunk;
  b.n_chunks = N / chunk + 1;
  b.n_words = b.n_chunks * (chunk / 32);
  uint8_t* d_in = c.take<uint8_t>(need_input_copy ? (size_t)N + 64 : 0);
  if (need_input_copy) b.in = d_in;
  b.rstart = c.take<uint32_t>(b.n_words);
  // the four event bitmaps are contiguous so that one memset clears them
  b.b_end = c.take<uint32_t>((size_t)b.n_words * 4);
  b.b_skip = b.b_end ? b.b_end + b.n_words : nullptr;
  b.b_sent = b.b_end ? b.b_end + 2 * (size_t)b.n_words : nullptr;
  b.b_tend = b.b_end ? b.b_end + 3 * (size_t)b.n_words : nullptr;
  b.E = c.take<WState>(b.n_chunks);
  b.exitA = c.take<WState>(b.n_chunks);
  b.Enew = c.take<WState>(b.n_chunks);
  b.Ytmp = c.take<WState>(b.n_chunks);
  b.sync = c.take<uint32_t>(b.n_chunks);
  b.first_hw = c.take<uint32_t>(b.n_chunks);
  b.cflags = c.take<uint32_t>(b.n_chunks);
  b.list_cur = c.take<uint32_t>(b.n_chunks);
  b.list_next = c.take<uint32_t>(b.n_chunks);
  b.list_rewalk = c.take<uint32_t>(b.n_chunks);
  b.counters = c.take<uint32_t>(8);
  b.err_key = c.take<unsigned long long>(1);
  const uint32_t wpb = COMPACT_THREADS * COMPACT_WPT;
  cb.n_blocks = (b.n_words + wpb - 1) / wpb;
  cb.block_agg = c.take<Agg>(cb.n_blocks);
  cb.block_carry = c.take<Agg>(cb.n_blocks);
  cb.warp_agg = c.take<Agg>((size_t)cb.n_blocks * (COMPACT_THREADS / 32));
  cb.super_agg = c.take<Agg>(cb.n_blocks / SCAN_THREADS + 1);
  cb.super_carry = c.take<Agg>(cb.n_blocks / SCAN_THREADS + 1);
  cb.total = c.take<Agg>(2);
  return align_up(c.off, 256);
}

int ensure_workspace(datok_model* m, size_t need) {
  if (need <= m->ws_bytes) return DATOK_OK;
  if (m->ws) cudaFree(m->ws);
  m->ws = nullptr;
  m->ws_bytes = 0;
  size_t want = need + need / 16;
  CUDA_TRY(cudaMalloc(&m->ws, want));
  m->ws_bytes = want;
  return DATOK_OK;
}

struct PhaseTimer {
  datok_model* m;
  size_t used = 0;
  std::vector<std::pair<int, size_t>> spans;  // (timer id, index of start event)
  cudaEvent_t next() {
    if (used == m->tev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      m->tev.push_back(e);
    }
    return m->tev[used++];
  }
  void begin(int id) {
    cudaEvent_t e = next();
    cudaEventRecord(e, m->stream);
    spans.emplace_back(id, used - 1);
  }
  void end() {
    cudaEvent_t e = next();
    cudaEventRecord(e, m->stream);
  }
  void collect() {  // after a stream synchronize
    for (auto& sp : spans) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, m->tev[sp.second], m->tev[sp.second + 1]) == cudaSuccess) m->t_ms[sp.first] += ms;
    }
  }
};

int upload_model(datok_model* m) {
  HostModel& h = m->hm;
  if (m->d_tables) { cudaFree(m->d_tables); m->d_tables = nullptr; }
  if (m->d_cls_tables) { cudaFree(m->d_cls_tables); m->d_cls_tables = nullptr; }
  if (m->d_rune_key) { cudaFree(m->d_rune_key); m->d_rune_key = nullptr; }
  const size_t t2_bytes = align_up(h.table2.size() * sizeof(uint32_t), 256);
  const size_t t1_bytes = align_up(h.table.size() * sizeof(uint16_t), 256);
  const size_t h16_bytes = h.hot16.size() * sizeof(uint16_t);
  m->d_tables_bytes = t2_bytes + t1_bytes + h16_bytes;
  CUDA_TRY(cudaMalloc(&m->d_tables, m->d_tables_bytes));
  CUDA_TRY(cudaMemcpy(m->d_tables, h.table2.data(), h.table2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(m->d_tables + t2_bytes, h.table.data(), h.table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(m->d_tables + t2_bytes + t1_bytes, h.hot16.data(), h16_bytes, cudaMemcpyHostToDevice));
  const size_t nr = h.rune_key.size();
  std::vector<uint8_t> blob(256 + nr + 16, 0);
  std::memcpy(blob.data(), h.ascii_cls, 128);
  std::memcpy(blob.data() + 128, h.latin1_cls, 128);
  if (nr) std::memcpy(blob.data() + 256, h.rune_cls.data(), nr);
  CUDA_TRY(cudaMalloc(&m->d_cls_tables, blob.size()));
  CUDA_TRY(cudaMemcpy(m->d_cls_tables, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&m->d_rune_key, (nr + 4) * sizeof(uint32_t)));
  if (nr) CUDA_TRY(cudaMemcpy(m->d_rune_key, h.rune_key.data(), nr * sizeof(uint32_t), cudaMemcpyHostToDevice));
  DeviceModel& d = m->dm;
  d.table2 = reinterpret_cast<const uint32_t*>(m->d_tables);
  d.table = reinterpret_cast<const uint16_t*>(m->d_tables + t2_bytes);
  d.hot16 = reinterpret_cast<const uint16_t*>(m->d_tables + t2_bytes + t1_bytes);
  d.row_shift = h.row_shift; d.start = h.start; d.n_classes = h.n_classes; d.stride2 = h.stride2;
  d.stride16 = h.stride16; d.hot16_rows = h.hot16_rows; d.hot_cols = h.hot_cols;
  d.eot_rewind = h.eot_rewind ? 1u : 0u;
  d.cls.ascii_cls = m->d_cls_tables;
  d.cls.latin1_cls = m->d_cls_tables + 128;
  d.cls.rune_cls = m->d_cls_tables + 256;
  d.cls.rune_key = m->d_rune_key;
  d.cls.n_rune = (uint32_t)nr;
  d.cls.identity_cls = h.identity_cls;
  d.cls.self = nullptr;  // (set by the kernels: a copy in shared memory)
  std::memcpy(d.sync_ascii, h.sync_ascii, sizeof d.sync_ascii);
  std::memcpy(d.sync_cls, h.sync_mask, sizeof d.sync_cls);
  m->n_hot = fused_max_hot_rows(d, m->smem_optin, (uint32_t)h.stateCount, m->fused_threads);
  if (const char* s = std::getenv("DATOK_HOT_ROWS")) {
    long v = std::atol(s);
    if (v >= 1 && (uint32_t)v < m->n_hot) m->n_hot = (uint32_t)v;
  }
  return DATOK_OK;
}

// Keep the transition tables resident in L2 (they are gathered from once per input byte).
void pin_table_in_l2(datok_model* m) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, m->device) != cudaSuccess) return;
  const size_t bytes = m->d_tables_bytes;
  if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
  size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes * 2);
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
  cudaStreamAttrValue attr;
  std::memset(&attr, 0, sizeof attr);
  attr.accessPolicyWindow.base_ptr = m->d_tables;
  attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
  attr.accessPolicyWindow.hitRatio = 1.0f;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(m->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
  cudaGetLastError();
}

// Re-orders the states by measured visit frequency on a sample of the caller's own
// text (GPU histogram), so that the rows kept in shared memory are the hot ones.
int calibrate_locked(datok_model* m, const WalkBuffers& full) {
  WalkBuffers b = full;
  const uint32_t sample = std::min<uint32_t>(full.N, 8u << 20);
  b.N = sample;
  b.n_chunks = sample / b.chunk + 1;
  const size_t S = (size_t)m->hm.stateCount;
  uint32_t* d_hist = nullptr;
  const size_t hist_words = S + 2 + 256;  // visits per state, then occurrences per class
  CUDA_TRY(cudaMalloc(&d_hist, hist_words * sizeof(uint32_t)));
  CUDA_TRY(cudaMemsetAsync(d_hist, 0, hist_words * sizeof(uint32_t), m->stream));
  CUDA_TRY(cudaMemsetAsync(b.b_end, 0, (size_t)full.n_words * 4 * sizeof(uint32_t), m->stream));
  launch_hist(m->dm, b, d_hist, (uint32_t)(S + 2), m->stream);
  std::vector<uint32_t> hist(hist_words);
  CUDA_TRY(cudaMemcpyAsync(hist.data(), d_hist, hist_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  cudaFree(d_hist);
  std::vector<uint64_t> hist_old(S + 1, 0), cls_hist(256, 0);
  for (size_t t = 1; t <= S; t++) hist_old[m->hm.old_of_new[t]] = hist[t];
  for (size_t c = 0; c < m->hm.n_classes; c++) cls_hist[m->hm.cls_base[c]] = hist[S + 2 + c];
  std::string why;
  {  // what the walk kernel can spend on compact rows (kernels.cu fused_max_hot_rows)
    DeviceModel probe = m->dm;
    probe.stride16 = 2;  // (4-byte rows: the row count then is the byte budget / 4)
    probe.hot16_rows = 0x7FFFFFFF;
    m->hm.row_budget_bytes = 4u * fused_max_hot_rows(probe, m->smem_optin, 0x7FFFFFF0u, m->fused_threads);
  }
  int rc = build_layout(m->hm, why, hist_old.data(), cls_hist.data());
  if (rc) { g_last_error = why; return rc; }
  rc = upload_model(m);
  if (rc) return rc;
  pin_table_in_l2(m);
  m->calibrated = true;
  return DATOK_OK;
}

static bool init_context(datok_model* m);
datok_model* finish_load(datok_model* m, int device, int* err) {
  DeviceGuard guard;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    g_last_error = "no usable CUDA device (this library has no CPU path)";
    cudaGetLastError();
    *err = DATOK_ERR_NO_DEVICE;
    delete m;
    return nullptr;
  }
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
      prop.major != 10) {
    g_last_error = "device is not an sm_100 (B200) GPU; the kernels are built for sm_100a only";
    cudaGetLastError();
    *err = DATOK_ERR_NO_DEVICE;
    delete m;
    return nullptr;
  }
  m->device = device;
  if (!init_context(m)) {
    g_last_error = "creating the model's streams / mailbox failed";
    cudaGetLastError();
    *err = DATOK_ERR_CUDA;
    datok_free(m);
    return nullptr;
  }
  if (const char* s = std::getenv("DATOK_MAX_CONCURRENCY")) {
    const long v = std::atol(s);
    if (v >= 1 && v <= 64) m->max_ctx = (int)v;
  }
  if (const char* s = std::getenv("DATOK_NO_PIPELINE")) m->pipelined = !(s[0] == '1');
  if (const char* s = std::getenv("DATOK_PIECE_MB")) {
    long v = std::atol(s);
    if (v >= 1 && v <= 2048) { m->piece_bytes = (size_t)v << 20; m->piece_fixed = true; }
  }
  m->n_sms = prop.multiProcessorCount;
  m->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  m->fused_threads = fused_threads_from_env();
  if (const char* s = std::getenv("DATOK_NO_CALIBRATE")) m->auto_calibrate = !(s[0] == '1');
  if (const char* s = std::getenv("DATOK_HOT_COLS")) {  // columns of the compact rows (default: from the calibration)
    const long v = std::atol(s);
    if (v >= 1 && v <= 256) {
      m->hm.force_hot_cols = (uint32_t)v;
      std::string why;
      const int rc2 = build_layout(m->hm, why);
      if (rc2) { g_last_error = why; *err = rc2; datok_free(m); return nullptr; }
    }
  }
  int rc = upload_model(m);
  if (rc) { *err = rc; datok_free(m); return nullptr; }
  pin_table_in_l2(m);
  if (const char* s = std::getenv("DATOK_CHUNK")) {
    long v = std::atol(s);
    if (v >= 32 && v <= 65536 && v % 32 == 0) m->chunk = (uint32_t)v;
  }
  *err = DATOK_OK;
  return m;
}

// the per-context resources of finish_load (stream, events, mailbox, copy streams)
static bool init_context(datok_model* m) {
  if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) return false;
  for (auto& e : m->ev) cudaEventCreate(&e);
  if (cudaHostAlloc((void**)&m->h_mail, 4096, cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&m->d_mail, m->h_mail, 0) != cudaSuccess)
    return false;
  cudaStreamCreateWithFlags(&m->s_h2d, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&m->s_d2h, cudaStreamNonBlocking);
  for (int i = 0; i < 3; i++) {
    cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&m->ev_free[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < 2; i++) {
    cudaEventCreateWithFlags(&m->ev_emit[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&m->ev_out[i], cudaEventDisableTiming);
  }
  return true;
}

// another execution context of P: its own stream, workspace and buffers, P's device-resident tables
static datok_model* clone_context(datok_model* P) {
  DeviceGuard guard;
  if (cudaSetDevice(P->device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  datok_model* c = new datok_model();
  c->primary = P;
  c->hm = P->hm;  // (host-side numbering tables; the device tables are shared)
  c->device = P->device;
  c->d_tables = P->d_tables; c->d_tables_bytes = P->d_tables_bytes;
  c->d_cls_tables = P->d_cls_tables; c->d_rune_key = P->d_rune_key;
  c->dm = P->dm;
  c->n_sms = P->n_sms; c->smem_optin = P->smem_optin; c->n_hot = P->n_hot; c->fused_threads = P->fused_threads;
  c->calibrated = true; c->auto_calibrate = false;
  c->chunk = P->chunk; c->piece_bytes = P->piece_bytes; c->piece_fixed = P->piece_fixed; c->pipelined = P->pipelined;
  if (!init_context(c)) { cudaGetLastError(); destroy_model(c); return nullptr; }
  pin_table_in_l2(c);
  return c;
}

// K1+K2: clear, fused walk, fix-up rounds (host-synchronised on the round counters), error collection
int do_walk(datok_model* m, WalkBuffers& b, uint32_t start_state, PhaseTimer& pt) {
  cudaStream_t s = m->stream;
  pt.begin(T_CLEAR);  // (the boundary bitmaps are fully written by the walk itself)
  CUDA_TRY(cudaMemsetAsync(b.counters, 0, 8 * sizeof(uint32_t), s));
  CUDA_TRY(cudaMemsetAsync(b.err_key, 0xFF, sizeof(unsigned long long), s));
  pt.end();
  pt.begin(T_WALK);
  {
    const int e = launch_walk_fused(m->dm, b, start_state, m->n_hot, m->n_sms, s, m->fused_threads);
    if (e != 0) { g_last_error = std::string("walk_fused launch: ") + cudaGetErrorString((cudaError_t)e); return DATOK_ERR_CUDA; }
  }
  pt.end();
  m->launches += 1;
  uint32_t n_list = b.n_chunks - 1;
  const uint32_t* list = nullptr;  // round 1: every chunk but the first
  uint32_t* cur = b.list_cur;
  uint32_t* nxt = b.list_next;
  uint32_t single_rounds = 0;  // consecutive rounds whose list held one chunk: a chain of dependent chunks
  while (n_list) {
    m->last_rounds++;
    b.list_next = nxt;
    single_rounds = (n_list == 1 && list) ? single_rounds + 1 : 0;
    if (single_rounds > 3) {
      // text without sync points (no whitespace the root state skips: minified markup, CSV ...): every chunk waits
      // for its predecessor's exit state.  The chain is followed on the device, thousands of chunks per launch,
      // instead of one host-synchronised round per chunk.
      pt.begin(T_STITCH);
      launch_chain(m->dm, b, list, 4096, s);
      pt.end();
      m->launches += 2;
      {
        MailSrc ms;
        std::memset(&ms, 0, sizeof ms);
        ms.p[0] = b.counters; ms.words[0] = 2; ms.off[0] = 0;
        launch_mail(ms, m->d_mail, s);
      }
      CUDA_TRY(cudaMemsetAsync(b.counters, 0, 2 * sizeof(uint32_t), s));
      CUDA_TRY(cudaStreamSynchronize(s));
      n_list = m->h_mail[0];
      std::swap(cur, nxt);
      list = cur;
      continue;
    }
    pt.begin(T_STITCH);
    launch_stitch(m->dm, b, list, n_list, s);
    pt.end();
    pt.begin(T_REWALK);
    {
      const int e = launch_rewalk_fused(m->dm, b, n_list, m->n_hot, m->n_sms, s);
      if (e != 0) { g_last_error = std::string("rewalk launch: ") + cudaGetErrorString((cudaError_t)e); return DATOK_ERR_CUDA; }
    }
    pt.end();
    pt.begin(T_COMMIT);
    launch_commit(b, list, n_list, s);
    pt.end();
    m->launches += 4;  // stitch, re-walk, commit, mailbox
    {
      MailSrc ms;
      std::memset(&ms, 0, sizeof ms);
      ms.p[0] = b.counters; ms.words[0] = 2; ms.off[0] = 0;
      launch_mail(ms, m->d_mail, s);
    }
    CUDA_TRY(cudaMemsetAsync(b.counters, 0, 2 * sizeof(uint32_t), s));
    CUDA_TRY(cudaStreamSynchronize(s));
    n_list = m->h_mail[0];
    std::swap(cur, nxt);
    list = cur;
  }
  launch_collect_errors(b, s);
  m->launches++;
  return DATOK_OK;
}

struct PieceHead {
  Agg tot;
  unsigned long long err;
  uint32_t invalid;
  WState last;
};

// K3 reduce + scan, then the stream summary, the walk's error key and its final state come back
int do_count(datok_model* m, const WalkBuffers& b, CompactCtx& c, const CompactBuffers& cb, uint32_t flags,
             bool sentence_end_in, PhaseTimer& pt, PieceHead& h) {
  cudaStream_t s = m->stream;
  std::memset(&c, 0, sizeof c);
  c.in = b.in; c.N = b.N; c.n_words = b.n_words;
  c.rstart = b.rstart; c.b_end = b.b_end; c.b_skip = b.b_skip; c.b_sent = b.b_sent; c.b_tend = b.b_tend;
  c.flags = flags; c.err_key = b.err_key;
  c.eot_rewind = m->hm.eot_rewind ? 1u : 0u;
  pt.begin(T_REDUCE);
  launch_compact_reduce(c, cb, s);
  pt.end();
  pt.begin(T_SCAN);
  launch_compact_scan(cb, sentence_end_in, s);
  pt.end();
  m->launches += 4;  // reduce, two scan levels, mailbox
  {
    MailSrc ms;
    static_assert(sizeof(Agg) == 32 && sizeof(WState) % 4 == 0 && sizeof(WState) <= 32, "mailbox layout");
    ms.p[0] = reinterpret_cast<const uint32_t*>(cb.total); ms.words[0] = 8; ms.off[0] = 0;
    ms.p[1] = reinterpret_cast<const uint32_t*>(b.err_key); ms.words[1] = 2; ms.off[1] = 8;
    ms.p[2] = b.counters + 2; ms.words[2] = 1; ms.off[2] = 10;
    ms.p[3] = reinterpret_cast<const uint32_t*>(b.E + (b.n_chunks - 1)); ms.words[3] = sizeof(WState) / 4; ms.off[3] = 12;
    launch_mail(ms, m->d_mail, s);
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  std::memcpy(&h.tot, m->h_mail, sizeof(Agg));
  std::memcpy(&h.err, m->h_mail + 8, sizeof h.err);
  h.invalid = m->h_mail[10];
  std::memcpy(&h.last, m->h_mail + 12, sizeof(WState));
  if (h.err != ~0ull) {  // the walk itself hit a reference panic
    const int code = (int)(h.err & 0xFF);
    g_last_error = std::string("reference would panic: ") + datok_strerror(code);
    return code >= 0xF0 ? DATOK_ERR_CUDA : code;
  }
  return DATOK_OK;
}

void sort_escapes(std::vector<uint32_t>& esc) {
  const size_t n = esc.size() / 2;
  std::vector<uint64_t> key(n);
  for (size_t i = 0; i < n; i++) key[i] = ((uint64_t)esc[2 * i] << 32) | esc[2 * i + 1];
  std::sort(key.begin(), key.end());
  for (size_t i = 0; i < n; i++) { esc[2 * i] = (uint32_t)(key[i] >> 32); esc[2 * i + 1] = (uint32_t)key[i]; }
}

// the walk of a non-final input stopped at the loop top at N: it must be at a rewind point with nothing pending
bool at_text_boundary(const WState& last, uint32_t N) {
  return last.tstart == N && last.base == N && last.eps_state == 0 && !(last.flags & WS_PEND);
}

bool grow_block(Block& b, size_t bytes, bool host) {
  if (b.p && b.bytes >= bytes) return true;
  if (b.p) { if (host) cudaFreeHost(b.p); else cudaFree(b.p); }
  b.p = nullptr;
  b.bytes = bytes + bytes / 4 + 4096;
  b.host = host;
  cudaError_t e = host ? cudaHostAlloc(&b.p, b.bytes, cudaHostAllocDefault) : cudaMalloc(&b.p, b.bytes);
  if (e != cudaSuccess) { b.p = nullptr; b.bytes = 0; cudaGetLastError(); return false; }
  return true;
}

// Host-to-host path for large inputs: the stream is cut after EOT bytes into pieces that are copied in,
// transduced and copied out in a software pipeline (copy engines in both directions overlap the
// kernels).  A piece after an EOT starts a new text: only the walk state and "a token was seen"
// carry over (matrix.go:593-605).  Returns -1 if the input cannot be cut (the caller then runs it in
// one pass).
int run_pipelined(datok_model* m, const uint8_t* in, size_t n, uint32_t flags, const datok_carry* carry_in,
                  datok_result** out) {
  // ---- cut points: the last EOT before each multiple of piece_bytes ----
  // A piece costs the kernels a fixed ~1 ms (fix-up rounds, host round trips) on top of ~7 us per MiB,
  // the copy engine ~20 us per MiB: pieces of >= 96 MiB keep the call copy-bound.  The first piece is a
  // third of that, so that the kernels start early.
  std::vector<size_t> cut;  // piece k = [cut[k], cut[k+1])
  cut.push_back(0);
  size_t piece = m->piece_bytes;
  if (!m->piece_fixed) piece = std::min<size_t>(std::max<size_t>(n / 10, (size_t)64 << 20), (size_t)128 << 20);
  while (n - cut.back() > piece + piece / 2) {
    const size_t lo = cut.back(), sz = (cut.size() == 1 && !m->piece_fixed) ? piece / 3 : piece, want = lo + sz;
    const void* q = memrchr(in + lo, 0x04, want - lo);
    if (!q) return -1;
    const size_t c = (size_t)((const uint8_t*)q - in) + 1;
    if (c - lo < sz / 2) return -1;  // texts longer than half a piece: not worth cutting
    cut.push_back(c);
  }
  cut.push_back(n);
  const size_t np = cut.size() - 1;
  if (np < 2) return -1;

  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard guard;
  CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  std::memset(m->t_ms, 0, sizeof m->t_ms);
  m->launches = 0;
  m->last_rounds = 0;
  const bool call_final = !(flags & DATOK_NOT_FINAL);

  size_t max_piece = 0;
  for (size_t k = 0; k < np; k++) max_piece = std::max(max_piece, cut[k + 1] - cut[k]);
  for (int i = 0; i < 3; i++)
    if (!grow_block(m->d_in[i], max_piece + 64, false)) { g_last_error = "cudaMalloc (input piece) failed"; return DATOK_ERR_CUDA; }
  WalkBuffers b;
  CompactBuffers cb;
  std::memset(&b, 0, sizeof b);
  std::memset(&cb, 0, sizeof cb);
  int rc = ensure_workspace(m, carve(nullptr, (uint32_t)max_piece, m->chunk, false, b, cb));
  if (rc) return rc;

  // DATOK_FORMAT: every piece is formatted on the device from its (piece-relative, absolute-form) arrays, which
  // stay there; the texts of the pieces are copied out one after the other
  const bool fmt = (flags & DATOK_FORMAT) != 0;
  if (fmt) flags &= ~(uint32_t)(DATOK_COMPACT | DATOK_COMPACT8);
  // in the compact modes the token slot 0 holds the deltas (8 or 4 bytes per token); slot 1 is unused by
  // DATOK_COMPACT and holds the escape list of DATOK_COMPACT8
  const bool compact8 = (flags & DATOK_COMPACT8) != 0, compact = compact8 || (flags & DATOK_COMPACT) != 0;
  const size_t tok_rec = compact8 ? 4 : 8;  // bytes per token in slot 0
  const bool want_bytes = compact ? (flags & (DATOK_TOKENS | DATOK_TOKEN_POS)) != 0 : (flags & DATOK_TOKENS) != 0;
  const bool want_pos = !compact && (flags & DATOK_TOKEN_POS) != 0;
  const bool want_spos = (flags & DATOK_SENTENCE_POS) != 0, want_stok = (flags & DATOK_SENTENCES) != 0;
  const bool want[5] = {want_bytes, want_pos || (compact8 && want_bytes), want_spos, want_stok, true};

  // DATOK_PIPE_TRACE=1: per-piece timeline (copy-in, kernels, copy-out; ms since the call started) on stderr
  const bool trace = std::getenv("DATOK_PIPE_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;  // 6 per piece: h2d begin/end, kernels begin/end, d2h begin/end
  if (trace) {
    tev.resize(6 * np);
    for (auto& e : tev) cudaEventCreate(&e);
  }
  // first pieces on their way
  auto issue_h2d = [&](size_t k) -> int {
    const int slot = (int)(k % 3);
    CUDA_TRY(cudaStreamWaitEvent(m->s_h2d, m->ev_free[slot], 0));  // the piece that used this buffer is done
    if (trace) cudaEventRecord(tev[6 * k], m->s_h2d);
    CUDA_TRY(cudaMemcpyAsync(m->d_in[slot].p, in + cut[k], cut[k + 1] - cut[k], cudaMemcpyHostToDevice, m->s_h2d));
    if (trace) cudaEventRecord(tev[6 * k + 1], m->s_h2d);
    CUDA_TRY(cudaEventRecord(m->ev_in[slot], m->s_h2d));
    return DATOK_OK;
  };
  CUDA_TRY(cudaEventRecord(m->ev[0], s));

  if (m->auto_calibrate && !m->calibrated && cut[1] >= (256u << 10)) {
    // one-time specialisation of the table layout to the caller's text
    carve(m->ws, (uint32_t)cut[1], m->chunk, false, b, cb);
    b.in = (const uint8_t*)m->d_in[0].p;
    b.final_input = 0;
    CUDA_TRY(cudaMemcpyAsync(m->d_in[0].p, in, std::min<size_t>(cut[1], 8u << 20), cudaMemcpyHostToDevice, s));
    rc = calibrate_locked(m, b);
    if (rc) return rc;
  }
  for (int i = 0; i < 3; i++) CUDA_TRY(cudaEventRecord(m->ev_free[i], s));
  for (int i = 0; i < 2; i++) CUDA_TRY(cudaEventRecord(m->ev_out[i], s));
  for (size_t k = 0; k < 3 && k < np; k++)
    if ((rc = issue_h2d(k))) return rc;

  datok_result* r = new datok_result();
  r->model = m;
  root_of(m)->live_results++;
  r->device = false;
  std::memset(&r->view, 0, sizeof r->view);
  datok_view& v = r->view;
  Block host[9];  // tok_bytes, tok_pos, sent_pos, sent_tok, text_tok_end, text_sent_end, text_sentpos_end, text_byte_end, text
  uint64_t text_bytes = 0;  // DATOK_FORMAT: text written so far
  struct PieceBase { uint64_t x0, x1, tok, sent, sentpos, byte; };
  std::vector<PieceBase> piece_bases;  // DATOK_FORMAT: the per-text bounds come back piece-relative
  auto fail = [&](int code) {
    cudaStreamSynchronize(m->s_h2d); cudaStreamSynchronize(m->s_d2h); cudaStreamSynchronize(s);
    for (auto& hb : host) release(m, hb);
    free_result_locked(r);
    return code;
  };
  // makes room for `need` bytes in host array i, keeping what is there (rare after the first piece)
  auto host_room = [&](int i, size_t need, size_t used) -> bool {
    if (host[i].p && host[i].bytes >= need) return true;
    int rc2 = DATOK_OK;
    Block nb = acquire(m, need + need / 8, true, &rc2);
    if (!nb.p) return false;
    if (host[i].p) {
      cudaStreamSynchronize(m->s_d2h);
      std::memcpy(nb.p, host[i].p, used);
      release(m, host[i]);
    }
    host[i] = nb;
    return true;
  };

  PhaseTimer pt{m};
  uint32_t state = m->hm.start;
  bool sentence_end_in = false, text_end_in = false;
  if (carry_in) {
    if (carry_in->state) {
      if (carry_in->state > (uint32_t)m->hm.stateCount) { g_last_error = "carry state out of range"; return fail(DATOK_ERR_INVALID_ARG); }
      state = m->hm.new_of_old[carry_in->state];
    }
    sentence_end_in = carry_in->sentence_end != 0;
    text_end_in = carry_in->text_end != 0;
  }
  uint64_t base_tok = 0, base_sent = 0, base_sentpos = 0, base_text = 0, runes = 0;
  uint32_t invalid = 0;
  float ms_kernels = 0;
  std::vector<cudaEvent_t> kev;

  for (size_t k = 0; k < np; k++) {
    const int slot = (int)(k & 1), islot = (int)(k % 3);  // result slot, input buffer
    // the copy of the piece after next goes out now: its buffer was freed by piece k - 1, whose kernels are done
    if (k >= 1 && k + 2 < np && (rc = issue_h2d(k + 2))) return fail(rc);
    const uint32_t N = (uint32_t)(cut[k + 1] - cut[k]);
    const bool last_piece = k + 1 == np;
    const bool final_input = last_piece && call_final;
    uint32_t pflags = flags | (final_input ? 0u : (uint32_t)DATOK_NOT_FINAL);
    if (base_tok > 0) pflags |= DATOK_WRITER_USED;
    carve(m->ws, N, m->chunk, false, b, cb);
    b.in = (const uint8_t*)m->d_in[islot].p;
    b.final_input = final_input ? 1u : 0u;
    CUDA_TRYF(cudaStreamWaitEvent(s, m->ev_in[islot], 0));
    cudaEvent_t k0 = pt.next(), k1 = pt.next();
    CUDA_TRYF(cudaEventRecord(k0, s));
    if (trace) cudaEventRecord(tev[6 * k + 2], s);
    if ((rc = do_walk(m, b, state, pt))) return fail(rc);
    CompactCtx c;
    PieceHead h;
    if ((rc = do_count(m, b, c, cb, pflags, sentence_end_in, pt, h))) { pt.collect(); return fail(rc); }
    if (!final_input && !at_text_boundary(h.last, N)) {
      g_last_error = last_piece ? "DATOK_NOT_FINAL input does not end at a text boundary"
                                : "internal: piece does not end at a text boundary";
      return fail(DATOK_ERR_NOT_AT_BOUNDARY);
    }
    if (fmt && h.invalid) { fail(DATOK_OK); return -2; }  // malformed UTF-8: the single-pass path formats on the host
    // ---- room for this piece's results: device slot and the host arrays ----
    const size_t nt = h.tot.n_tok, ns = (size_t)h.tot.n_sent + 1, nx = (size_t)h.tot.n_text + 1,
                 nsp = (size_t)h.tot.n_sentpos + 1;
    const size_t esc_cap = nt / 16 + 4096;
    const size_t dev_bytes[5] = {nt * tok_rec + 4, compact8 ? esc_cap * 8 : 2 * nt * 4, nsp * 4, ns * 4,
                                 nx * 16 + (nx + 1) * sizeof(DocRec)};
    CUDA_TRYF(cudaStreamWaitEvent(s, m->ev_out[slot], 0));  // the slot's previous results have left the device
    for (int i = 0; i < 5; i++) {
      if (!want[i]) continue;
      if (m->d_out[slot][i].bytes < dev_bytes[i]) CUDA_TRYF(cudaEventSynchronize(m->ev_out[slot]));
      if (!grow_block(m->d_out[slot][i], dev_bytes[i], false)) { g_last_error = "cudaMalloc (results) failed"; return fail(DATOK_ERR_CUDA); }
    }
    {
      // estimate for the whole stream from what has been seen so far (+12 %), exact for the last piece
      const double done = (double)cut[k + 1], scale = last_piece ? 1.0 : 1.12 * (double)n / done;
      auto est = [&](uint64_t have, size_t add) { return (size_t)((double)(have + add) * scale) + 4096; };
      const size_t need8[8] = {est(base_tok, nt) * tok_rec, est(base_tok, nt) * 8, est(base_sentpos, nsp) * 4, est(base_sent, ns) * 4,
                               est(base_text, nx) * 4, est(base_text, nx) * 4, est(base_text, nx) * 4, est(base_text, nx) * 4};
      const size_t used8[8] = {base_tok * tok_rec, base_tok * 8, base_sentpos * 4, base_sent * 4, base_text * 4, base_text * 4,
                               base_text * 4, base_text * 4};
      const bool want8[8] = {want_bytes && !fmt, want_pos && !fmt, want_spos && !fmt, want_stok && !fmt, true, true, true, true};
      for (int i = 0; i < 8; i++)
        if (want8[i] && !host_room(i, need8[i], used8[i])) { g_last_error = "cudaHostAlloc (results) failed"; return fail(DATOK_ERR_CUDA); }
    }
    c.base_tok = (uint32_t)base_tok; c.base_sent = (uint32_t)base_sent; c.base_sentpos = (uint32_t)base_sentpos;
    c.base_byte = (uint32_t)cut[k];
    if (fmt) c.base_tok = c.base_sent = c.base_sentpos = c.base_byte = 0;  // (the formatter reads these arrays: piece-relative)
    c.tok_bytes = (want_bytes && !compact) ? (uint32_t*)m->d_out[slot][0].p : nullptr;
    c.tok_delta = (want_bytes && compact && !compact8) ? (uint16_t*)m->d_out[slot][0].p : nullptr;
    c.tok_delta8 = (want_bytes && compact8) ? (uint8_t*)m->d_out[slot][0].p : nullptr;
    c.esc = (want_bytes && compact8) ? (uint32_t*)m->d_out[slot][1].p : nullptr;
    c.esc_count = b.counters + 3; c.esc_cap = (uint32_t)esc_cap;
    c.tok_pos = want_pos ? (int32_t*)m->d_out[slot][1].p : nullptr;
    c.sent_pos = want_spos ? (int32_t*)m->d_out[slot][2].p : nullptr;
    c.sent_tok = want_stok ? (uint32_t*)m->d_out[slot][3].p : nullptr;
    c.text_tok_end = (uint32_t*)m->d_out[slot][4].p;
    c.text_sent_end = c.text_tok_end + nx;
    c.text_sentpos_end = c.text_tok_end + 2 * nx;
    c.text_byte_end = c.text_tok_end + 3 * nx;
    c.docs = reinterpret_cast<DocRec*>(c.text_tok_end + 4 * nx);
    pt.begin(T_TEXTS);
    launch_compact_texts(c, cb, s);
    pt.end();
    pt.begin(T_EMIT);
    {
      const int e = launch_compact_emit(c, cb, s);
      if (e != 0) { g_last_error = std::string("compact_emit launch: ") + cudaGetErrorString((cudaError_t)e); return fail(DATOK_ERR_CUDA); }
    }
    launch_compact_finalize(c, cb, text_end_in, final_input, s);
    pt.end();
    m->launches += 4;  // texts, emit, finalize, mailbox
    struct { StreamTotals fin; unsigned long long err; } tail;
    FmtCtx f;
    std::memset(&f, 0, sizeof f);
    if (fmt) {
      // the prefix sums of the formatter right behind the emit pass (the counts are those of the reduce pass plus
      // the end-of-stream events, which only the last piece has: it is sized with them)
      f.in = b.in;
      f.tok_bytes = c.tok_bytes; f.tok_pos = c.tok_pos; f.sent_pos = c.sent_pos; f.sent_tok = c.sent_tok;
      f.text_tok_end = c.text_tok_end; f.text_sent_end = c.text_sent_end; f.text_sentpos_end = c.text_sentpos_end;
      f.flags = flags & 15u;
    }
    {
      MailSrc ms;
      std::memset(&ms, 0, sizeof ms);
      ms.p[0] = reinterpret_cast<const uint32_t*>(cb.total + 1); ms.words[0] = 8; ms.off[0] = 0;
      ms.p[1] = reinterpret_cast<const uint32_t*>(b.err_key); ms.words[1] = 2; ms.off[1] = 8;
      ms.p[2] = b.counters + 3; ms.words[2] = 1; ms.off[2] = 10;
      launch_mail(ms, m->d_mail, s);
    }
    if (!fmt) CUDA_TRYF(cudaEventRecord(k1, s));
    if (trace) cudaEventRecord(tev[6 * k + 3], s);
    if (!fmt) {
      CUDA_TRYF(cudaEventRecord(m->ev_emit[slot], s));
      CUDA_TRYF(cudaEventRecord(m->ev_free[islot], s));
    }
    CUDA_TRYF(cudaStreamSynchronize(s));
    std::memcpy(&tail.fin, m->h_mail, sizeof(StreamTotals));
    std::memcpy(&tail.err, m->h_mail + 8, sizeof tail.err);
    kev.push_back(k0); kev.push_back(k1);
    if (tail.err != ~0ull) {
      pt.collect();
      const int code = (int)(tail.err & 0xFF);
      g_last_error = std::string("reference would panic: ") + datok_strerror(code);
      return fail(code);
    }
    if (compact8 && want_bytes && m->h_mail[10] > esc_cap) {
      // more escape pairs than the list holds: the emit pass runs again with a list of the counted size
      const size_t ne = m->h_mail[10];
      if (!grow_block(m->d_out[slot][1], ne * 8, false)) { g_last_error = "cudaMalloc (escape list) failed"; return fail(DATOK_ERR_CUDA); }
      c.esc = (uint32_t*)m->d_out[slot][1].p; c.esc_cap = (uint32_t)ne;
      CUDA_TRYF(cudaMemsetAsync(b.counters + 3, 0, sizeof(uint32_t), s));
      {
        const int e = launch_compact_emit(c, cb, s);
        if (e != 0) { g_last_error = std::string("compact_emit launch: ") + cudaGetErrorString((cudaError_t)e); return fail(DATOK_ERR_CUDA); }
      }
      launch_compact_finalize(c, cb, text_end_in, final_input, s);
      m->launches += 2;
      CUDA_TRYF(cudaEventRecord(m->ev_emit[slot], s));
      CUDA_TRYF(cudaEventRecord(m->ev_free[islot], s));
      CUDA_TRYF(cudaStreamSynchronize(s));
    }
    if (compact8 && want_bytes && m->h_mail[10]) {  // rare: this piece's escape pairs (a plain, blocking copy)
      const size_t ne = m->h_mail[10], at = r->esc.size();
      r->esc.resize(at + 2 * ne);
      CUDA_TRYF(cudaMemcpy(r->esc.data() + at, m->d_out[slot][1].p, ne * 8, cudaMemcpyDeviceToHost));
    }
    unsigned long long piece_text = 0;
    if (fmt) {
      // ---- the piece's text: prefix sums, its length, room for it, the writers ----
      f.n_tok = tail.fin.n_tok; f.n_sent = tail.fin.n_sent; f.n_sentpos = tail.fin.n_sentpos; f.n_text = tail.fin.n_text;
      if (!grow_block(m->d_out[slot][5], format_scratch_bytes(f.n_tok, f.n_sentpos, f.n_text), false)) {
        g_last_error = "cudaMalloc (formatter) failed";
        return fail(DATOK_ERR_CUDA);
      }
      format_carve(f, (uint8_t*)m->d_out[slot][5].p);
      unsigned long long* d_total = reinterpret_cast<unsigned long long*>(b.counters + 4);
      int e = launch_format_scan(f, d_total, s);
      MailSrc ms;
      std::memset(&ms, 0, sizeof ms);
      ms.p[0] = reinterpret_cast<const uint32_t*>(d_total); ms.words[0] = 2; ms.off[0] = 0;
      launch_mail(ms, m->d_mail, s);
      m->launches += 10;
      if (e == 0) e = (int)cudaStreamSynchronize(s);
      if (e != 0) { g_last_error = std::string("device formatter: ") + cudaGetErrorString((cudaError_t)e); return fail(DATOK_ERR_CUDA); }
      std::memcpy(&piece_text, m->h_mail, sizeof piece_text);
      if (!grow_block(m->d_out[slot][6], (size_t)piece_text + 16, false)) { g_last_error = "cudaMalloc (text) failed"; return fail(DATOK_ERR_CUDA); }
      {
        const double done = (double)cut[k + 1], scale = last_piece ? 1.0 : 1.12 * (double)n / done;
        const size_t need = (size_t)((double)(text_bytes + piece_text) * scale) + 4096;
        if (!host_room(8, need, (size_t)text_bytes)) { g_last_error = "cudaHostAlloc (text) failed"; return fail(DATOK_ERR_CUDA); }
      }
      f.out = (uint8_t*)m->d_out[slot][6].p;
      e = launch_format_write(f, s);
      m->launches += 2;
      if (e != 0) { g_last_error = std::string("device formatter: ") + cudaGetErrorString((cudaError_t)e); return fail(DATOK_ERR_CUDA); }
      CUDA_TRYF(cudaEventRecord(k1, s));
      CUDA_TRYF(cudaEventRecord(m->ev_emit[slot], s));
      CUDA_TRYF(cudaEventRecord(m->ev_free[islot], s));
      piece_bases.push_back({base_text, base_text + tail.fin.n_text, base_tok, base_sent, base_sentpos, (uint64_t)cut[k]});
    }
    // ---- results out, behind the kernels of the following pieces ----
    CUDA_TRYF(cudaStreamWaitEvent(m->s_d2h, m->ev_emit[slot], 0));
    if (trace) cudaEventRecord(tev[6 * k + 4], m->s_d2h);
    if (fmt && piece_text)
      CUDA_TRYF(cudaMemcpyAsync((uint8_t*)host[8].p + text_bytes, m->d_out[slot][6].p, (size_t)piece_text, cudaMemcpyDeviceToHost, m->s_d2h));
    text_bytes += piece_text;
    const uint32_t* dtx = (const uint32_t*)m->d_out[slot][4].p;
    struct Cp { int hi; const void* src; size_t off, bytes; };
    const Cp cps[8] = {{0, m->d_out[slot][0].p, base_tok * tok_rec, (size_t)tail.fin.n_tok * tok_rec},
                       {1, m->d_out[slot][1].p, base_tok * 8, (size_t)tail.fin.n_tok * 8},
                       {2, m->d_out[slot][2].p, base_sentpos * 4, (size_t)tail.fin.n_sentpos * 4},
                       {3, m->d_out[slot][3].p, base_sent * 4, (size_t)tail.fin.n_sent * 4},
                       {4, dtx, base_text * 4, (size_t)tail.fin.n_text * 4},
                       {5, dtx + nx, base_text * 4, (size_t)tail.fin.n_text * 4},
                       {6, dtx + 2 * nx, base_text * 4, (size_t)tail.fin.n_text * 4},
                       {7, dtx + 3 * nx, base_text * 4, (size_t)tail.fin.n_text * 4}};
    const bool want8[8] = {want_bytes && !fmt, want_pos && !fmt, want_spos && !fmt, want_stok && !fmt, true, true, true, true};
    for (const Cp& cp : cps)
      if (want8[cp.hi] && cp.bytes)
        CUDA_TRYF(cudaMemcpyAsync((uint8_t*)host[cp.hi].p + cp.off, cp.src, cp.bytes, cudaMemcpyDeviceToHost, m->s_d2h));
    if (trace) cudaEventRecord(tev[6 * k + 5], m->s_d2h);
    CUDA_TRYF(cudaEventRecord(m->ev_out[slot], m->s_d2h));
    // ---- carry to the next piece ----
    base_tok += tail.fin.n_tok; base_sent += tail.fin.n_sent; base_sentpos += tail.fin.n_sentpos;
    base_text += tail.fin.n_text; runes += tail.fin.n_rune;
    invalid |= h.invalid;
    state = h.last.t;
    v.carry_out.state = m->hm.old_of_new[h.last.t];
    v.carry_out.sentence_end = 1;
    v.carry_out.text_end = 1;
    if (!final_input) {
      const uint32_t lk = tail.fin.last_kind;
      v.carry_out.sentence_end = (lk == EV_SENT || lk == EV_TEND) ? 1u : 0u;
      v.carry_out.text_end = (tail.fin.n_text > 0 && tail.fin.tokless) ? 1u : (text_end_in ? 1u : 0u);
    }
    sentence_end_in = v.carry_out.sentence_end != 0;
    text_end_in = v.carry_out.text_end != 0;
  }
  CUDA_TRYF(cudaStreamSynchronize(m->s_d2h));
  CUDA_TRYF(cudaEventRecord(m->ev[3], s));
  CUDA_TRYF(cudaStreamSynchronize(s));
  pt.collect();
  if (trace) {
    for (size_t k = 0; k < np; k++) {
      float t[6];
      for (int j = 0; j < 6; j++) cudaEventElapsedTime(&t[j], m->ev[0], tev[6 * k + j]);
      std::fprintf(stderr, "piece %2zu %9zu B  h2d %7.3f-%7.3f  kernels %7.3f-%7.3f  d2h %7.3f-%7.3f\n", k, cut[k + 1] - cut[k],
                   t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    for (auto& e : tev) cudaEventDestroy(e);
  }
  for (size_t i = 0; i + 1 < kev.size(); i += 2) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, kev[i], kev[i + 1]) == cudaSuccess) ms_kernels += ms;
  }
  v.n_tokens = base_tok; v.n_sentences = base_sent; v.n_texts = base_text; v.n_sent_pos = base_sentpos; v.n_runes = runes;
  v.has_invalid_utf8 = invalid;
  if (fmt) {
    // the per-text bounds of a formatted result came back relative to their piece
    uint32_t* te[4] = {(uint32_t*)host[4].p, (uint32_t*)host[5].p, (uint32_t*)host[6].p, (uint32_t*)host[7].p};
    for (const PieceBase& pb : piece_bases)
      for (uint64_t d = pb.x0; d < pb.x1; d++) {
        te[0][d] += (uint32_t)pb.tok; te[1][d] += (uint32_t)pb.sent; te[2][d] += (uint32_t)pb.sentpos; te[3][d] += (uint32_t)pb.byte;
      }
    v.text = (const uint8_t*)host[8].p;
    v.text_len = text_bytes;
  }
  v.tok_bytes = (want_bytes && !compact && !fmt) ? (const uint32_t*)host[0].p : nullptr;
  v.tok_delta = (want_bytes && compact && !compact8) ? (const uint16_t*)host[0].p : nullptr;
  v.tok_delta8 = (want_bytes && compact8) ? (const uint8_t*)host[0].p : nullptr;
  if (v.tok_delta8) {
    sort_escapes(r->esc);
    v.tok_esc = r->esc.data();
    v.n_esc = r->esc.size() / 2;
  }
  v.tok_pos = (want_pos && !fmt) ? (const int32_t*)host[1].p : nullptr;
  v.sent_pos = (want_spos && !fmt) ? (const int32_t*)host[2].p : nullptr;
  v.sent_tok = (want_stok && !fmt) ? (const uint32_t*)host[3].p : nullptr;
  v.text_tok_end = (const uint32_t*)host[4].p;
  v.text_sent_end = (const uint32_t*)host[5].p;
  v.text_sentpos_end = (const uint32_t*)host[6].p;
  v.text_byte_end = (const uint32_t*)host[7].p;
  // overlapped: only the wall time of the whole call and the summed kernel time are meaningful
  float total_ms = 0;
  cudaEventElapsedTime(&total_ms, m->ev[0], m->ev[3]);
  v.ms_kernels = ms_kernels;
  v.ms_h2d = 0;
  v.ms_d2h = total_ms > ms_kernels ? total_ms - ms_kernels : 0;
  for (auto& hb : host) if (hb.p) r->blocks.push_back(hb);
  *out = r;
  return DATOK_OK;
}


int run_pipeline(datok_model* m, const uint8_t* in, bool in_is_device, size_t n, uint32_t flags,
                 const datok_carry* carry_in, bool device_out, datok_result** out) {
  if (!m || !out || (!in && n)) { g_last_error = "invalid argument"; return DATOK_ERR_INVALID_ARG; }
  if (n >= 0xFFFFFFFFull - (1u << 20)) { g_last_error = "input too large for one call"; return DATOK_ERR_TOO_LARGE; }
  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard guard;
  CUDA_TRY(cudaSetDevice(m->device));
  const uint32_t N = (uint32_t)n;
  cudaStream_t s = m->stream;
  std::memset(m->t_ms, 0, sizeof m->t_ms);
  m->launches = 0;
  m->last_rounds = 0;

  uint32_t start_state = m->hm.start;
  bool sentence_end_in = false, text_end_in = false;
  if (carry_in) {
    if (carry_in->state) {
      if (carry_in->state > (uint32_t)m->hm.stateCount) { g_last_error = "carry state out of range"; return DATOK_ERR_INVALID_ARG; }
      start_state = m->hm.new_of_old[carry_in->state];
    }
    sentence_end_in = carry_in->sentence_end != 0;
    text_end_in = carry_in->text_end != 0;
  }

  WalkBuffers b;
  CompactBuffers cb;
  std::memset(&b, 0, sizeof b);
  std::memset(&cb, 0, sizeof cb);
  const size_t need = carve(nullptr, N, m->chunk, !in_is_device, b, cb);
  int rc = ensure_workspace(m, need);
  if (rc) return rc;
  carve(m->ws, N, m->chunk, !in_is_device, b, cb);
  if (in_is_device) b.in = in;
  const bool final_input = !(flags & DATOK_NOT_FINAL);
  b.final_input = final_input ? 1u : 0u;

  if (m->auto_calibrate && !m->calibrated && N >= (256u << 10)) {
    // one-time specialisation of the table layout to the caller's text
    if (!in_is_device) CUDA_TRY(cudaMemcpyAsync(const_cast<uint8_t*>(b.in), in, std::min<uint32_t>(N, 8u << 20),
                                                cudaMemcpyHostToDevice, s));
    rc = calibrate_locked(m, b);
    if (rc) return rc;
    if (carry_in && carry_in->state) start_state = m->hm.new_of_old[carry_in->state];
    else start_state = m->hm.start;
  }

  PhaseTimer pt{m};
  CUDA_TRY(cudaEventRecord(m->ev[0], s));
  if (!in_is_device && N) CUDA_TRY(cudaMemcpyAsync(const_cast<uint8_t*>(b.in), in, N, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaEventRecord(m->ev[1], s));

  if ((rc = do_walk(m, b, start_state, pt))) return rc;
  CompactCtx c;
  PieceHead hdr;
  if ((rc = do_count(m, b, c, cb, flags, sentence_end_in, pt, hdr))) { pt.collect(); return rc; }

  // output arrays: sized from the scan totals (+1 for the end-of-stream events)
  datok_result* r = new datok_result();
  r->model = m;
  root_of(m)->live_results++;
  r->device = device_out;
  std::memset(&r->view, 0, sizeof r->view);
  const size_t nt = hdr.tot.n_tok, ns = (size_t)hdr.tot.n_sent + 1, nx = (size_t)hdr.tot.n_text + 1,
               np = (size_t)hdr.tot.n_sentpos + 1;
  // DATOK_FORMAT: the device formats the text from the absolute arrays, which then stay on the device (with
  // malformed UTF-8 in the input the host formatter does it, from the arrays: surfaces are re-encoded)
  const bool fmt = (flags & DATOK_FORMAT) != 0, fmt_dev = fmt && !hdr.invalid;
  if (fmt) flags &= ~(uint32_t)(DATOK_COMPACT | DATOK_COMPACT8);
  const bool compact8 = (flags & DATOK_COMPACT8) != 0, compact = compact8 || (flags & DATOK_COMPACT) != 0;
  const bool want_tok = (flags & (DATOK_TOKENS | DATOK_TOKEN_POS)) != 0;
  const bool want_delta = compact && !compact8 && want_tok, want_delta8 = compact8 && want_tok;
  const bool want_bytes = !compact && (flags & DATOK_TOKENS) != 0, want_pos = !compact && (flags & DATOK_TOKEN_POS) != 0;
  const size_t esc_cap = nt / 16 + 4096;
  const bool want_spos = (flags & DATOK_SENTENCE_POS) != 0, want_stok = (flags & DATOK_SENTENCES) != 0;
  struct Out { size_t bytes; bool want; void** dev; Block d, h; };
  void *d_tok_bytes = nullptr, *d_tok_pos = nullptr, *d_sent_pos = nullptr, *d_sent_tok = nullptr, *d_text = nullptr,
       *d_delta = nullptr, *d_delta8 = nullptr, *d_esc = nullptr;
  // the per-text arrays are followed by the DocRec table (device scratch, not copied back)
  Out outs[8] = {{2 * nt * 4, want_bytes, &d_tok_bytes, {}, {}}, {2 * nt * 4, want_pos, &d_tok_pos, {}, {}},
                 {np * 4, want_spos, &d_sent_pos, {}, {}},       {ns * 4, want_stok, &d_sent_tok, {}, {}},
                 {nx * 4 * 4 + (nx + 1) * sizeof(DocRec), true, &d_text, {}, {}},
                 {nt * 8, want_delta, &d_delta, {}, {}},         {nt * 4 + 4, want_delta8, &d_delta8, {}, {}},
                 {esc_cap * 8, want_delta8, &d_esc, {}, {}}};
  for (auto& o : outs) {
    if (!o.want) continue;
    o.d = acquire(m, o.bytes, false, &rc);
    const bool to_host = !device_out && (!fmt_dev || o.dev == &d_text);  // (formatted on the device: only the per-text bounds leave it)
    if (to_host) o.h = acquire(m, o.bytes, true, &rc);
    r->blocks.push_back(o.d);
    if (to_host) r->blocks.push_back(o.h);
    *o.dev = o.d.p;
  }
  if (rc) { free_result_locked(r); return rc; }
  c.tok_bytes = (uint32_t*)d_tok_bytes;
  c.tok_pos = (int32_t*)d_tok_pos;
  c.tok_delta = (uint16_t*)d_delta;
  c.tok_delta8 = (uint8_t*)d_delta8;
  c.esc = (uint32_t*)d_esc; c.esc_count = b.counters + 3; c.esc_cap = (uint32_t)esc_cap;
  c.sent_pos = (int32_t*)d_sent_pos;
  c.sent_tok = (uint32_t*)d_sent_tok;
  c.text_tok_end = (uint32_t*)d_text;
  c.text_sent_end = c.text_tok_end + nx;
  c.text_sentpos_end = c.text_tok_end + 2 * nx;
  c.text_byte_end = c.text_tok_end + 3 * nx;
  c.docs = reinterpret_cast<DocRec*>(c.text_tok_end + 4 * nx);
  pt.begin(T_TEXTS);
  launch_compact_texts(c, cb, s);
  pt.end();
  pt.begin(T_EMIT);
  {
    const int e = launch_compact_emit(c, cb, s);
    if (e != 0) { g_last_error = std::string("compact_emit launch: ") + cudaGetErrorString((cudaError_t)e); free_result_locked(r); return DATOK_ERR_CUDA; }
  }
  launch_compact_finalize(c, cb, text_end_in, final_input, s);
  pt.end();
  m->launches += 4;  // texts, emit, finalize, mailbox
  struct { StreamTotals fin; unsigned long long err; WState last; uint32_t n_esc; } tail;
  tail.last = hdr.last;
  {
    MailSrc ms;
    std::memset(&ms, 0, sizeof ms);
    ms.p[0] = reinterpret_cast<const uint32_t*>(cb.total + 1); ms.words[0] = 8; ms.off[0] = 0;
    ms.p[1] = reinterpret_cast<const uint32_t*>(b.err_key); ms.words[1] = 2; ms.off[1] = 8;
    ms.p[2] = b.counters + 3; ms.words[2] = 1; ms.off[2] = 10;
    launch_mail(ms, m->d_mail, s);
  }
  CUDA_TRY(cudaEventRecord(m->ev[2], s));
  CUDA_TRY(cudaStreamSynchronize(s));
  std::memcpy(&tail.fin, m->h_mail, sizeof(StreamTotals));
  std::memcpy(&tail.err, m->h_mail + 8, sizeof tail.err);
  tail.n_esc = m->h_mail[10];
  if (want_delta8 && tail.err == ~0ull && tail.n_esc > esc_cap) {
    // more escape pairs than the list holds (many long tokens or gaps): the pass is repeated once with a
    // list of the counted size -- it writes the same arrays again
    Block big = acquire(m, (size_t)tail.n_esc * 8, false, &rc);
    if (rc) { pt.collect(); free_result_locked(r); return rc; }
    r->blocks.push_back(big);
    if (!device_out) {
      Block bigh = acquire(m, (size_t)tail.n_esc * 8, true, &rc);
      if (rc) { pt.collect(); free_result_locked(r); return rc; }
      r->blocks.push_back(bigh);
      outs[7].h = bigh;
    }
    outs[7].d = big;
    d_esc = big.p;
    c.esc = (uint32_t*)d_esc; c.esc_cap = tail.n_esc;
    CUDA_TRY(cudaMemsetAsync(b.counters + 3, 0, sizeof(uint32_t), s));
    {
      const int e = launch_compact_emit(c, cb, s);
      if (e != 0) { g_last_error = std::string("compact_emit launch: ") + cudaGetErrorString((cudaError_t)e); free_result_locked(r); return DATOK_ERR_CUDA; }
    }
    launch_compact_finalize(c, cb, text_end_in, final_input, s);
    m->launches += 2;
    CUDA_TRY(cudaEventRecord(m->ev[2], s));
    CUDA_TRY(cudaStreamSynchronize(s));
  }
  pt.collect();
  if (tail.err != ~0ull) {
    const int code = (int)(tail.err & 0xFF);
    g_last_error = std::string("reference would panic: ") + datok_strerror(code);
    free_result_locked(r);
    return code;
  }
  datok_view& v = r->view;
  v.n_tokens = tail.fin.n_tok;
  v.n_sentences = tail.fin.n_sent;
  v.n_texts = tail.fin.n_text;
  v.n_sent_pos = tail.fin.n_sentpos;
  v.n_runes = tail.fin.n_rune;
  v.has_invalid_utf8 = hdr.invalid;
  v.carry_out.state = m->hm.old_of_new[tail.last.t];
  v.carry_out.sentence_end = 1;
  v.carry_out.text_end = 1;
  if (!final_input) {
    // the walk stopped at the loop top at N: it must be at a rewind point with nothing pending
    if (!at_text_boundary(tail.last, N)) {
      g_last_error = "DATOK_NOT_FINAL input does not end at a text boundary";
      free_result_locked(r);
      return DATOK_ERR_NOT_AT_BOUNDARY;
    }
    const uint32_t lk = tail.fin.last_kind;
    v.carry_out.sentence_end = (lk == EV_SENT || lk == EV_TEND) ? 1u : 0u;
    v.carry_out.text_end = (tail.fin.n_text > 0 && tail.fin.tokless) ? 1u : (text_end_in ? 1u : 0u);
  }
  // ---- the text, on the device ----
  Block text_d, text_h;
  unsigned long long text_len = 0;
  if (fmt_dev) {
    FmtCtx f;
    std::memset(&f, 0, sizeof f);
    f.in = b.in;
    f.tok_bytes = c.tok_bytes; f.tok_pos = c.tok_pos; f.sent_pos = c.sent_pos; f.sent_tok = c.sent_tok;
    f.text_tok_end = c.text_tok_end; f.text_sent_end = c.text_sent_end; f.text_sentpos_end = c.text_sentpos_end;
    f.n_tok = tail.fin.n_tok; f.n_sent = tail.fin.n_sent; f.n_sentpos = tail.fin.n_sentpos; f.n_text = tail.fin.n_text;
    f.flags = flags & 15u;
    Block scratch = acquire(m, format_scratch_bytes(f.n_tok, f.n_sentpos, f.n_text), false, &rc);
    if (rc) { free_result_locked(r); return rc; }
    r->blocks.push_back(scratch);
    format_carve(f, (uint8_t*)scratch.p);
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(b.counters + 4);
    int e = launch_format_scan(f, d_total, s);
    MailSrc ms;
    std::memset(&ms, 0, sizeof ms);
    ms.p[0] = reinterpret_cast<const uint32_t*>(d_total); ms.words[0] = 2; ms.off[0] = 0;
    launch_mail(ms, m->d_mail, s);
    m->launches += 10;
    if (e == 0) e = (int)cudaStreamSynchronize(s);
    if (e != 0) { g_last_error = std::string("device formatter: ") + cudaGetErrorString((cudaError_t)e); free_result_locked(r); return DATOK_ERR_CUDA; }
    std::memcpy(&text_len, m->h_mail, sizeof text_len);
    text_d = acquire(m, (size_t)text_len + 16, false, &rc);
    if (!device_out) text_h = acquire(m, (size_t)text_len + 16, true, &rc);
    if (rc) { free_result_locked(r); return rc; }
    r->blocks.push_back(text_d);
    if (!device_out) r->blocks.push_back(text_h);
    f.out = (uint8_t*)text_d.p;
    e = launch_format_write(f, s);
    m->launches += 2;
    if (e != 0) { g_last_error = std::string("device formatter: ") + cudaGetErrorString((cudaError_t)e); free_result_locked(r); return DATOK_ERR_CUDA; }
    CUDA_TRY(cudaEventRecord(m->ev[2], s));
    if (!device_out && text_len) CUDA_TRY(cudaMemcpyAsync(text_h.p, text_d.p, (size_t)text_len, cudaMemcpyDeviceToHost, s));
    v.text = (const uint8_t*)(device_out ? text_d.p : text_h.p);
    v.text_len = text_len;
  }
  // ---- D2H ----
  for (auto& o : outs) {
    if (!o.want || device_out || !o.h.p) continue;
    size_t bytes = o.bytes;
    if (o.dev == &d_tok_bytes || o.dev == &d_tok_pos || o.dev == &d_delta) bytes = 2 * (size_t)v.n_tokens * 4;
    else if (o.dev == &d_delta8) bytes = (size_t)v.n_tokens * 4;
    else if (o.dev == &d_esc) bytes = (size_t)tail.n_esc * 8;
    else if (o.dev == &d_sent_pos) bytes = (size_t)v.n_sent_pos * 4;
    else if (o.dev == &d_sent_tok) bytes = (size_t)v.n_sentences * 4;
    else if (o.dev == &d_text) bytes = nx * 4 * 4;
    if (bytes) CUDA_TRY(cudaMemcpyAsync(o.h.p, o.d.p, bytes, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaEventRecord(m->ev[3], s));
  CUDA_TRY(cudaStreamSynchronize(s));
  auto pick = [&](int i) -> void* { return outs[i].want ? (device_out ? (fmt_dev && i != 4 ? nullptr : outs[i].d.p) : outs[i].h.p) : nullptr; };
  v.tok_bytes = (const uint32_t*)pick(0);
  v.tok_pos = (const int32_t*)pick(1);
  v.sent_pos = (const int32_t*)pick(2);
  v.sent_tok = (const uint32_t*)pick(3);
  v.tok_delta = (const uint16_t*)pick(5);
  v.tok_delta8 = (const uint8_t*)pick(6);
  if (want_delta8 && !device_out) {  // the escape list is tiny: sorted copy owned by the result
    const uint32_t* e = (const uint32_t*)outs[7].h.p;
    r->esc.assign(e, e + 2 * (size_t)tail.n_esc);
    sort_escapes(r->esc);
    v.tok_esc = r->esc.data();
    v.n_esc = tail.n_esc;
  } else if (want_delta8) {
    v.tok_esc = (const uint32_t*)outs[7].d.p;  // device-resident and unordered
    v.n_esc = tail.n_esc;
  }
  const uint32_t* tx = (const uint32_t*)pick(4);
  v.text_tok_end = tx;
  v.text_sent_end = tx + nx;
  v.text_sentpos_end = tx + 2 * nx;
  v.text_byte_end = tx + 3 * nx;
  cudaEventElapsedTime(&v.ms_h2d, m->ev[0], m->ev[1]);
  cudaEventElapsedTime(&v.ms_kernels, m->ev[1], m->ev[2]);
  cudaEventElapsedTime(&v.ms_d2h, m->ev[2], m->ev[3]);
  if (fmt && !fmt_dev && !device_out) {
    // malformed UTF-8: Go re-encodes those bytes (U+FFFD), surfaces are not verbatim copies -> the host formatter
    const size_t need = datok_format(r, in, n, flags & 15u, nullptr, 0);
    if (need == (size_t)-1) { g_last_error = "formatter: arrays missing"; free_result_locked(r); return DATOK_ERR_INVALID_ARG; }
    Block th = acquire(m, need + 16, true, &rc);
    if (rc) { free_result_locked(r); return rc; }
    r->blocks.push_back(th);
    datok_format(r, in, n, flags & 15u, (uint8_t*)th.p, need);
    v.text = (const uint8_t*)th.p;
    v.text_len = need;
  }
  if (!device_out) {  // the device copies are no longer needed
    std::vector<Block> keep;
    for (auto& blk : r->blocks) { if (blk.host) keep.push_back(blk); else release(m, blk); }
    r->blocks.swap(keep);
  }
  *out = r;
  return DATOK_OK;
}

}  // namespace

extern "C" {

datok_model* datok_load(const char* path, int device, int* err) {
  int dummy;
  if (!err) err = &dummy;
  datok_model* m = new datok_model();
  std::string why;
  int rc = load_matok_file(path, m->hm, why);
  if (rc) { g_last_error = why; *err = rc; delete m; return nullptr; }
  return finish_load(m, device, err);
}

datok_model* datok_load_image(const uint8_t* image, size_t n, int device, int* err) {
  int dummy;
  if (!err) err = &dummy;
  datok_model* m = new datok_model();
  std::string why;
  int rc = parse_matok_image(image, n, m->hm, why);
  if (!rc) rc = build_layout(m->hm, why);
  if (rc) { g_last_error = why; *err = rc; delete m; return nullptr; }
  return finish_load(m, device, err);
}

datok_model* datok_load_foma(const char* path, int device, int* err) {
  int dummy;
  if (!err) err = &dummy;
  datok_model* m = new datok_model();
  std::string why;
  int rc = load_foma_file(path, m->hm, why);
  if (rc) { g_last_error = why; *err = rc; delete m; return nullptr; }
  return finish_load(m, device, err);
}

int datok_compile_foma(const char* foma_path, const char* matok_path) {
  HostModel hm;
  std::string why;
  int rc = load_foma_file(foma_path, hm, why);  // (with the layout build: a model the kernels cannot run is reported here)
  if (!rc) rc = save_matok_file(hm, matok_path, why);
  if (rc) g_last_error = why;
  return rc;
}

int datok_save(const datok_model* m, const char* path) {
  if (!m || !path) return DATOK_ERR_INVALID_ARG;
  std::string why;
  int rc = save_matok_file(m->hm, path, why);
  if (rc) g_last_error = why;
  return rc;
}

size_t datok_write_image(const datok_model* m, uint8_t* dst, size_t cap) {
  if (!m) return 0;
  std::vector<uint8_t> img;
  std::string why;
  if (write_matok_image(m->hm, img, why)) { g_last_error = why; return 0; }
  if (dst && cap >= img.size()) std::memcpy(dst, img.data(), img.size());
  return img.size();
}

void datok_free(datok_model* m) {
  if (!m) return;
  m->freed_by_user = true;
  if (m->live_results == 0) destroy_model(m);  // otherwise the last datok_result_free() does it
}

const char* datok_type(void) { return "MATOK"; }
const char* datok_model_type(const datok_model* m) { return (m && !m->hm.eot_rewind) ? "DATOK" : "MATOK"; }

int datok_model_info(const datok_model* m, uint32_t* state_count, uint32_t* sigma_count, uint32_t* n_classes,
                     uint32_t* epsilon, uint32_t* unknown, uint32_t* identity) {
  if (!m) return DATOK_ERR_INVALID_ARG;
  if (state_count) *state_count = (uint32_t)m->hm.stateCount;
  if (sigma_count) *sigma_count = (uint32_t)m->hm.sigmaCount;
  if (n_classes) *n_classes = m->hm.n_classes;
  if (epsilon) *epsilon = (uint32_t)m->hm.epsilon;
  if (unknown) *unknown = (uint32_t)m->hm.unknown;
  if (identity) *identity = (uint32_t)m->hm.identity;
  return DATOK_OK;
}

// An idle execution context of the model for one call (released by ctx_release).
static datok_model* ctx_acquire(datok_model* P) {
  std::lock_guard<std::mutex> lock(P->ctx_mu);
  if (!P->busy) { P->busy = true; return P; }
  for (datok_model* c : P->clones)
    if (!c->busy) { c->busy = true; return c; }
  // a new context only once the table layout is final (the one-time calibration re-uploads the tables)
  if ((P->calibrated || !P->auto_calibrate) && (int)P->clones.size() + 1 < P->max_ctx) {
    if (datok_model* c = clone_context(P)) { c->busy = true; P->clones.push_back(c); return c; }
  }
  return nullptr;  // every context is busy: the call queues on the primary
}
static void ctx_release(datok_model* P, datok_model* c) {
  std::lock_guard<std::mutex> lock(P->ctx_mu);
  c->busy = false;
}

// A double-array model (datok.go) does not rewind its buffer at an EOT (datok.go:1019-1030): a text's first Token call
// reaches back into the text before, so a stream of such a model cannot be cut at an EOT (no DATOK_NOT_FINAL, no
// pieces), and the delta-coded transport forms, whose cursors restart at every text, do not apply: absolute arrays.
static int check_norewind(const datok_model* m, uint32_t& flags) {
  if (!m || m->hm.eot_rewind) return DATOK_OK;
  if (flags & DATOK_NOT_FINAL) {
    g_last_error = "double-array model (.datok): its stream cannot be continued behind an EOT (DATOK_NOT_FINAL)";
    return DATOK_ERR_UNSUPPORTED_MODEL;
  }
  flags &= ~(uint32_t)(DATOK_COMPACT | DATOK_COMPACT8);
  return DATOK_OK;
}

int datok_transduce(datok_model* m, const uint8_t* in, size_t n, uint32_t flags, const datok_carry* carry_in,
                    datok_result** out) {
  if (const int rc = check_norewind(m, flags)) return rc;
  if (!m) { g_last_error = "invalid argument"; return DATOK_ERR_INVALID_ARG; }
  datok_model* x = ctx_acquire(m);   // (nullptr: all busy -- queue on the primary's mutex)
  datok_model* c = x ? x : m;
  int rc = -1;
  if (c->hm.eot_rewind && out && in && n < 0xFFFFFFFFull - (1u << 20) && n >= 2 * c->piece_bytes && c->pipelined)
    rc = run_pipelined(c, in, n, flags, carry_in, out);  // -1: no EOT to cut at; -2: DATOK_FORMAT with malformed UTF-8
  if (rc == -1 || rc == -2) rc = run_pipeline(c, in, false, n, flags, carry_in, false, out);
  if (x) ctx_release(m, x);
  return rc;
}

int datok_transduce_device(datok_model* m, const uint8_t* d_in, size_t n, uint32_t flags,
                           const datok_carry* carry_in, datok_result** out) {
  if (const int rc = check_norewind(m, flags)) return rc;
  if (!m) { g_last_error = "invalid argument"; return DATOK_ERR_INVALID_ARG; }
  datok_model* x = ctx_acquire(m);
  const int rc = run_pipeline(x ? x : m, d_in, true, n, flags, carry_in, true, out);
  if (x) ctx_release(m, x);
  return rc;
}

const datok_view* datok_result_view(const datok_result* r) { return r ? &r->view : nullptr; }

void datok_result_free(datok_result* r) {
  if (!r) return;
  datok_model* m = r->model;
  datok_model* root = root_of(m);
  {
    std::lock_guard<std::mutex> lock(m->mu);
    free_result_locked(r);
  }
  if (root->freed_by_user && root->live_results == 0) destroy_model(root);
}

int datok_last_kernel_times(const datok_model* m, const char** names, float* ms, int cap) {
  if (!m) return 0;
  int n = 0;
  for (int i = 0; i < T_COUNT && n < cap; i++, n++) {
    if (names) names[n] = kTimerNames[i];
    if (ms) ms[n] = m->t_ms[i];
  }
  return n;
}

int datok_last_launch_count(const datok_model* m) { return m ? m->launches : 0; }

int datok_last_stats(const datok_model* m, uint32_t* fixup_rounds, uint32_t* hot_rows, uint32_t* hot_cols, uint32_t* chunk_bytes) {
  if (!m) return DATOK_ERR_INVALID_ARG;
  if (fixup_rounds) *fixup_rounds = m->last_rounds;
  if (hot_rows) *hot_rows = m->n_hot;
  if (hot_cols) *hot_cols = m->hm.hot_cols;
  if (chunk_bytes) *chunk_bytes = m->chunk;
  return DATOK_OK;
}

int datok_measure_gather_bound(datok_model* m, double* byte_steps_per_s) {
  if (!m || !byte_steps_per_s) return DATOK_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard guard;
  CUDA_TRY(cudaSetDevice(m->device));
  const uint32_t row16 = m->dm.stride16 * 2u, segs = 256;
  int rc = ensure_workspace(m, 4096);
  if (rc) return rc;
  uint32_t* sink = reinterpret_cast<uint32_t*>(m->ws);
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  float best = 0;
  for (int it = 0; it < 4; it++) {  // the first launch warms up
    CUDA_TRY(cudaEventRecord(a, m->stream));
    const int e = launch_gather_bound(m->n_hot, row16, segs, m->n_sms, sink, m->stream);
    if (e != 0) { g_last_error = std::string("gather_bound launch: ") + cudaGetErrorString((cudaError_t)e); return DATOK_ERR_CUDA; }
    CUDA_TRY(cudaEventRecord(b, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    if (it > 0 && (best == 0 || ms < best)) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *byte_steps_per_s = (double)m->n_sms * 1024.0 * segs * 32.0 / (best * 1e-3);
  return DATOK_OK;
}

void* datok_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void datok_host_free(void* p) { if (p) cudaFreeHost(p); }

const char* datok_last_error(void) { return g_last_error.c_str(); }

const char* datok_strerror(int code) {
  switch (code) {
    case DATOK_OK: return "ok";
    case DATOK_ERR_BUFFER_OVERFLOW: return "more than 1024 runes without a token boundary (matrix.go:365,406)";
    case DATOK_ERR_SENT_NO_TOKEN: return "SentenceEnd before any token of the text under SENTENCE_POS (token_writer.go:108)";
    case DATOK_ERR_TEXT_NO_TOKEN: return "TextEnd on a token-less text under TOKEN_POS (token_writer.go:135)";
    case DATOK_ERR_TEXT_NO_SENT: return "TextEnd on a sentence-less text under SENTENCE_POS (token_writer.go:145)";
    case DATOK_ERR_DEGENERATE: return "degenerate event sequence (empty token slice or repeated SentenceEnd)";
    case DATOK_ERR_IO: return "cannot read model file";
    case DATOK_ERR_FORMAT: return "not a MATOK v1 model";
    case DATOK_ERR_UNSUPPORTED_MODEL: return "model not supported by the GPU layout";
    case DATOK_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    case DATOK_ERR_CUDA: return "CUDA error";
    case DATOK_ERR_TOO_LARGE: return "input too large for one call";
    case DATOK_ERR_INVALID_ARG: return "invalid argument";
    case DATOK_ERR_NOT_AT_BOUNDARY: return "non-final input does not end at a text boundary";
    case DATOK_ERR_COMPACT_RANGE: return "DATOK_COMPACT: a token delta does not fit 16 bits";
    default: return "unknown error";
  }
}

}  // extern "C"
