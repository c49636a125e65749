// multi.cpp -- the two front-ends of the C ABI that sit on top of datok_transduce():
//
//   datok_stream_*            the reference reads any io.Reader through a sliding window (matrix.go:373,388-408,
//                             cmd/datok.go:108-132 incl. STDIN); here the caller pushes blocks as they arrive, the
//                             stream cuts them after EOT bytes (a text boundary: matrix.go:593-605) and carries
//                             the walk state, sentenceEnd / textEnd and the writer's `init` flag from batch to batch.
//   datok_transduce_sharded   one corpus over the GPUs of a box: EOT-aligned byte-balanced shards,
//                             one per device, walked concurrently from the guessed carry; an NCCL all-gather of
//                             {bytes, tokens, sentences, texts, sent entries, carry-out state} gives every shard its
//                             global index bases and tells which guesses were wrong; those shards are walked again.
//
// Host code only (no kernels).  NCCL is bound at run time (dlopen): the library itself does not depend on it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/datok_b200.h"

// ------------------------------------------------------------------ streaming

struct datok_stream {
  datok_model* model;
  uint32_t flags;
  std::vector<uint8_t> pending;  // bytes pushed and not yet transduced (the tail behind the last EOT)
  datok_carry carry;
  bool have_carry = false;
  bool writer_used;
  bool finished = false;
  uint64_t bytes_done = 0;       // bytes of the stream before pending[0]
};

extern "C" {

datok_stream* datok_stream_open(datok_model* m, uint32_t flags) {
  if (!m) return nullptr;
  datok_stream* s = new datok_stream();
  s->model = m;
  s->flags = flags & ~(uint32_t)DATOK_NOT_FINAL;
  s->writer_used = (flags & DATOK_WRITER_USED) != 0;
  std::memset(&s->carry, 0, sizeof s->carry);
  return s;
}

static int stream_run(datok_stream* s, size_t n, bool final, datok_result** out) {
  uint32_t f = s->flags | (final ? 0u : (uint32_t)DATOK_NOT_FINAL) | (s->writer_used ? (uint32_t)DATOK_WRITER_USED : 0u);
  datok_result* r = nullptr;
  const int rc = datok_transduce(s->model, s->pending.data(), n, f, s->have_carry ? &s->carry : nullptr, &r);
  if (rc != DATOK_OK) return rc;
  const datok_view* v = datok_result_view(r);
  s->carry = v->carry_out;
  s->have_carry = true;
  if (v->n_tokens) s->writer_used = true;
  s->bytes_done += n;
  s->pending.erase(s->pending.begin(), s->pending.begin() + (ptrdiff_t)n);
  *out = r;
  return DATOK_OK;
}

int datok_stream_push(datok_stream* s, const uint8_t* data, size_t n, datok_result** out) {
  if (!s || !out || (!data && n) || s->finished) return DATOK_ERR_INVALID_ARG;
  *out = nullptr;
  s->pending.insert(s->pending.end(), data, data + n);
  // the batch ends right after the last EOT seen so far; without one the text is still open: keep reading
  size_t cut = s->pending.size();
  while (cut > 0 && s->pending[cut - 1] != 0x04) cut--;
  if (cut == 0) return DATOK_OK;
  return stream_run(s, cut, false, out);
}

int datok_stream_finish(datok_stream* s, datok_result** out) {
  if (!s || !out || s->finished) return DATOK_ERR_INVALID_ARG;
  *out = nullptr;
  s->finished = true;
  return stream_run(s, s->pending.size(), true, out);  // end-of-input processing (matrix.go:650-695)
}

uint64_t datok_stream_bytes_done(const datok_stream* s) { return s ? s->bytes_done : 0; }

void datok_stream_close(datok_stream* s) { delete s; }

}  // extern "C"

// ------------------------------------------------------------------ one corpus over several GPUs

namespace {

// the few NCCL entry points, bound at run time
struct Nccl {
  void* lib = nullptr;
  int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*AllGather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t s) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int NCCL_UINT64 = 5;  // ncclUint64 (nccl.h ncclDataType_t)

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.lib) break;
    }
    if (!n.lib) return;
    n.CommInitAll = (int (*)(void**, int, const int*))dlsym(n.lib, "ncclCommInitAll");
    n.CommDestroy = (int (*)(void*))dlsym(n.lib, "ncclCommDestroy");
    n.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(n.lib, "ncclAllGather");
    n.GroupStart = (int (*)())dlsym(n.lib, "ncclGroupStart");
    n.GroupEnd = (int (*)())dlsym(n.lib, "ncclGroupEnd");
    n.GetErrorString = (const char* (*)(int))dlsym(n.lib, "ncclGetErrorString");
    n.ok = n.CommInitAll && n.CommDestroy && n.AllGather && n.GroupStart && n.GroupEnd;
  });
  return n;
}

constexpr int N_COUNTS = 6;  // bytes, tokens, sentences, texts, sent entries, carry-out state

thread_local std::string g_multi_error;
thread_local int g_used_nccl = 0, g_rewalked = 0;

}  // namespace

extern "C" {

const char* datok_sharded_last_error(void) { return g_multi_error.c_str(); }
int datok_sharded_last_info(int* used_nccl, int* shards_rewalked) {
  if (used_nccl) *used_nccl = g_used_nccl;
  if (shards_rewalked) *shards_rewalked = g_rewalked;
  return DATOK_OK;
}

int datok_plan_shards(const uint8_t* in, size_t n, int n_shards, uint64_t* bounds) {
  if (n_shards < 1 || !bounds || (!in && n)) return DATOK_ERR_INVALID_ARG;
  bounds[0] = 0;
  for (int r = 1; r < n_shards; r++) {
    size_t p = (size_t)((unsigned __int128)n * (unsigned)r / (unsigned)n_shards);
    if (p < bounds[r - 1]) p = (size_t)bounds[r - 1];
    const void* q = p < n ? std::memchr(in + p, 0x04, n - p) : nullptr;  // the first EOT at or behind the balanced position
    bounds[r] = q ? (uint64_t)((const uint8_t*)q - in) + 1 : (uint64_t)n;
  }
  bounds[n_shards] = n;
  return DATOK_OK;
}

int datok_transduce_sharded(datok_model* const* models, const int* devices, int ndev, const uint8_t* in, size_t n,
                            uint32_t flags, const datok_carry* carry_in, datok_result** outs, uint64_t* bases,
                            uint64_t* bounds_out) {
  if (!models || !devices || ndev < 1 || !outs || (!in && n)) { g_multi_error = "invalid argument"; return DATOK_ERR_INVALID_ARG; }
  struct DeviceGuard {  // the caller's current device is left as it was
    int dev = -1;
    DeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) dev = -1; }
    ~DeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
  } guard;
  for (int i = 0; i < ndev; i++) outs[i] = nullptr;
  std::vector<uint64_t> bounds((size_t)ndev + 1);
  datok_plan_shards(in, n, ndev, bounds.data());
  const bool call_final = !(flags & DATOK_NOT_FINAL);
  const uint32_t base_flags = flags & ~(uint32_t)(DATOK_NOT_FINAL | DATOK_WRITER_USED);
  const bool used_in = (flags & DATOK_WRITER_USED) != 0;

  std::vector<int> rcs((size_t)ndev, DATOK_OK);
  std::vector<std::string> errs((size_t)ndev);
  std::vector<datok_carry> carry((size_t)ndev);   // the carry each shard is walked from
  std::vector<char> used((size_t)ndev, 0);        // DATOK_WRITER_USED of each shard
  // guesses: every shard but the first starts in the root state right behind a finished text, and some shard before
  // it has produced a token
  for (int i = 0; i < ndev; i++) {
    std::memset(&carry[i], 0, sizeof(datok_carry));
    if (i == 0) { if (carry_in) carry[0] = *carry_in; used[0] = used_in; }
    else { carry[i].state = 1; carry[i].sentence_end = 1; carry[i].text_end = 1; used[i] = 1; }
  }
  auto walk = [&](int i) {
    const bool last = i == ndev - 1;
    const uint32_t f = base_flags | ((last && call_final) ? 0u : (uint32_t)DATOK_NOT_FINAL) | (used[i] ? (uint32_t)DATOK_WRITER_USED : 0u);
    if (outs[i]) { datok_result_free(outs[i]); outs[i] = nullptr; }
    rcs[i] = datok_transduce(models[i], in + bounds[i], (size_t)(bounds[i + 1] - bounds[i]), f, (i == 0 && !carry_in) ? nullptr : &carry[i], &outs[i]);
    if (rcs[i] != DATOK_OK) errs[i] = datok_last_error();
  };
  auto walk_all = [&](const std::vector<int>& which) {
    std::vector<std::thread> th;
    for (int i : which) th.emplace_back(walk, i);
    for (auto& t : th) t.join();
    for (int i : which)
      if (rcs[i] != DATOK_OK) { g_multi_error = "shard " + std::to_string(i) + ": " + errs[i]; return rcs[i]; }
    return (int)DATOK_OK;
  };
  auto fail = [&](int rc) {
    for (int i = 0; i < ndev; i++) if (outs[i]) { datok_result_free(outs[i]); outs[i] = nullptr; }
    return rc;
  };
  g_used_nccl = 0; g_rewalked = 0;
  std::vector<int> todo;
  for (int i = 0; i < ndev; i++) todo.push_back(i);
  int rc = walk_all(todo);
  if (rc) return fail(rc);

  // ---- the exchange: per-shard counts and carry-out, all-gathered over NCCL (NVLink / NVSwitch) ----
  std::vector<uint64_t> all((size_t)ndev * N_COUNTS);
  Nccl& nc = nccl();
  std::vector<void*> comms((size_t)ndev, nullptr);
  std::vector<uint64_t*> d_send((size_t)ndev, nullptr), d_recv((size_t)ndev, nullptr);
  std::vector<cudaStream_t> streams((size_t)ndev, nullptr);
  bool use_nccl = ndev > 1 && nc.ok;
  if (use_nccl) {
    int e = nc.CommInitAll(comms.data(), ndev, devices);
    if (e != 0) { g_multi_error = std::string("ncclCommInitAll: ") + (nc.GetErrorString ? nc.GetErrorString(e) : "?"); use_nccl = false; }
  }
  if (use_nccl) {
    for (int i = 0; i < ndev; i++) {
      cudaSetDevice(devices[i]);
      cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking);
      cudaMalloc((void**)&d_send[i], N_COUNTS * sizeof(uint64_t));
      cudaMalloc((void**)&d_recv[i], (size_t)ndev * N_COUNTS * sizeof(uint64_t));
    }
  }
  auto gather = [&]() -> int {
    std::vector<uint64_t> mine((size_t)ndev * N_COUNTS);
    for (int i = 0; i < ndev; i++) {
      const datok_view* v = datok_result_view(outs[i]);
      uint64_t* c = &mine[(size_t)i * N_COUNTS];
      c[0] = bounds[i + 1] - bounds[i]; c[1] = v->n_tokens; c[2] = v->n_sentences; c[3] = v->n_texts; c[4] = v->n_sent_pos;
      c[5] = v->carry_out.state;
    }
    if (!use_nccl) { all = mine; return 0; }  // one device, or no NCCL in this process: nothing to exchange
    for (int i = 0; i < ndev; i++) {
      cudaSetDevice(devices[i]);
      cudaMemcpyAsync(d_send[i], &mine[(size_t)i * N_COUNTS], N_COUNTS * sizeof(uint64_t), cudaMemcpyHostToDevice, streams[i]);
    }
    nc.GroupStart();
    for (int i = 0; i < ndev; i++) {
      const int e = nc.AllGather(d_send[i], d_recv[i], N_COUNTS, NCCL_UINT64, comms[i], streams[i]);
      if (e != 0) { nc.GroupEnd(); g_multi_error = std::string("ncclAllGather: ") + (nc.GetErrorString ? nc.GetErrorString(e) : "?"); return e; }
    }
    const int e = nc.GroupEnd();
    if (e != 0) { g_multi_error = std::string("ncclGroupEnd: ") + (nc.GetErrorString ? nc.GetErrorString(e) : "?"); return e; }
    g_used_nccl = 1;
    // every device holds the whole table now; the host reads device 0's copy
    cudaSetDevice(devices[0]);
    cudaMemcpyAsync(all.data(), d_recv[0], all.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, streams[0]);
    for (int i = 0; i < ndev; i++) { cudaSetDevice(devices[i]); cudaStreamSynchronize(streams[i]); }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  };
  auto cleanup = [&] {
    if (!use_nccl) return;
    for (int i = 0; i < ndev; i++) {
      cudaSetDevice(devices[i]);
      if (d_send[i]) cudaFree(d_send[i]);
      if (d_recv[i]) cudaFree(d_recv[i]);
      if (streams[i]) cudaStreamDestroy(streams[i]);
      if (comms[i]) nc.CommDestroy(comms[i]);
    }
  };

  // ---- verify the guesses; walk the shards whose guess was wrong again (matrix.go:593-605: the state behind an
  // EOT is whatever the matrix says; token_writer.go:42,70: `init`), until nothing changes ----
  for (int round = 0; round <= ndev; round++) {
    if (gather() != 0) { cleanup(); return fail(DATOK_ERR_CUDA); }
    todo.clear();
    uint64_t tokens_before = 0;
    uint32_t state = carry_in && carry_in->state ? carry_in->state : 1;  // state the stream is in at the shard's start
    bool any_before = used_in;
    for (int i = 0; i < ndev; i++) {
      const uint64_t* c = &all[(size_t)i * N_COUNTS];
      if (i > 0) {
        const bool want_used = any_before || tokens_before > 0;
        if (c[0] != 0 && (carry[i].state != state || (bool)used[i] != want_used)) {
          carry[i].state = state; used[i] = want_used;
          todo.push_back(i);
        }
      }
      if (c[0] != 0) state = (uint32_t)c[5];  // (an empty shard passes the carry through)
      tokens_before += c[1];
    }
    if (todo.empty()) break;
    // a corrected shard may end in another state than before: only the first one is certain, the rest is checked again
    todo.resize(1);
    g_rewalked++;
    if ((rc = walk_all(todo))) { cleanup(); return fail(rc); }
  }
  cleanup();
  if (bases) {
    uint64_t acc[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < ndev; i++) {
      for (int k = 0; k < 5; k++) bases[(size_t)i * 5 + k] = acc[k];
      for (int k = 0; k < 5; k++) acc[k] += all[(size_t)i * N_COUNTS + k];
    }
  }
  if (bounds_out) std::memcpy(bounds_out, bounds.data(), bounds.size() * sizeof(uint64_t));
  return DATOK_OK;
}

}  // extern "C"
