// kernels.cu -- sm_100a kernels of the Datok matrix-FSA transduction path.
//
//   K1+K2a walk_fused_kernel UTF-8 decode + sigma map + speculative per-chunk walk, fused
//                            (matrix.go:388-435 and 437-695); persistent, hot table rows in smem
//   K2b stitch_kernel        head of each chunk from the predecessor's exit state
//   K2c rewalk_kernel        chunks whose guessed start state was wrong
//   K2d commit_kernel        publish changed exit states, queue successors
//   K3a compact_reduce       per-block summaries of the boundary bitmaps  (token_writer.go:36-175)
//   K3b compact_scan         exclusive scan of the block summaries
//   K3c compact_emit         offset arrays
//   K3d compact_finalize     end-of-stream SentenceEnd / TextEnd          (matrix.go:680-695)
//
// The per-thread bodies live in walk_core.cuh / chunk_core.cuh / compact_core.cuh.
#include "kernels.cuh"

#include <cstdlib>

namespace datok {

// ------------------------------------------------------------------ K1 + K2a (fused)

constexpr int WALK_THREADS = 128;
constexpr int LANE_CLS_STRIDE = 36;  // bytes of class scratch per lane (9 words: bank spread)
// Staging slot per lane for the next segment's raw bytes (cp.async, LDGSTS): a build option, off by default.
// Measured (profiles/r2_stage_ab.txt, 1 GiB German): walk 3.88 ms without, 4.51 ms with -- the 32 KB of slots
// cost ~200 resident table rows and the slot bookkeeping two registers in a kernel that is at its register limit,
// which outweighs the input latency it hides (the L1 prefetch of the next sector already hides most of it).
#if defined(DATOK_STAGE_TMA)
#if !defined(DATOK_STAGE_ASYNC)
#error "DATOK_STAGE_TMA is a flavour of DATOK_STAGE_ASYNC: define both"
#endif
constexpr int LANE_STAGE_BYTES = 48;  // slot + the lane's mbarrier (16-byte aligned slots)
#elif defined(DATOK_STAGE_ASYNC)
constexpr int LANE_STAGE_BYTES = 32;
#else
constexpr int LANE_STAGE_BYTES = 0;
#endif
constexpr int LANE_SMEM = LANE_CLS_STRIDE + LANE_STAGE_BYTES;
constexpr int MAX_RUNES_SMEM = 96;   // the rune table of sigma is kept in shared memory up to this size
constexpr int WALK_LUT_BYTES = 544 + 5 * MAX_RUNES_SMEM;  // LUTs of the walk kernel (multiple of 16)
static_assert(WALK_LUT_BYTES % 16 == 0, "the compact rows start 16-byte aligned");

// Persistent kernel, one CTA per SM.  The compact (u16) rows of the hottest states and the
// byte->class LUTs live in shared memory; every lane owns one chunk at a time and walks it
// segment by segment (chunk_spec_fast).
// REWALK: the same tables, but the lanes re-walk the chunks of list_rewalk (K2c) from their true states.
template <int THREADS, bool REWALK>
__global__ void __launch_bounds__(THREADS, 1)
walk_fused_kernel(DeviceModel m, WalkBuffers b, uint32_t start_state, uint32_t n_hot) {
  extern __shared__ __align__(16) uint32_t smem[];
  // (the list length is only known on the device.  The list is spread over all warps of the grid, `per` chunks to
  // a warp (lanes 0..per-1).  The usual list -- a fraction of a percent of the chunks -- gives one chunk per warp:
  // the kernel's time then is the latency of one chunk's walk, and a lane that walks alone neither waits for the
  // slowest lane of its warp at every segment nor sits through the other lanes' rare paths.  Lists too long for
  // that are walked one chunk per lane like the first walk.)
  uint32_t per = 0;
  if (REWALK) {
    const uint32_t n = b.counters[1], W = gridDim.x * (THREADS / 32);
    per = (n + W - 1) / W;
    if (per > 32) per = 0;  // dense
    if (blockIdx.x * (per ? per * (THREADS / 32) : THREADS) >= n) return;
  }
  // layout: lane scratch | byte -> class LUTs | compact rows.  The first two have compile-time offsets, so a
  // lane's scratch address is threadIdx.x * stride away from the window base wherever it is needed again.
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(smem);           // THREADS staging slots (16-byte aligned)
  uint8_t* s_cls = s_stage + THREADS * LANE_STAGE_BYTES;
  uint8_t* s_lut = s_cls + THREADS * LANE_CLS_STRIDE;
  uint16_t* s_hot = reinterpret_cast<uint16_t*>(s_lut + WALK_LUT_BYTES);
  const uint32_t hot_entries = n_hot * m.stride16;
  for (uint32_t k = threadIdx.x; k < hot_entries; k += THREADS) {
    uint32_t e = m.hot16[k];
    if ((e & F16_TGT) >= n_hot) e = 0;  // the target's row is not resident: that step goes through T3
    s_hot[k] = (uint16_t)e;
  }
  for (uint32_t k = threadIdx.x; k < m.stride16; k += THREADS) s_hot[hot_entries + k] = 0;  // row n_hot: "see T3"
  // s_lut: ascii_cls[128] latin1_cls[128] | doubled, capped class per byte [256] | sync classes [32 B] |
  // rune_key[MAX_RUNES_SMEM] u32, rune_cls[MAX_RUNES_SMEM] (the sorted non-Latin-1 runes of sigma)
  uint32_t* s_sync = reinterpret_cast<uint32_t*>(s_lut + 512);
  uint32_t* s_rkey = reinterpret_cast<uint32_t*>(s_lut + 544);
  uint8_t* s_rcls = s_lut + 544 + 4 * MAX_RUNES_SMEM;
  if (threadIdx.x < 128) {
    s_lut[threadIdx.x] = m.cls.ascii_cls[threadIdx.x];
    s_lut[128 + threadIdx.x] = m.cls.latin1_cls[threadIdx.x];
    s_lut[256 + threadIdx.x] = (uint8_t)cap_cl2(m.cls.ascii_cls[threadIdx.x], 2u * m.hot_cols);
    s_lut[384 + threadIdx.x] = (uint8_t)(128 + threadIdx.x);
    if (threadIdx.x < 8) s_sync[threadIdx.x] = m.sync_cls[threadIdx.x];
    if (threadIdx.x < MAX_RUNES_SMEM && threadIdx.x < m.cls.n_rune) {
      s_rkey[threadIdx.x] = m.cls.rune_key[threadIdx.x];
      s_rcls[threadIdx.x] = m.cls.rune_cls[threadIdx.x];
    }
  }
  __syncthreads();
  DeviceModel lm = m;
  lm.cls.ascii_cls = s_lut;
  lm.cls.latin1_cls = s_lut + 128;
  if (m.cls.n_rune <= MAX_RUNES_SMEM) { lm.cls.rune_key = s_rkey; lm.cls.rune_cls = s_rcls; }
  {  // the struct itself, for the out-of-line classify_pos
    __shared__ ClsTables s_ct;
    lm.cls.self = &s_ct;
    if (threadIdx.x == 0) s_ct = lm.cls;
    __syncthreads();
  }
  FastTables FT;
  FT.hot16 = s_hot; FT.t3 = m.table2; FT.n_hot = n_hot; FT.row16 = m.stride16 * 2u; FT.stride3 = m.stride2;
  FT.ascii_cls2 = s_lut + 256;
  FT.sync_cls = s_sync;
  FT.stop_cl2 = 2u * m.hot_cols;
  FT.eot_rewind = m.eot_rewind;
  FT.row16_inv = 0xFFFFFFFFu / FT.row16 + 1u;  // (row16 is not a power of two: floor(2^32 / row16) == floor((2^32 - 1) / row16))
  {  // opaque to the compiler: otherwise the shared-window base is re-derived in every step of the hot loop
    const unsigned long long sa = __cvta_generic_to_shared(s_hot);
    asm volatile("cvt.u32.u64 %0, %1;" : "=r"(FT.hot_saddr) : "l"(sa));
  }
  uint8_t* my_cls = s_cls + threadIdx.x * LANE_CLS_STRIDE;
  const uint32_t my_stage = LANE_STAGE_BYTES ? (uint32_t)__cvta_generic_to_shared(s_stage + threadIdx.x * LANE_STAGE_BYTES) : 0u;
  if (REWALK) {
    const uint32_t n = b.counters[1];
    if (per) {
      const uint32_t lane = threadIdx.x & 31u, k = (blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * per + lane;
      if (lane < per && k < n) chunk_rewalk_fast(lm, b, FT, b.list_rewalk[k], my_cls, my_stage);
    } else {
      for (uint32_t k = blockIdx.x * THREADS + threadIdx.x; k < n; k += gridDim.x * THREADS)
        chunk_rewalk_fast(lm, b, FT, b.list_rewalk[k], my_cls, my_stage);
    }
  } else {
    for (;;) {
      uint32_t base = 0;
      if ((threadIdx.x & 31u) == 0) base = atomicAdd(&b.counters[6], 32u);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (base >= b.n_chunks) break;
      const uint32_t i = base + (threadIdx.x & 31u);
      if (i < b.n_chunks) chunk_spec_fast(lm, b, FT, i, start_state, my_cls, nullptr, my_stage);
      __syncwarp();
    }
  }
}

int fused_threads_from_env() {
  int t = 1024;
  if (const char* s = std::getenv("DATOK_FUSED_THREADS")) {
    const long v = std::atol(s);
    if (v == 256 || v == 512 || v == 768 || v == 1024) t = (int)v;
  }
  return t;
}

static size_t fused_smem_bytes_t(const DeviceModel& m, uint32_t n_hot, int threads) {
  return ((((size_t)n_hot + 1) * m.stride16 * 2 + 15) & ~(size_t)15) + (size_t)threads * LANE_SMEM + WALK_LUT_BYTES;
}
static_assert(LANE_CLS_STRIDE % 4 == 0, "the LUTs and the compact rows behind the lane scratch stay 4-byte aligned");
size_t fused_smem_bytes(const DeviceModel& m, uint32_t n_hot, int threads) { return fused_smem_bytes_t(m, n_hot, threads); }

uint32_t fused_max_hot_rows(const DeviceModel& m, size_t smem_limit, uint32_t n_states, int threads) {
  const size_t fixed = (size_t)threads * LANE_SMEM + WALK_LUT_BYTES + 16 + 1024 + (size_t)m.stride16 * 2;
  if (smem_limit <= fixed) return 1;
  size_t rows = (smem_limit - fixed) / ((size_t)m.stride16 * 2);
  if (rows > (size_t)n_states + 1) rows = (size_t)n_states + 1;
  if (rows > m.hot16_rows) rows = m.hot16_rows;
  return rows ? (uint32_t)rows : 1u;
}

template <int THREADS>
static int launch_walk_fused_t(const DeviceModel& m, const WalkBuffers& b, uint32_t start_state, uint32_t n_hot,
                               int n_sms, cudaStream_t s) {
  const size_t smem = fused_smem_bytes_t(m, n_hot, THREADS);
  {  // per launch: the attribute belongs to the current device, and a process may drive several
    cudaError_t e = cudaFuncSetAttribute(walk_fused_kernel<THREADS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  uint32_t blocks = (b.n_chunks + THREADS - 1) / THREADS;
  if (blocks > (uint32_t)n_sms) blocks = (uint32_t)n_sms;
  walk_fused_kernel<THREADS, false><<<blocks, THREADS, smem, s>>>(m, b, start_state, n_hot);
  return (int)cudaGetLastError();
}

// K2c: the mismatched chunks, with the hot rows in shared memory like the first walk
constexpr int REWALK_THREADS = 512;
int launch_rewalk_fused(const DeviceModel& m, const WalkBuffers& b, uint32_t n_rewalk_max, uint32_t n_hot, int n_sms,
                        cudaStream_t s) {
  if (!n_rewalk_max) return 0;
  const size_t smem = fused_smem_bytes_t(m, n_hot, REWALK_THREADS);
  {
    cudaError_t e = cudaFuncSetAttribute(walk_fused_kernel<REWALK_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  // (one CTA per SM whenever the list could need them: the device picks one chunk per warp if the list is short)
  uint32_t blocks = (n_rewalk_max + REWALK_THREADS / 32 - 1) / (REWALK_THREADS / 32);
  if (blocks > (uint32_t)n_sms) blocks = (uint32_t)n_sms;
  walk_fused_kernel<REWALK_THREADS, true><<<blocks, REWALK_THREADS, smem, s>>>(m, b, 0, n_hot);
  return (int)cudaGetLastError();
}

int launch_walk_fused(const DeviceModel& m, const WalkBuffers& b, uint32_t start_state, uint32_t n_hot,
                      int n_sms, cudaStream_t s, int threads) {
  switch (threads) {
    case 256: return launch_walk_fused_t<256>(m, b, start_state, n_hot, n_sms, s);
    case 768: return launch_walk_fused_t<768>(m, b, start_state, n_hot, n_sms, s);
    case 512: return launch_walk_fused_t<512>(m, b, start_state, n_hot, n_sms, s);
    default: return launch_walk_fused_t<1024>(m, b, start_state, n_hot, n_sms, s);
  }
}

// Calibration: visits per state on a sample, walked speculatively chunk by chunk
// with the exact walker.  Wrong guesses only add noise to the ranking.
__global__ void __launch_bounds__(WALK_THREADS) hist_kernel(DeviceModel m, WalkBuffers b, uint32_t* hist, uint32_t hist_cls_offset) {
  __shared__ ClsTables s_ct;
  if (threadIdx.x == 0) { s_ct = m.cls; s_ct.self = &s_ct; }
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n_chunks) return;
  WalkCtx c = make_walk_ctx(m, b);
  c.cls.self = &s_ct;
  c.hist = hist;
  c.hist_cls = hist + hist_cls_offset;
  const uint32_t lo = i * b.chunk, hi = lo + b.chunk;
  const uint32_t s = i == 0 ? 0u : find_sync(b.in, b.N, m.sync_ascii, lo, hi);
  if (s == K_NOPOS) return;
  WState st;
  st.pos = st.tstart = st.base = st.hw = s;
  st.eps_pos = 0; st.eps_state = 0; st.flags = 0; st.t = (uint16_t)m.start;
  SpecInfo si;
  walk_run<true, true, false>(c, st, hi, &si);
}

void launch_hist(const DeviceModel& m, const WalkBuffers& b, uint32_t* hist, uint32_t hist_cls_offset, cudaStream_t s) {
  hist_kernel<<<(b.n_chunks + WALK_THREADS - 1) / WALK_THREADS, WALK_THREADS, 0, s>>>(m, b, hist, hist_cls_offset);
}

// ------------------------------------------------------------------ gather bound (measurement only)

// What the hardware allows a lane-per-chunk table walk of this shape: the walk's configuration (one
// 1024-thread CTA per SM, the compact rows in shared memory) with nothing but the dependent chain of one
// byte step -- class byte from the lane's scratch, row entry of (state, class), next state -- and one
// XOR to keep the result.  bench.py reports the fused walk against this figure next to the HBM roofline
// (R_gather).  The rows hold pseudo-random targets below n_rows.
__global__ void __launch_bounds__(1024, 1) gather_bound_kernel(uint32_t n_rows, uint32_t row16, uint32_t segs, uint32_t* sink) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint16_t* s_hot = reinterpret_cast<uint16_t*>(smem);
  const uint32_t entries = n_rows * (row16 / 2);
  uint8_t* s_cls = reinterpret_cast<uint8_t*>(smem) + ((entries * 2u + 15u) & ~15u);
  uint32_t x = 0x9E3779B9u * (blockIdx.x + 1);
  for (uint32_t k = threadIdx.x; k < entries; k += 1024) {
    uint32_t h = (k + x) * 2654435761u;
    h ^= h >> 15;
    s_hot[k] = (uint16_t)(h % n_rows);
  }
  uint8_t* my = s_cls + threadIdx.x * LANE_CLS_STRIDE;
  for (uint32_t k = 0; k < 32; k++) {
    uint32_t h = (threadIdx.x * 32 + k + x) * 2246822519u;
    h ^= h >> 13;
    my[k] = (uint8_t)(2u * (h % (row16 / 2)));
  }
  __syncthreads();
  const uint32_t hot_saddr = (uint32_t)__cvta_generic_to_shared(s_hot), cls_saddr = (uint32_t)__cvta_generic_to_shared(my);
  uint32_t t = threadIdx.x % n_rows, acc = 0;
  for (uint32_t sg = 0; sg < segs; sg++) {
#pragma unroll 1
    for (uint32_t off = 0; off < 32; off++) {
      uint32_t cl2, e;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(cl2) : "r"(cls_saddr + off));
      asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(hot_saddr + t * row16 + cl2));
      acc ^= e + off;
      t = e;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;  // (keeps the chain alive)
}

int launch_gather_bound(uint32_t n_rows, uint32_t row16, uint32_t segs, int n_sms, uint32_t* sink, cudaStream_t s) {
  const size_t smem = (((size_t)n_rows * row16 + 15) & ~(size_t)15) + 1024 * LANE_CLS_STRIDE;
  cudaError_t e = cudaFuncSetAttribute(gather_bound_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  gather_bound_kernel<<<n_sms, 1024, smem, s>>>(n_rows, row16, segs, sink);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ K2b-d

#ifndef DATOK_STITCH_MINB
#define DATOK_STITCH_MINB 6
#endif
__global__ void __launch_bounds__(WALK_THREADS, DATOK_STITCH_MINB) stitch_kernel(DeviceModel m, WalkBuffers b, const uint32_t* list,
                                                              uint32_t n_list) {
  __shared__ ClsTables s_ct;
  if (threadIdx.x == 0) { s_ct = m.cls; s_ct.self = &s_ct; }
  __syncthreads();
  m.cls.self = &s_ct;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_list) return;
  const uint32_t i = list ? list[k] : k + 1;
  if (chunk_stitch(m, b, i, list == nullptr)) {  // list == nullptr: the first round over all chunks
    const uint32_t slot = atomicAdd(&b.counters[1], 1u);
    b.list_rewalk[slot] = i;
  }
}

__global__ void __launch_bounds__(256) commit_kernel(WalkBuffers b, const uint32_t* list, uint32_t n_list) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_list) return;
  const uint32_t i = list ? list[k] : k + 1;
  if (chunk_commit(b, i) && i + 1 < b.n_chunks) {
    const uint32_t slot = atomicAdd(&b.counters[0], 1u);
    b.list_next[slot] = i + 1;
  }
}

__global__ void __launch_bounds__(256) collect_errors_kernel(WalkBuffers b) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n_chunks) return;
  const uint32_t f = b.E[i].flags;
  if (f & WS_INVALID) {
    const uint32_t code = f >> WS_ERR_SHIFT;
    atomicMin(b.err_key, ((unsigned long long)(i * b.chunk) << 8) | (code ? code : 0xFFu));
  }
  if (i == b.n_chunks - 1 && !(f & (WS_DONE | WS_INVALID)))
    atomicMin(b.err_key, ((unsigned long long)(i * b.chunk) << 8) | 0xFEu);
}

// A single chain of dependent chunks (text without sync points: every chunk can only be walked once its
// predecessor's exit state is known) followed on the device: stitch, re-walk, commit, next chunk -- up to
// max_steps chunks per launch instead of one host-synchronised round per chunk.
__global__ void chain_kernel(DeviceModel m, WalkBuffers b, const uint32_t* list, uint32_t max_steps) {
  __shared__ ClsTables s_ct;
  s_ct = m.cls; s_ct.self = &s_ct;
  m.cls.self = &s_ct;
  uint32_t i = list[0], n_next = 0;
  for (uint32_t steps = 0;; steps++) {
    if (chunk_stitch(m, b, i, false)) chunk_rewalk(m, b, i);
    const bool changed = chunk_commit(b, i);
    if (!changed || i + 1 >= b.n_chunks) break;
    i++;
    if (steps + 1 == max_steps) { b.list_next[0] = i; n_next = 1; break; }
  }
  b.counters[0] = n_next;
  b.counters[1] = 0;
}
void launch_chain(const DeviceModel& m, const WalkBuffers& b, const uint32_t* list, uint32_t max_steps, cudaStream_t s) {
  chain_kernel<<<1, 1, 0, s>>>(m, b, list, max_steps);
}

void launch_stitch(const DeviceModel& m, const WalkBuffers& b, const uint32_t* list, uint32_t n_list, cudaStream_t s) {
  if (!n_list) return;
  stitch_kernel<<<(n_list + WALK_THREADS - 1) / WALK_THREADS, WALK_THREADS, 0, s>>>(m, b, list, n_list);
}
void launch_commit(const WalkBuffers& b, const uint32_t* list, uint32_t n_list, cudaStream_t s) {
  if (!n_list) return;
  commit_kernel<<<(n_list + 255) / 256, 256, 0, s>>>(b, list, n_list);
}
void launch_collect_errors(const WalkBuffers& b, cudaStream_t s) {
  collect_errors_kernel<<<(b.n_chunks + 255) / 256, 256, 0, s>>>(b);
}

// ------------------------------------------------------------------ host mailbox

// Small results the host waits for (round counters, stream summary, error key, last walk state) are
// written by this kernel straight into mapped pinned host memory.  A cudaMemcpy would queue on the
// device-to-host copy engine behind the result arrays of the previous piece and stall the pipeline.
__global__ void mail_kernel(MailSrc src, uint32_t* dst) {
  for (int k = 0; k < 4; k++)
    for (uint32_t i = threadIdx.x; i < src.words[k]; i += blockDim.x) dst[src.off[k] + i] = src.p[k][i];
  __threadfence_system();
}
void launch_mail(const MailSrc& src, uint32_t* dst_mapped, cudaStream_t s) { mail_kernel<<<1, 32, 0, s>>>(src, dst_mapped); }

// ------------------------------------------------------------------ K3

union AggWords {
  Agg a;
  uint32_t w[AGG_WORDS];
};
static_assert(sizeof(Agg) == AGG_WORDS * 4, "Agg layout");
static_assert(sizeof(StreamTotals) == sizeof(Agg), "totals share the Agg slots");

__device__ __forceinline__ Agg agg_shfl_up(const Agg& v, int delta) {
  AggWords in, out;
  in.a = v;
#pragma unroll
  for (int k = 0; k < AGG_WORDS; k++) out.w[k] = __shfl_up_sync(0xFFFFFFFFu, in.w[k], delta);
  return out.a;
}
__device__ __forceinline__ Agg agg_shfl_down(const Agg& v, int delta) {
  AggWords in, out;
  in.a = v;
#pragma unroll
  for (int k = 0; k < AGG_WORDS; k++) out.w[k] = __shfl_down_sync(0xFFFFFFFFu, in.w[k], delta);
  return out.a;
}

// Ordered (non-commutative) block scan.  Returns the exclusive prefix of `mine`
// within the block combined after `seed`.
template <int THREADS>
__device__ __forceinline__ Agg block_exclusive_scan(const Agg& mine, const Agg& seed) {
  __shared__ Agg s_warp[THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Agg incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Agg o = agg_shfl_up(incl, d);
    if (lane >= d) incl = agg_combine(o, incl);
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  Agg prefix = seed;
  for (int wi = 0; wi < warp; wi++) prefix = agg_combine(prefix, s_warp[wi]);
  Agg prev = agg_shfl_up(incl, 1);
  if (lane > 0) prefix = agg_combine(prefix, prev);
  __syncthreads();
  return prefix;
}

// Ordered block reduction of the reduce pass.  The texts and emit passes work in UNITS of 32 * COMPACT_WPT words (one
// warp of theirs); the reduce pass may give a thread more words (REDUCE_WPT, a multiple of COMPACT_WPT: fewer, longer
// threads measured faster for this pass), a unit then is LANES = 32 * COMPACT_WPT / REDUCE_WPT of its lanes.
// unit_out: for each of the block's units the summary of the units BEFORE it in the block (what the other passes seed
// their warp scans with: one combine instead of a loop), with WAGG_HAS_TEXT set in `kinds` when the unit's own words
// hold a TextEnd (the texts pass only visits those).  Returns the block's summary, valid in thread UNITS - 1.
template <int THREADS, int LANES>
__device__ __forceinline__ Agg block_reduce_units(const Agg& mine, Agg* unit_out) {
  constexpr int UNITS = THREADS / LANES;
  static_assert(LANES <= 32 && (LANES & (LANES - 1)) == 0 && UNITS <= 32, "unit shape");
  __shared__ Agg s_unit[UNITS];
  const int lane = threadIdx.x & 31;
  Agg v = mine;
#pragma unroll
  for (int d = 1; d < LANES; d <<= 1) {
    Agg o = agg_shfl_down(v, d);
    if ((lane & (2 * d - 1)) == 0) v = agg_combine(v, o);  // lanes lane..lane+2d-1, in order
  }
  if ((lane & (LANES - 1)) == 0) s_unit[threadIdx.x / LANES] = v;
  __syncthreads();
  Agg tot = agg_zero();
  if (threadIdx.x < 32) {  // ordered scan of the unit totals by the first warp: lane j ends up with the units before unit j
    const Agg own = lane < UNITS ? s_unit[lane] : agg_zero();
    Agg incl = own;
#pragma unroll
    for (int d = 1; d < UNITS; d <<= 1) {
      Agg o = agg_shfl_up(incl, d);
      if (lane >= d) incl = agg_combine(o, incl);
    }
    Agg ex = agg_shfl_up(incl, 1);
    if (lane == 0) ex = agg_zero();
    tot = incl;  // (lane UNITS - 1: the block)
    if (lane < UNITS) {
      if (own.n_text) ex.kinds |= WAGG_HAS_TEXT;
      unit_out[lane] = ex;
    }
  }
  return tot;
}

// Ordered exclusive scan within one warp, combined after `seed`.
__device__ __forceinline__ Agg warp_exclusive_scan(const Agg& mine, const Agg& seed) {
  const int lane = threadIdx.x & 31;
  Agg incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Agg o = agg_shfl_up(incl, d);
    if (lane >= d) incl = agg_combine(o, incl);
  }
  Agg prev = agg_shfl_up(incl, 1);
  return lane > 0 ? agg_combine(seed, prev) : seed;
}

enum { K3_REDUCE = 0, K3_TEXTS = 1, K3_EMIT = 2 };

// K3a: a CTA takes one block of THREADS * WPT words after the other (grid-stride) and leaves its summary in
// block_agg[block] (and the warp prefixes in warp_agg) -- the same units the texts and emit passes use.
constexpr int REDUCE_WPT = DATOK_REDUCE_WPT;
constexpr int REDUCE_THREADS = COMPACT_THREADS * COMPACT_WPT / REDUCE_WPT;  // the same block of words as the other passes
static_assert(REDUCE_WPT % COMPACT_WPT == 0 && REDUCE_THREADS % 32 == 0 && REDUCE_THREADS * REDUCE_WPT == COMPACT_THREADS * COMPACT_WPT,
              "the reduce pass covers the blocks of the texts / emit passes");
__global__ void __launch_bounds__(REDUCE_THREADS) compact_reduce_kernel(CompactCtx c, CompactBuffers cb) {
  constexpr int UNITS = COMPACT_THREADS / 32;
  for (uint32_t vb = blockIdx.x; vb < cb.n_blocks; vb += gridDim.x) {
    const uint32_t w0 = (vb * REDUCE_THREADS + threadIdx.x) * REDUCE_WPT;
    Agg ta = agg_zero();
#pragma unroll
    for (int k = 0; k < REDUCE_WPT; k++) {
      const uint32_t w = w0 + k;
      if (w < c.n_words) ta = agg_combine(ta, word_agg(w, word_load(c, w)));
    }
    const Agg tot = block_reduce_units<REDUCE_THREADS, 32 * COMPACT_WPT / REDUCE_WPT>(ta, cb.warp_agg + (size_t)vb * UNITS);
    if (threadIdx.x == UNITS - 1) cb.block_agg[vb] = tot;
    __syncthreads();  // the shared slots are reused by the next block
  }
}


// One pass over the bitmaps, COMPACT_WPT words per thread.
//   (block summaries: compact_reduce_kernel)
//   K3_TEXTS   TextEnd events: per-text bounds and the DocRec table
//   K3_EMIT    Token and SentenceEnd events; a block's tokens are staged in shared memory and
//              written out as contiguous 8-byte pairs
template <int MODE>
__global__ void __launch_bounds__(COMPACT_THREADS) compact_kernel(CompactCtx c, CompactBuffers cb) {
  extern __shared__ __align__(16) uint32_t s_stage[];
  if (MODE == K3_TEXTS) {
    // TextEnds are rare (one per text).  The reduce pass has marked the warp units (32 * COMPACT_WPT words) that
    // hold one: a warp reads 32 marks at a time and visits the marked units only, each with the whole warp.
    if (blockIdx.x == 0 && threadIdx.x == 0) c.docs[0] = doc_stream_start(c);
    constexpr uint32_t UNIT = 32 * COMPACT_WPT, WARPS = COMPACT_THREADS / 32;
    const uint32_t n_units = cb.n_blocks * WARPS, lane = threadIdx.x & 31;
    const uint32_t n_groups = (n_units + 31) / 32;
    for (uint32_t g = blockIdx.x * WARPS + (threadIdx.x >> 5); g < n_groups; g += gridDim.x * WARPS) {
      const uint32_t mu = g * 32 + lane;
      uint32_t marks = __ballot_sync(0xFFFFFFFFu, mu < n_units && (cb.warp_agg[mu].kinds & WAGG_HAS_TEXT) != 0);
      while (marks) {
        const uint32_t u = g * 32 + (uint32_t)__ffs((int)marks) - 1u;
        marks &= marks - 1;
        const uint32_t wt0 = u * UNIT + lane * COMPACT_WPT;
        WordBits tb[COMPACT_WPT];
#pragma unroll
        for (int k = 0; k < COMPACT_WPT; k++)
          if (wt0 + k < c.n_words) tb[k] = word_load(c, wt0 + k);
        const uint32_t blk = u / WARPS;  // the reduce pass's block of this unit
        const Agg seed = agg_combine(agg_combine(cb.super_carry[blk / SCAN_THREADS], cb.block_carry[blk]),
                                     load_warp_prefix(cb.warp_agg, u));
        Agg tw[COMPACT_WPT];
        Agg mine = agg_zero();
#pragma unroll
        for (int k = 0; k < COMPACT_WPT; k++) {
          if (wt0 + k < c.n_words) {
            tw[k] = word_agg(wt0 + k, tb[k]);
            mine = agg_combine(mine, tw[k]);
          }
        }
        Agg run = warp_exclusive_scan(mine, seed);
#pragma unroll
        for (int k = 0; k < COMPACT_WPT; k++) {
          if (wt0 + k < c.n_words) {
            emit_texts(c, wt0 + k, tb[k], run);
            run = agg_combine(run, tw[k]);
          }
        }
      }
    }
    return;
  }
  const uint32_t w0 = (blockIdx.x * COMPACT_THREADS + threadIdx.x) * COMPACT_WPT;
  WordBits wb[COMPACT_WPT];
  Agg wa[COMPACT_WPT];
  Agg ta = agg_zero();
#pragma unroll
  for (int k = 0; k < COMPACT_WPT; k++) {
    const uint32_t w = w0 + k;
    if (w < c.n_words) {
      wb[k] = word_load(c, w);
      wa[k] = word_agg(w, wb[k]);
      ta = agg_combine(ta, wa[k]);
    }
  }
  // summary of everything before this block: (groups of 1024 blocks before) + (blocks before, in the group)
  const Agg block_start = agg_combine(cb.super_carry[blockIdx.x / SCAN_THREADS], cb.block_carry[blockIdx.x]);
  // prefix of this thread: warps before it in the block (their summary, left by the reduce pass), lanes before it
  const Agg seed = agg_combine(block_start, load_warp_prefix(cb.warp_agg, (size_t)blockIdx.x * (COMPACT_THREADS / 32) + (threadIdx.x >> 5)));
  Agg carry = warp_exclusive_scan(ta, seed);
  // K3_EMIT
  const uint32_t blk_tok0 = block_start.n_tok, blk_ntok = cb.block_agg[blockIdx.x].n_tok;
  const bool staged = blk_ntok <= STAGE_TOKENS;
  // staging layout: absolute form tok_bytes | tok_pos (2 + 2 words per token), compact form 2 words per token
  uint32_t* s_tb = s_stage;
  int32_t* s_tp = reinterpret_cast<int32_t*>(s_stage + 2 * STAGE_TOKENS);
  uint16_t* s_td = reinterpret_cast<uint16_t*>(s_stage);
  // specialised copies of the loop: staged ones address shared memory directly, compact / absolute form
#define DATOK_EMIT_LOOP(FORM, TB, TP, TD, BASE)                                                   \
  _Pragma("unroll") for (int k = 0; k < COMPACT_WPT; k++) {                                      \
    const uint32_t w = w0 + k;                                                                   \
    if (w < c.n_words) {                                                                         \
      if (wb[k].e | wb[k].s | wb[k].t) {                                                         \
        const WordMasks m = word_masks(wb[k], agg_last(carry));                                  \
        emit_tokens<FORM>(c, w, wb[k], m, carry, TB, TP, TD, BASE);                               \
        emit_sentences(c, w, wb[k], m, carry);                                                   \
      }                                                                                          \
      carry = agg_combine(carry, wa[k]);                                                         \
    }                                                                                            \
  }
  if (c.tok_delta8) {
    uint16_t* g8 = reinterpret_cast<uint16_t*>(c.tok_delta8);
    if (staged) { DATOK_EMIT_LOOP(2, nullptr, nullptr, s_td, blk_tok0) }
    else { DATOK_EMIT_LOOP(2, nullptr, nullptr, g8, 0u) }
  } else if (c.tok_delta) {
    if (staged) { DATOK_EMIT_LOOP(1, nullptr, nullptr, s_td, blk_tok0) }
    else { DATOK_EMIT_LOOP(1, nullptr, nullptr, c.tok_delta, 0u) }
  } else {
    if (staged) { DATOK_EMIT_LOOP(0, c.tok_bytes ? s_tb : nullptr, c.tok_pos ? s_tp : nullptr, nullptr, blk_tok0) }
    else { DATOK_EMIT_LOOP(0, c.tok_bytes, c.tok_pos, nullptr, 0u) }
  }
#undef DATOK_EMIT_LOOP
  if (!staged) return;
  __syncthreads();
  if (c.tok_bytes) {
    uint2* dst = reinterpret_cast<uint2*>(c.tok_bytes) + blk_tok0;
    const uint2* src = reinterpret_cast<const uint2*>(s_tb);
    for (uint32_t i = threadIdx.x; i < blk_ntok; i += COMPACT_THREADS) dst[i] = src[i];
  }
  if (c.tok_pos) {
    uint2* dst = reinterpret_cast<uint2*>(c.tok_pos) + blk_tok0;
    const uint2* src = reinterpret_cast<const uint2*>(s_tp);
    for (uint32_t i = threadIdx.x; i < blk_ntok; i += COMPACT_THREADS) dst[i] = src[i];
  }
  if (c.tok_delta) {
    uint2* dst = reinterpret_cast<uint2*>(c.tok_delta) + blk_tok0;
    const uint2* src = reinterpret_cast<const uint2*>(s_td);
    for (uint32_t i = threadIdx.x; i < blk_ntok; i += COMPACT_THREADS) dst[i] = src[i];
  }
  if (c.tok_delta8) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(c.tok_delta8) + blk_tok0;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(s_td);
    for (uint32_t i = threadIdx.x; i < blk_ntok; i += COMPACT_THREADS) dst[i] = src[i];
  }
}

// Scan of the block summaries, two levels: groups of SCAN_THREADS block summaries are scanned by one
// block each (compact_scan_groups), then the group totals by a single block (compact_scan_top).  A
// consumer block combines super_carry[group] with block_carry[block].
__global__ void __launch_bounds__(SCAN_THREADS) compact_scan_groups_kernel(CompactBuffers cb) {
  const uint32_t i = blockIdx.x * SCAN_THREADS + threadIdx.x;
  const Agg mine = i < cb.n_blocks ? cb.block_agg[i] : agg_zero();
  const Agg excl = block_exclusive_scan<SCAN_THREADS>(mine, agg_zero());
  if (i < cb.n_blocks) cb.block_carry[i] = excl;
  if (threadIdx.x == SCAN_THREADS - 1) cb.super_agg[blockIdx.x] = agg_combine(excl, mine);
}
__global__ void __launch_bounds__(SCAN_THREADS) compact_scan_top_kernel(CompactBuffers cb, bool sentence_end_in) {
  const uint32_t n = (cb.n_blocks + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t lo = threadIdx.x * per;
  const uint32_t hi = lo + per < n ? lo + per : n;
  Agg mine = agg_zero();
  for (uint32_t i = lo; i < hi; i++) mine = agg_combine(mine, cb.super_agg[i]);
  const Agg start = agg_stream_start(sentence_end_in);
  Agg run = block_exclusive_scan<SCAN_THREADS>(mine, start);
  for (uint32_t i = lo; i < hi; i++) {
    cb.super_carry[i] = run;
    run = agg_combine(run, cb.super_agg[i]);
  }
  // the thread that owns the last group holds the stream summary
  if ((lo < hi && hi == n) || (n == 0 && threadIdx.x == 0)) cb.total[0] = run;
}

__global__ void compact_finalize_kernel(CompactCtx c, CompactBuffers cb, bool text_end_in, bool final_input) {
  const StreamTotals t = finalize_stream(c, cb.total[0], text_end_in, final_input);
  *reinterpret_cast<StreamTotals*>(&cb.total[1]) = t;
}

void launch_compact_reduce(const CompactCtx& c, const CompactBuffers& cb, cudaStream_t s) {
  // one CTA per block of words: measured faster than a persistent grid of 8 CTAs per SM looping over the
  // blocks (0.26 against 0.31 ms per GiB); the kernel's loop then runs once
  compact_reduce_kernel<<<cb.n_blocks, REDUCE_THREADS, 0, s>>>(c, cb);
}
void launch_compact_scan(const CompactBuffers& cb, bool sentence_end_in, cudaStream_t s) {
  const uint32_t groups = (cb.n_blocks + SCAN_THREADS - 1) / SCAN_THREADS;
  if (groups) compact_scan_groups_kernel<<<groups, SCAN_THREADS, 0, s>>>(cb);
  compact_scan_top_kernel<<<1, SCAN_THREADS, 0, s>>>(cb, sentence_end_in);
}
void launch_compact_texts(const CompactCtx& c, const CompactBuffers& cb, cudaStream_t s) {
  // (any grid works: the warps stride over groups of 32 warp units)
  constexpr uint32_t WARPS = COMPACT_THREADS / 32;
  const uint32_t n_groups = (cb.n_blocks * WARPS + 31) / 32;
  uint32_t blocks = (n_groups + WARPS - 1) / WARPS;
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  if (blocks == 0) blocks = 1;  // (docs[0] is written by block 0)
  compact_kernel<K3_TEXTS><<<blocks, COMPACT_THREADS, 0, s>>>(c, cb);
}
int launch_compact_emit(const CompactCtx& c, const CompactBuffers& cb, cudaStream_t s) {
  const int smem_max = 4 * STAGE_TOKENS * (int)sizeof(uint32_t);
  const int smem = (c.tok_delta8 ? 1 : c.tok_delta ? 2 : 4) * STAGE_TOKENS * (int)sizeof(uint32_t);
  {
    cudaError_t e = cudaFuncSetAttribute(compact_kernel<K3_EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) return (int)e;
  }
  compact_kernel<K3_EMIT><<<cb.n_blocks, COMPACT_THREADS, smem, s>>>(c, cb);
  return (int)cudaGetLastError();
}
void launch_compact_finalize(const CompactCtx& c, const CompactBuffers& cb, bool text_end_in, bool final_input,
                             cudaStream_t s) {
  compact_finalize_kernel<<<1, 1, 0, s>>>(c, cb, text_end_in, final_input);
}

}  // namespace datok
