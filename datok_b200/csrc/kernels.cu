// kernels.cu -- sm_100a kernels of the Datok matrix-FSA transduction path.
//
//   K1  classify_kernel      bytes -> class stream + rune-start bitmap   (matrix.go:388-435)
//   K2a walk_spec_kernel     speculative per-chunk walk                   (matrix.go:437-695)
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

namespace datok {

// ------------------------------------------------------------------ K1

// One thread per input byte; a warp's 32 rune-start flags become one bitmap word.
__global__ void __launch_bounds__(256) classify_kernel(DeviceModel m, WalkBuffers b) {
  __shared__ uint8_t s_ascii[128];
  __shared__ uint8_t s_latin1[128];
  if (threadIdx.x < 128) {
    s_ascii[threadIdx.x] = m.cls.ascii_cls[threadIdx.x];
    s_latin1[threadIdx.x] = m.cls.latin1_cls[threadIdx.x];
  }
  __syncthreads();
  ClsTables T = m.cls;
  T.ascii_cls = s_ascii;
  T.latin1_cls = s_latin1;
  const uint32_t total = b.n_words * 32u;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    bool start = false, invalid = false;
    if (p < b.N) b.cls[p] = (uint8_t)classify_pos(b.in, b.N, p, T, &start, &invalid);
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, start);
    const uint32_t inv = __ballot_sync(0xFFFFFFFFu, invalid);
    if ((threadIdx.x & 31) == 0) {
      b.rstart[p >> 5] = word;
      if (inv) atomicOr(&b.counters[2], 1u);
    }
  }
}

void launch_classify(const DeviceModel& m, const WalkBuffers& b, cudaStream_t s) {
  const uint32_t total = b.n_words * 32u;
  uint32_t blocks = (total + 255) / 256;
  const uint32_t cap = 148u * 8u * 16u;
  if (blocks > cap) blocks = cap;
  classify_kernel<<<blocks, 256, 0, s>>>(m, b);
}

// ------------------------------------------------------------------ K2

constexpr int WALK_THREADS = 128;

__global__ void __launch_bounds__(WALK_THREADS) walk_spec_kernel(DeviceModel m, WalkBuffers b, uint32_t start_state) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < b.n_chunks) chunk_spec(m, b, i, start_state);
}

__global__ void __launch_bounds__(WALK_THREADS) stitch_kernel(DeviceModel m, WalkBuffers b, const uint32_t* list,
                                                              uint32_t n_list) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_list) return;
  const uint32_t i = list ? list[k] : k + 1;
  if (chunk_stitch(m, b, i)) {
    const uint32_t slot = atomicAdd(&b.counters[1], 1u);
    b.list_rewalk[slot] = i;
  }
}

__global__ void __launch_bounds__(WALK_THREADS) rewalk_kernel(DeviceModel m, WalkBuffers b, uint32_t n_rewalk) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  // n_rewalk is only the launch bound; the list length was counted on the device by stitch_kernel
  if (k < n_rewalk && k < b.counters[1]) chunk_rewalk(m, b, b.list_rewalk[k]);
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

void launch_walk_spec(const DeviceModel& m, const WalkBuffers& b, uint32_t start_state, cudaStream_t s) {
  walk_spec_kernel<<<(b.n_chunks + WALK_THREADS - 1) / WALK_THREADS, WALK_THREADS, 0, s>>>(m, b, start_state);
}
void launch_stitch(const DeviceModel& m, const WalkBuffers& b, const uint32_t* list, uint32_t n_list, cudaStream_t s) {
  if (!n_list) return;
  stitch_kernel<<<(n_list + WALK_THREADS - 1) / WALK_THREADS, WALK_THREADS, 0, s>>>(m, b, list, n_list);
}
void launch_rewalk(const DeviceModel& m, const WalkBuffers& b, uint32_t n_rewalk, cudaStream_t s) {
  if (!n_rewalk) return;
  rewalk_kernel<<<(n_rewalk + WALK_THREADS - 1) / WALK_THREADS, WALK_THREADS, 0, s>>>(m, b, n_rewalk);
}
void launch_commit(const WalkBuffers& b, const uint32_t* list, uint32_t n_list, cudaStream_t s) {
  if (!n_list) return;
  commit_kernel<<<(n_list + 255) / 256, 256, 0, s>>>(b, list, n_list);
}
void launch_collect_errors(const WalkBuffers& b, cudaStream_t s) {
  collect_errors_kernel<<<(b.n_chunks + 255) / 256, 256, 0, s>>>(b);
}

// ------------------------------------------------------------------ K3

union AggWords {
  Agg a;
  uint32_t w[AGG_WORDS];
};
static_assert(sizeof(Agg) == AGG_WORDS * 4, "Agg layout");

__device__ __forceinline__ Agg agg_shfl_up(const Agg& v, int delta) {
  AggWords in, out;
  in.a = v;
#pragma unroll
  for (int k = 0; k < AGG_WORDS; k++) out.w[k] = __shfl_up_sync(0xFFFFFFFFu, in.w[k], delta);
  return out.a;
}

// Ordered (non-commutative) block scan.  Returns the exclusive prefix of `mine`
// within the block combined after `seed`; *block_total = all threads' values combined.
__device__ __forceinline__ Agg block_exclusive_scan(const Agg& mine, const Agg& seed, Agg* block_total) {
  __shared__ Agg s_warp[COMPACT_THREADS / 32];
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
  if (block_total) {
    Agg tot = s_warp[0];
    for (int wi = 1; wi < COMPACT_THREADS / 32; wi++) tot = agg_combine(tot, s_warp[wi]);
    *block_total = tot;
  }
  Agg prev = agg_shfl_up(incl, 1);
  if (lane > 0) prefix = agg_combine(prefix, prev);
  __syncthreads();
  return prefix;
}

__global__ void __launch_bounds__(COMPACT_THREADS) compact_reduce_kernel(CompactCtx c, CompactBuffers cb) {
  const uint32_t w0 = (blockIdx.x * COMPACT_THREADS + threadIdx.x) * COMPACT_WPT;
  Agg ta = agg_zero();
#pragma unroll
  for (int k = 0; k < COMPACT_WPT; k++) {
    const uint32_t w = w0 + k;
    if (w < c.n_words) ta = agg_combine(ta, process_word<false>(c, w, ta));
  }
  Agg tot;
  block_exclusive_scan(ta, agg_zero(), &tot);
  if (threadIdx.x == 0) cb.block_agg[blockIdx.x] = tot;
}

// Single block: thread t owns a contiguous run of block summaries.
__global__ void __launch_bounds__(COMPACT_THREADS) compact_scan_kernel(CompactCtx c, CompactBuffers cb,
                                                                        bool sentence_end_in) {
  const uint32_t per = (cb.n_blocks + COMPACT_THREADS - 1) / COMPACT_THREADS;
  const uint32_t lo = threadIdx.x * per;
  const uint32_t hi = lo + per < cb.n_blocks ? lo + per : cb.n_blocks;
  Agg mine = agg_zero();
  for (uint32_t i = lo; i < hi; i++) mine = agg_combine(mine, cb.block_agg[i]);
  Agg tot;
  Agg run = block_exclusive_scan(mine, agg_stream_start(c, sentence_end_in), &tot);
  for (uint32_t i = lo; i < hi; i++) {
    cb.block_carry[i] = run;
    run = agg_combine(run, cb.block_agg[i]);
  }
  if (threadIdx.x == 0) cb.total[0] = agg_combine(agg_stream_start(c, sentence_end_in), tot);
}

__global__ void __launch_bounds__(COMPACT_THREADS) compact_emit_kernel(CompactCtx c, CompactBuffers cb) {
  const uint32_t w0 = (blockIdx.x * COMPACT_THREADS + threadIdx.x) * COMPACT_WPT;
  Agg wa[COMPACT_WPT];
  Agg ta = agg_zero();
#pragma unroll
  for (int k = 0; k < COMPACT_WPT; k++) {
    const uint32_t w = w0 + k;
    wa[k] = (w < c.n_words) ? process_word<false>(c, w, ta) : agg_zero();
    ta = agg_combine(ta, wa[k]);
  }
  Agg carry = block_exclusive_scan(ta, cb.block_carry[blockIdx.x], nullptr);
#pragma unroll
  for (int k = 0; k < COMPACT_WPT; k++) {
    const uint32_t w = w0 + k;
    if (w < c.n_words) {
      process_word<true>(c, w, carry);
      carry = agg_combine(carry, wa[k]);
    }
  }
}

__global__ void compact_finalize_kernel(CompactCtx c, CompactBuffers cb, bool text_end_in) {
  Agg tot = cb.total[0];
  finalize_stream(c, tot, text_end_in);
  cb.total[1] = tot;
}

void launch_compact_reduce(const CompactCtx& c, const CompactBuffers& cb, cudaStream_t s) {
  compact_reduce_kernel<<<cb.n_blocks, COMPACT_THREADS, 0, s>>>(c, cb);
}
void launch_compact_scan(const CompactCtx& c, const CompactBuffers& cb, bool sentence_end_in, cudaStream_t s) {
  compact_scan_kernel<<<1, COMPACT_THREADS, 0, s>>>(c, cb, sentence_end_in);
}
void launch_compact_emit(const CompactCtx& c, const CompactBuffers& cb, cudaStream_t s) {
  compact_emit_kernel<<<cb.n_blocks, COMPACT_THREADS, 0, s>>>(c, cb);
}
void launch_compact_finalize(const CompactCtx& c, const CompactBuffers& cb, bool text_end_in, cudaStream_t s) {
  compact_finalize_kernel<<<1, 1, 0, s>>>(c, cb, text_end_in);
}

}  // namespace datok
