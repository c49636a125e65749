// format_kernels.cu -- the text half of the TokenWriter on the device (format_core.cuh): prefix sums over the
// item lengths, then one thread per token / SentenceEnd / TextEnd / `sent` entry writes its bytes.
#include "format_kernels.cuh"

namespace datok {

namespace {

constexpr int FMT_THREADS = 256;
constexpr int FMT_IPT = FMT_TILE / FMT_THREADS;  // items per thread in a scan tile
static_assert(FMT_IPT * FMT_THREADS == (int)FMT_TILE, "tile shape");

enum { FMT_TOK = 0, FMT_POS = 1, FMT_SP = 2, FMT_TEXT = 3 };

template <int KIND>
__device__ __forceinline__ uint32_t fmt_value(const FmtCtx& c, uint32_t i) {
  if (KIND == FMT_TOK) return fmt_len_tok(c, i);
  if (KIND == FMT_POS) return fmt_len_pos(c, i);
  if (KIND == FMT_SP) return fmt_len_sp(c, i);
  return fmt_len_text(c, i);
}

// Tile-local exclusive prefix sums of the n item lengths (entry n: the tile's running total, so that P(n) is
// the grand total) and the tile totals.
template <int KIND>
__global__ void __launch_bounds__(FMT_THREADS) fmt_scan_local_kernel(FmtCtx c, uint32_t n, FmtScan out) {
  __shared__ uint32_t s_warp[FMT_THREADS / 32];
  const uint32_t tile0 = blockIdx.x * FMT_TILE, i0 = tile0 + threadIdx.x * FMT_IPT;
  uint32_t v[FMT_IPT], sum = 0;
#pragma unroll
  for (int k = 0; k < FMT_IPT; k++) {
    v[k] = (i0 + k < n) ? fmt_value<KIND>(c, i0 + k) : 0u;
    sum += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t before = incl - sum;
  for (int w = 0; w < warp; w++) before += s_warp[w];
#pragma unroll
  for (int k = 0; k < FMT_IPT; k++) {
    if (i0 + k <= n) out.local[i0 + k] = before;
    before += v[k];
  }
  if (threadIdx.x == FMT_THREADS - 1) out.base[blockIdx.x] = before;  // (the tile total; scanned in place next)
}

// exclusive scan of the tile totals, in place (one block)
__global__ void __launch_bounds__(1024) fmt_scan_base_kernel(unsigned long long* base, uint32_t n_tiles) {
  __shared__ unsigned long long s_part[1024];
  const uint32_t per = (n_tiles + 1023) / 1024, lo = threadIdx.x * per, hi = min(lo + per, n_tiles);
  unsigned long long sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += base[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < 1024; t++) { const unsigned long long x = s_part[t]; s_part[t] = run; run += x; }
  }
  __syncthreads();
  unsigned long long run = s_part[threadIdx.x];
  for (uint32_t i = lo; i < hi; i++) { const unsigned long long x = base[i]; base[i] = run; run += x; }
}

__global__ void __launch_bounds__(FMT_THREADS) fmt_write_tokens_kernel(FmtCtx c) {
  const uint32_t k = blockIdx.x * FMT_THREADS + threadIdx.x;
  if (k < c.n_tok) fmt_write_token(c, k);
}
__global__ void __launch_bounds__(FMT_THREADS) fmt_write_rest_kernel(FmtCtx c) {
  const uint32_t i = blockIdx.x * FMT_THREADS + threadIdx.x;
  if (i < c.n_sent) fmt_write_sentence(c, i);
  if (i < c.n_text) fmt_write_text(c, i);
  if (i < c.n_sentpos) fmt_write_sentpos(c, i);
}

template <int KIND>
void scan(const FmtCtx& c, uint32_t n, const FmtScan& out, cudaStream_t s) {
  const uint32_t tiles = n / FMT_TILE + 1;  // entries 0..n
  fmt_scan_local_kernel<KIND><<<tiles, FMT_THREADS, 0, s>>>(c, n, out);
  fmt_scan_base_kernel<<<1, 1024, 0, s>>>(out.base, tiles);
}

}  // namespace

size_t format_scratch_bytes(uint32_t n_tok, uint32_t n_sentpos, uint32_t n_text) {
  auto one = [](uint32_t n) { return (((size_t)n + 1) * 4 + 255) / 256 * 256 + (((size_t)n / FMT_TILE + 2) * 8 + 255) / 256 * 256; };
  return 2 * one(n_tok) + one(n_sentpos) + one(n_text) + 256;
}

void format_carve(FmtCtx& c, uint8_t* scratch) {
  auto carve = [&](uint32_t n, FmtScan& sc) {
    sc.local = reinterpret_cast<uint32_t*>(scratch);
    scratch += (((size_t)n + 1) * 4 + 255) / 256 * 256;
    sc.base = reinterpret_cast<unsigned long long*>(scratch);
    scratch += (((size_t)n / FMT_TILE + 2) * 8 + 255) / 256 * 256;
  };
  carve(c.n_tok, c.ptok);
  carve(c.n_tok, c.ppos);
  carve(c.n_sentpos, c.psp);
  carve(c.n_text, c.px);
}

__global__ void fmt_total_kernel(FmtCtx c, unsigned long long* total) { *total = fmt_total(c); }

int launch_format_scan(const FmtCtx& c, unsigned long long* d_total, cudaStream_t s) {
  scan<FMT_TOK>(c, c.n_tok, c.ptok, s);
  scan<FMT_POS>(c, c.n_tok, c.ppos, s);
  scan<FMT_SP>(c, c.n_sentpos, c.psp, s);
  scan<FMT_TEXT>(c, c.n_text, c.px, s);   // (needs ppos and psp)
  fmt_total_kernel<<<1, 1, 0, s>>>(c, d_total);
  return (int)cudaGetLastError();
}

int launch_format_write(const FmtCtx& c, cudaStream_t s) {
  if (c.n_tok) fmt_write_tokens_kernel<<<(c.n_tok + FMT_THREADS - 1) / FMT_THREADS, FMT_THREADS, 0, s>>>(c);
  const uint32_t m = max(max(c.n_sent, c.n_text), c.n_sentpos);
  if (m) fmt_write_rest_kernel<<<(m + FMT_THREADS - 1) / FMT_THREADS, FMT_THREADS, 0, s>>>(c);
  return (int)cudaGetLastError();
}

}  // namespace datok
