// kernels.cuh -- launch interface between the C ABI (api.cu) and the sm_100a
// kernels (kernels.cu).  Device pointers only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chunk_core.cuh"

namespace datok {

struct CompactBuffers {
  Agg* block_agg;
  Agg* block_carry;
  Agg* warp_agg;           // per block and warp: summary of the warps before it in the block (+ WAGG_HAS_TEXT mark)
  Agg* super_agg;          // one per group of SCAN_THREADS blocks
  Agg* super_carry;
  Agg* total;              // [0] stream summary after the scan, [1] StreamTotals after finalize
  uint32_t n_blocks;
};

#ifndef DATOK_COMPACT_THREADS
#define DATOK_COMPACT_THREADS 256
#endif
#ifndef DATOK_COMPACT_WPT
#define DATOK_COMPACT_WPT 2
#endif
#ifndef DATOK_REDUCE_WPT
#define DATOK_REDUCE_WPT 4  // words per thread of the reduce pass (measured: 0.25 -> 0.19 ms per GiB against 2)
#endif
constexpr int COMPACT_THREADS = DATOK_COMPACT_THREADS;
constexpr int COMPACT_WPT = DATOK_COMPACT_WPT;     // bitmap words per thread
constexpr int SCAN_THREADS = 1024;
constexpr int STAGE_TOKENS = 4096;  // tokens of one block staged in shared memory (else written directly)

// fused classify + speculative walk; returns a cudaError_t value
// `threads`: lanes per CTA of the fused walk (256, 512, 768 or 1024; DATOK_FUSED_THREADS, default 1024)
int fused_threads_from_env();
int launch_walk_fused(const DeviceModel& m, const WalkBuffers& b, uint32_t start_state, uint32_t n_hot, int n_sms,
                      cudaStream_t s, int threads);
size_t fused_smem_bytes(const DeviceModel& m, uint32_t n_hot, int threads);
uint32_t fused_max_hot_rows(const DeviceModel& m, size_t smem_limit, uint32_t n_states, int threads);
// measurement only: the dependent shared-memory gather chain of one byte step, nothing else (bench.py)
int launch_gather_bound(uint32_t n_rows, uint32_t row16, uint32_t segs, int n_sms, uint32_t* sink, cudaStream_t s);
// calibration histogram (visits per state, GPU numbering)
void launch_hist(const DeviceModel& m, const WalkBuffers& b, uint32_t* hist, uint32_t hist_cls_offset, cudaStream_t s);
// one fix-up round over `n_list` chunks (list == nullptr: all chunks 1..n_chunks-1)
void launch_stitch(const DeviceModel& m, const WalkBuffers& b, const uint32_t* list, uint32_t n_list, cudaStream_t s);
int launch_rewalk_fused(const DeviceModel& m, const WalkBuffers& b, uint32_t n_rewalk_max, uint32_t n_hot, int n_sms,
                        cudaStream_t s);
void launch_commit(const WalkBuffers& b, const uint32_t* list, uint32_t n_list, cudaStream_t s);
// one chain of dependent chunks, starting at list[0], followed on the device for up to max_steps chunks; leaves the
// next round's list (0 or 1 entries) in b.list_next / b.counters[0]
void launch_chain(const DeviceModel& m, const WalkBuffers& b, const uint32_t* list, uint32_t max_steps, cudaStream_t s);
void launch_collect_errors(const WalkBuffers& b, cudaStream_t s);

// up to four small device regions (32-bit words) -> mapped host memory at word offsets off[k]
struct MailSrc {
  const uint32_t* p[4];
  uint32_t words[4];
  uint32_t off[4];
};
void launch_mail(const MailSrc& src, uint32_t* dst_mapped, cudaStream_t s);

void launch_compact_reduce(const CompactCtx& c, const CompactBuffers& cb, cudaStream_t s);
void launch_compact_scan(const CompactBuffers& cb, bool sentence_end_in, cudaStream_t s);
void launch_compact_texts(const CompactCtx& c, const CompactBuffers& cb, cudaStream_t s);
int launch_compact_emit(const CompactCtx& c, const CompactBuffers& cb, cudaStream_t s);
void launch_compact_finalize(const CompactCtx& c, const CompactBuffers& cb, bool text_end_in, bool final_input,
                             cudaStream_t s);

}  // namespace datok
