// format_kernels.cuh -- launch interface of the device formatter (format_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "format_core.cuh"

namespace datok {

// device scratch the prefix sums need for these counts
size_t format_scratch_bytes(uint32_t n_tok, uint32_t n_sentpos, uint32_t n_text);
// points the four prefix sums of c into `scratch` (format_scratch_bytes)
void format_carve(FmtCtx& c, uint8_t* scratch);
// phase 1: the prefix sums over the item lengths; the length of the text goes to *d_total
int launch_format_scan(const FmtCtx& c, unsigned long long* d_total, cudaStream_t s);
// phase 2: every token / SentenceEnd / TextEnd / `sent` entry writes its bytes to c.out (which holds *d_total bytes)
int launch_format_write(const FmtCtx& c, cudaStream_t s);

}  // namespace datok
