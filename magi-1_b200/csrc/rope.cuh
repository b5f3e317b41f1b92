// Stand-alone NeoX rotary embedding (the drop-in for the reference's flash_apply_rotary_emb call, for callers that do
// not use the fused prologue).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace magi {

struct RotaryParams {
  const void* x; void* out;
  int64_t batch, seqlen;
  int heads, head_dim, rotary_dim;
  int64_t x_bs, x_ss, x_hs;       // element strides of x: batch, sequence, head (innermost dim contiguous)
  int64_t o_bs, o_ss, o_hs;
  const float* cos; const float* sin;   // (seqlen, rotary_dim / 2) fp32, contiguous
};

cudaError_t rotary_launch(const RotaryParams& p, bool bf16, int grid, cudaStream_t stream);

}  // namespace magi
