// Host launcher of the fused MLP activation (see activation.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace magi {

// x: (rows, 2 f) bf16 when gated (out = silu(x[:, :f]) * x[:, f:]), (rows, f) otherwise (out = gelu(x)); out: (rows, f)
cudaError_t mlp_act_launch(const void* x, int64_t x_row_stride, void* out, int64_t out_row_stride, int64_t rows, int f,
                           bool gated, int sms, cudaStream_t stream);

}  // namespace magi
