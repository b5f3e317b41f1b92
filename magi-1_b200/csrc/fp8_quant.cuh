// Host launcher of the activation quantiser in front of the fp8 linears (see fp8_quant.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace magi {

cudaError_t fp8_quant_launch(const void* x, int64_t x_row_stride, const float* scale, int scale_len, void* out,
                             int64_t out_row_stride, int64_t rows, int cols, int sms, cudaStream_t stream);

}  // namespace magi
