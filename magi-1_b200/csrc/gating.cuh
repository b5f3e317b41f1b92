// Kernel-side parameter block of the fused adaLN gating chain (bias_modulate_add).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace magi {

struct GatingParams {
  const void* x;          // (rows, hidden) fp32 or bf16
  const void* residual;   // (rows, hidden) bf16
  const void* cond_map;   // (rows) int32 or int64 index into gate rows
  const void* gate;       // (n_gates, hidden) fp32 or bf16
  const float* ln_w; const float* ln_b;   // (hidden) fp32
  void* out;              // (rows, hidden) bf16
  int64_t rows;
  int hidden;
  int64_t x_row_stride, res_row_stride, out_row_stride, gate_row_stride, map_stride;   // elements
  int map_is_i64;
  float eps;
  int zero_centered_gamma;
};

// warps that share one row for this hidden size (1, 2, 4 or 8), 0 if unsupported
int gating_warps_per_row(int hidden);
cudaError_t gating_launch(const GatingParams& p, bool x_bf16, bool gate_bf16, int sms, cudaStream_t stream);

}  // namespace magi
