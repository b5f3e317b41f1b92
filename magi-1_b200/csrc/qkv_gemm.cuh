// Kernel-side parameter block and host launcher of the fused projection GEMM + prologue (see qkv_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace magi {

struct QkvGemmParams {
  int64_t rows;                 // rows of X ("(sq b)" order)
  int hidden;                   // K (multiple of 64)
  int Hq, Hkv, heads_total;     // heads_total = Hq + 2 Hkv (+ Hq when the qx projection is included): rows of W / 128
  int batch;                    // rope row = row / batch
  const void* ln_q_w; const void* ln_q_b; const void* ln_k_w; const void* ln_k_b;   // fp32, or bf16 if ln_bf16
  const void* ln_qx_w; const void* ln_qx_b;                                         // bf16 (may be null without qx)
  int ln_bf16;
  float eps;
  int zero_centered_gamma;
  const float* sin; const float* cos;   // (rows / batch, 48) or null
  void* q_out; int64_t q_out_token_stride, q_out_head_stride;
  void* kv_out; int64_t kv_out_token_stride, kv_out_head_stride;
  void* qx_out; int64_t qx_out_token_stride, qx_out_head_stride;
  void* cache; int64_t cache_token_stride, cache_head_stride, cache_row_offset, cache_rows;
  // fused Ulysses input all-to-all (context parallel): head slice c of q / [K|V] goes straight into rank c's receive
  // buffer ((all rows, hq_loc, 128) / (all rows, hkv_loc, 256)) at row peer_row_offset + r - the same routing as the
  // prologue kernel's peer mode (magi_prologue_args.peer_*), with KV-head replication when cp > Hkv. qx stays local.
  int cp;                        // 0 / 1 = off
  int hq_loc, hkv_loc;
  void* peer_q[8]; void* peer_kv[8];
  int64_t peer_row_offset;
  int align32;                   // every destination row is 32-byte aligned: 32-byte stores
  // fp8_quant: X and W hold e4m3 bytes, hidden is a multiple of 128, and out_scale (device, 4 floats: q, k, v, qx)
  // holds input_scale * weight_scale of each projection
  int fp8;
  const float* out_scale;
};

cudaError_t qkv_gemm_launch(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const QkvGemmParams& p, int num_ctas,
                            cudaStream_t stream);

}  // namespace magi
