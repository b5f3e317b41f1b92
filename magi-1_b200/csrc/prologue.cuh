// Kernel-side parameter block of the fused prologue and declarations of the memory-bound helper kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace magi {

struct PrologueParams {
  const void* x_q; const void* x_k; const void* x_v;
  int64_t rows;
  int64_t xq_row_stride, xk_row_stride, xv_row_stride;
  int64_t xq_head_stride, xk_head_stride, xv_head_stride;
  int Hq, Hkv, batch;
  const void* ln_q_w; const void* ln_q_b; const void* ln_k_w; const void* ln_k_b;
  float eps;
  int zero_centered_gamma;
  const float* sin; const float* cos;
  void* q_out; int64_t q_out_token_stride, q_out_head_stride;
  void* kv_out; int64_t kv_out_token_stride, kv_out_head_stride;
  void* cache; int64_t cache_token_stride, cache_head_stride, cache_row_offset, cache_rows;
  int head_perm_cp;
  // fused all-to-all: per destination rank receive buffers (same strides as q_out / kv_out), this rank's row offset
  int use_peers;
  void* peer_q[8]; void* peer_kv[8];
  int64_t peer_row_offset;
  // fused K/V all-gather (cp_ring_kv): peer_kv[0 .. gather_peers) are the ranks' gathered in-flight K/V buffers
  int gather_peers;
  int gather_multicast;   // peer_kv[0] is a multimem (NVSwitch multicast) address: store with multimem.st
  int64_t gather_rows_per_chunk, gather_chunk_stride, gather_row_offset;
  // Further tensors of the SAME launch (the cross-attention's half of a layer, get_xqkv dit_module.py:1005-1021): always
  // stored in place, LayerNorm parameters in bf16, no RoPE, whatever the mode of q / k / v above.
  //   qx: the cross-attention query, rows x Hq heads -> qx_out (rows, Hq, 128)
  //   tk / tv: the text K / V (rows_t x Hkv heads, normally two column halves of the packed projection)
  //            -> tkv_out (rows_t, Hkv, 256) packed [K | V]
  const void* x_qx; int64_t xqx_row_stride, xqx_head_stride;
  const void* ln_qx_w; const void* ln_qx_b;
  void* qx_out; int64_t qx_out_token_stride, qx_out_head_stride;
  const void* x_tk; const void* x_tv; int64_t rows_t, xt_row_stride, xt_head_stride;
  const void* ln_tk_w; const void* ln_tk_b;
  void* tkv_out; int64_t tkv_out_token_stride, tkv_out_head_stride;
};

// number of work units of a launch (-> *units) and the resident 256-thread CTAs per SM the kernel is built for
int prologue_units(const PrologueParams& p, int64_t* units);
// persistent launch sized from the SM count
cudaError_t prologue_launch(const PrologueParams& p, bool ln_bf16, int sms, cudaStream_t stream);
cudaError_t kv_append_launch(void* cache, int64_t cache_stride, const void* kv, int64_t kv_stride, int64_t row_offset,
                             int64_t rows, int row_vecs, int grid, cudaStream_t stream);
cudaError_t lse_merge_launch(void* out_a, float* lse_a, const void* out_b, const float* lse_b, int64_t rows, int heads,
                             int64_t a_ts, int64_t a_hs, int64_t b_ts, int64_t b_hs, int grid, cudaStream_t stream);
cudaError_t ulysses_relayout_launch(void* dst, const void* src, int64_t seq, int hn, int head_vecs, int cp, int rep,
                                    int mode, int dst_hn, int dst_off, int grid, cudaStream_t stream);
cudaError_t cso_unshuffle_launch(void* dst, int64_t dst_stride, const void* src, int64_t src_stride, int dn,
                                 int64_t spb, int64_t clip, int row_vecs, int grid, cudaStream_t stream);

}  // namespace magi
