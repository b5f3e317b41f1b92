// Kernel-side parameter block and host launchers of the range-masked attention kernel (see attn_fwd.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace magi {

struct AttnKernelParams {
  void* out;
  int64_t out_token_stride, out_head_stride;  // elements
  float* lse;                                 // (total_q, Hq) or nullptr
  const int* work;                            // produced by the setup kernel
  int Hq, Hkv;
  int len_a;                                  // rows of K/V segment A (KV-cache prefix)
  float scale_log2;                           // softmax_scale * log2(e)
  int out_group_heads;                        // 0 = plain head stride; else grouped output addressing (magi_attn.h)
  int64_t out_group_stride;
  // fused output all-to-all: row g of head h goes to peer_out[rho] (rho = owner of g), see magi_attn.h
  int num_peers;
  int peer_head_offset;
  void* peer_out[8];
  int64_t peer_bounds[9];
  // programmatic dependent launch: 0 = everything this launch touches may depend on the kernel before it in the stream
  // (griddepcontrol.wait right after the CTA's set-up); 1 = the caller asserts that nothing this launch reads or writes
  // is touched by that kernel (magi_attn_args.launch_dependency): the wait moves to the end of the kernel, so its CTAs
  // do their work on the SMs the earlier kernel has already left, and the launch still completes in stream order
  int pdl_wait_at_end;
  int* counters;                              // {next item, CTAs finished}, the last two ints of the plan workspace
  int out_align32;                            // every output row is 32-byte aligned: the epilogue may use STG.256
  int fill_uncovered;                         // rows covered by no q range get out = 0, lse = -inf (single-CTA kernel)
  int total_q;
};

// workspace ints: [total_items, n, 0, 0][prefix n+1][q0 n][q1 n][k0 n][k1 n][pad to 4][0, 0, next item, CTAs finished]
__host__ __device__ inline size_t attn_workspace_ints(int n) {
  return ((4 + (size_t)(n + 1) + 4 * (size_t)n + 3) & ~(size_t)3) + 4;
}

// Builds the LPT-ordered work list on the device from (q_ranges,k_ranges) or (cu_q,cu_k); no host sync.
cudaError_t attn_setup_launch(const int* q_ranges, const int* k_ranges, const int* cu_q, const int* cu_k, int n,
                              int Hq, int total_q, int total_k, int* ws, cudaStream_t stream);
// maps: {q, k_a, v_a, k_b, v_b}
// total_keys = rows of the whole key sequence (segments A + B): picks the scheduler mailbox flavour (attn_fwd.cu)
cudaError_t attn_fwd_launch(const CUtensorMap* maps, const AttnKernelParams& p, int num_ctas, int total_keys,
                            cudaStream_t stream);

// CTA-pair kernel (attn_pair.cu): maps {q (128-row box), k_a (64-row box), v_a (128), k_b (64), v_b (128)}; num_ctas is
// rounded down to an even number (one cluster of two CTAs per pair of SMs). Same params, work list and results.
cudaError_t attn_pair_launch(const CUtensorMap* maps, const AttnKernelParams& p, int num_ctas, cudaStream_t stream);

}  // namespace magi
