/*
 * magi_attn.h - C-ABI of the B200-native MAGI-1 attention hot path (libmagi_attn.so).
 *
 * Plain pointers and sizes only: no torch types cross this boundary. Every entry point is asynchronous on the
 * `stream` it is given (a cudaStream_t passed as void*), never synchronises the device, never allocates persistent
 * device memory, and is CUDA-graph capturable. All return an int status: 0 = ok, non-zero = error (the message is
 * available from magi_last_error(), thread-local).
 *
 * The reference (samiazirar/MAGI-1) has no FFI of its own: the "operator API" of this path is a set of Python call
 * sites inside inference/model/dit/dit_module.py. Each entry point below names the call site(s) it replaces.
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Layout conventions (reference: dit_module.py:953-996, 1023-1066):
 *   q / out : (tokens, Hq, 128) bf16, explicit token/head strides in elements, innermost dim contiguous
 *   k / v   : (tokens, Hkv, 128) bf16 with explicit strides, so the reference's packed KV tensor
 *             (tokens, Hkv, 256) = [K(128) | V(128)] (dit_module.py:996) is consumed in place:
 *             k = base, v = base + 128, token_stride = Hkv*256, head_stride = 256.
 *   GQA     : q head h reads kv head h / (Hq / Hkv) (flash-attn convention).
 */
#ifndef MAGI_ATTN_H_
#define MAGI_ATTN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MAGI_API __attribute__((visibility("default")))
#else
#define MAGI_API
#endif

#define MAGI_HEAD_DIM 128
#define MAGI_ROPE_DIM 96
#define MAGI_MAX_PEERS 8 /* ranks of one context-parallel group (one NVSwitch box) */

/* A (tokens, heads, 128) bf16 view. Strides are in elements. */
typedef struct magi_view {
  const void* ptr;
  int64_t tokens;
  int64_t token_stride;
  int64_t head_stride;
} magi_view;

/*
 * Range-masked ("chunk-causal") attention forward.
 * Replaces: flex_flash_attn_func      dit_module.py:1032-1043, 1074-1085, 1111-1122
 *           flash_attn_func loop      dit_module.py:1046-1065, 1087-1095
 *           flash_attn_varlen_func    dit_module.py:1124-1133   (give cu_seqlens_* instead of *_ranges)
 *           torch.cat/chunk of cache  dit_module.py:914-931, 945-946 (two-segment K/V source, no copy)
 *
 * For each i < n_ranges, query rows [q_ranges[i][0], q_ranges[i][1]) attend densely (non-causally) to key rows
 * [k_ranges[i][0], k_ranges[i][1]) of the *concatenation* seg_a ++ seg_b. seg_a is the KV-cache prefix, seg_b the
 * new tokens of this forward; either may be empty (tokens == 0). Query ranges must be disjoint (the reference passes
 * disable_fwd_atomic_reduction=True, dit_module.py:1042). Rows covered by no range are left untouched.
 * A range with an empty key interval yields out = 0, lse = -inf.
 * K/V are read in tiles of 128 rows: a tile may extend past the end of a key range (those scores are masked) and past
 * the end of a VIEW (rows beyond `tokens` are zero-filled by the TMA unit, never read). Rows that lie inside a view
 * but outside every key range must therefore hold finite values - a masked probability of exactly 0 times a NaN / Inf
 * in V would still poison the accumulator (pass `cache[:prefix_rows]`, not the whole torch.empty allocation).
 *
 * Exactly one of {q_ranges,k_ranges} / {cu_seqlens_q,cu_seqlens_k} must be non-NULL; all are int32 DEVICE pointers
 * ((n,2) row-major, resp. n+1 entries). lse is (total_q, Hq) fp32 natural-log, may be NULL.
 */
typedef struct magi_attn_args {
  magi_view q;
  magi_view k_a, v_a; /* segment A: KV-cache prefix (tokens may be 0) */
  magi_view k_b, v_b; /* segment B: new tokens                         */
  void* out;
  int64_t out_token_stride, out_head_stride;
  float* lse;
  const int32_t* q_ranges;
  const int32_t* k_ranges;
  const int32_t* cu_seqlens_q;
  const int32_t* cu_seqlens_k;
  int32_t n_ranges;
  int32_t num_heads_q, num_heads_kv, head_dim;
  float softmax_scale; /* <= 0 selects 1/sqrt(head_dim) (reference passes softmax_scale=None) */
  void* workspace;     /* >= magi_attn_workspace_bytes(n_ranges) bytes of device memory        */
  size_t workspace_bytes;
  int32_t num_sms;     /* 0 = all SMs of the current device                                    */
  /* Optional grouped output addressing (0 = off): head h is stored at
   *   out + row * out_token_stride + (h / out_group_heads) * out_group_stride + (h % out_group_heads) * out_head_stride
   * which lets self- and cross-attention write straight into the out-projection input of the reference,
   * "sq b (n hn hd) -> sq b (hn n hd)", n = 2, hn = 8 (attn_linear_proj, dit_module.py:1332-1338): no cat, no
   * rearrange copy. */
  int32_t out_group_heads;
  int64_t out_group_stride;
  /* Fused Ulysses output all-to-all (num_peers > 0): query row g (a row of ALL ranks' tokens) of local head h is
   * stored straight into the output buffer of the rank rho that owns the row, peer_row_bounds[rho] <= g <
   * peer_row_bounds[rho + 1], at
   *   peer_out[rho] + (g - peer_row_bounds[rho]) * out_token_stride + (peer_head_offset + h) * out_head_stride
   * i.e. all_to_all_output_split followed by "(cp sq) hn hd -> sq (cp hn) hd" (context_parallel.py:404-423, 596)
   * performed by the epilogue's own stores over NVLink peer memory. `out` is then unused (may be NULL); rows outside
   * every bound are dropped. The caller owns the cross-GPU synchronisation.
   * Together with out_group_heads > 0 the grouped addressing above applies to the destination head
   * H = peer_head_offset + h: (H / out_group_heads) * out_group_stride + (H % out_group_heads) * out_head_stride, i.e.
   * the rows land directly in the owner's out-projection input (single-CTA kernel only). */
  int32_t num_peers;
  int32_t peer_head_offset;
  void* peer_out[MAGI_MAX_PEERS];
  int64_t peer_row_bounds[MAGI_MAX_PEERS + 1];
  /* Which kernel runs: 0 = the library default (the environment variable MAGI_ATTN_KERNEL=single|pair overrides it),
   * 1 = attn_fwd_kernel (one CTA per SM, two query tiles per CTA), 2 = attn_pair_kernel (CTA pairs, cta_group::2;
   * experimental: correct, measured slower - DESIGN.md section 4.1). Same results within the documented tolerance. */
  int32_t kernel;
  /* plan_ready != 0: `workspace` already holds the work list magi_attn_plan built for exactly these ranges,
   * num_heads_q and token counts; magi_attn_fwd then launches the attention kernel only (one launch instead of two).
   * The ranges of a denoise forward are the same for every layer (dit_model.py:215-259): plan once, reuse 34 / 48 times. */
  int32_t plan_ready;
  /* Launch ordering against the kernel BEFORE this one in the stream. The attention and prologue kernels are launched
   * with programmatic stream serialization (their CTA set-up overlaps the tail of the earlier kernel; MAGI_PDL=0 in the
   * environment restores plain launches). 0 = this launch may depend on that kernel: its memory accesses start once
   * that kernel has completed. 1 = the caller asserts that NOTHING this launch reads or writes (q, k, v, out, lse,
   * ranges, the plan in `workspace` - hence plan_ready must be set) is read or written by that kernel: the CTAs of
   * this launch then run on the SMs the earlier kernel has already left, and the launch still COMPLETES in stream
   * order (everything after it sees both finished). The text cross-attention after the self-attention of the same
   * layer is such a launch (context_parallel.py:591: "no communication", no shared tensors). */
  int32_t launch_dependency;
  /* != 0: query rows that no q range covers are written as zeros (lse = -inf) by the kernel itself - the behaviour
   * flex_flash_attn_func / flash_attn_varlen_func callers rely on - instead of by a memset of the whole output before
   * the launch. Plain output addressing only (no peers, no groups). */
  int32_t fill_uncovered_rows;
} magi_attn_args;

MAGI_API size_t magi_attn_workspace_bytes(int32_t n_ranges);
MAGI_API int magi_attn_fwd(const magi_attn_args* args, void* stream);
/* Builds the work list of magi_attn_fwd (ranges clamped to the token counts, sorted longest-key-range first, prefix sums
 * of the 256-row work items) into args->workspace, on `stream`, without host reads. Uses only the range pointers,
 * n_ranges, num_heads_q, q.tokens, k_a.tokens + k_b.tokens and workspace / workspace_bytes of `args`. */
MAGI_API int magi_attn_plan(const magi_attn_args* args, void* stream);

/*
 * Fused QK-LayerNorm(+1p) + RoPE-96 + layout prologue.
 * Replaces: get_q / get_k / get_v / get_kv   dit_module.py:953-996 (7 eager kernels per tensor)
 *           FusedLayerNorm.forward           dit_module.py:355-357
 *           flash_attn rotary (Triton)       dit_module.py:961, 978
 *           KV-cache slice-assign            dit_module.py:918-929 (append fused: second store target)
 *
 * x_q / x_k / x_v are the bf16 projection outputs, (rows, H, 128) with explicit row/head strides, rows = (sq b).
 * LayerNorm runs in fp32 over 128 with gamma = (w + 1) if zero_centered_gamma else w, bias b, eps; weights are
 * fp32 (the reference promotes q/k_layernorm to fp32, dit_model.py:621-633) or bf16 (ln_dtype = 1; the +1 is
 * then rounded to bf16 like the reference's `self.weight + 1` in bf16).
 * RoPE is the non-interleaved (NeoX) rotation of dims [0,48) with [48,96) in fp32 by sin/cos (seq, 48) fp32,
 * row r uses table row r / batch; dims [96,128) pass through. One rounding to bf16 at the end.
 * Passing sin == NULL skips RoPE (cross-attention q/k, dit_module.py:1005-1021).
 *
 * q_out: (rows, Hq, 128) with q_out strides. kv_out: packed (rows, Hkv, 256) rows, K at +0 and V at +128, with
 * kv_out_token_stride (elements). If cache != NULL, rows [0, cache_rows) are ALSO stored to
 * cache + (cache_row_offset + r) * cache_token_stride (the KV-cache append; cache_rows excludes the last chunk
 * when distill_nearly_clean_chunk, dit_module.py:921-925).
 * Head remap for Ulysses: if head_perm_cp > 1 the outputs are written in all-to-all SEND layout
 * "(cp seq) hn hd" (context_parallel.py:396): output row = (h / (H/cp)) * rows + r, head = h % (H/cp).
 */
typedef struct magi_prologue_args {
  const void* x_q; /* may be NULL (kv only)  */
  const void* x_k; /* may be NULL (q only)   */
  const void* x_v; /* may be NULL            */
  int64_t rows;
  int64_t xq_row_stride, xk_row_stride, xv_row_stride;    /* elements */
  int64_t xq_head_stride, xk_head_stride, xv_head_stride; /* elements; 0 selects 128 (dense heads). 256 reads K and
                                                             V out of a packed (rows, Hkv, 256) projection
                                                             (linear_kv_xattn, dit_module.py:1015-1018) */
  int32_t num_heads_q, num_heads_kv;
  int32_t batch; /* rope row = r / batch */
  const void* ln_q_w;
  const void* ln_q_b;
  const void* ln_k_w;
  const void* ln_k_b;
  int32_t ln_dtype; /* 0 = fp32, 1 = bf16 */
  float eps;
  int32_t zero_centered_gamma;
  const float* sin; /* (rows / batch, 48) or NULL */
  const float* cos;
  void* q_out;
  int64_t q_out_token_stride, q_out_head_stride;
  void* kv_out;
  int64_t kv_out_token_stride, kv_out_head_stride; /* head stride normally 256 */
  void* cache;
  int64_t cache_token_stride, cache_head_stride;
  int64_t cache_row_offset, cache_rows;
  int32_t head_perm_cp; /* 0/1 = identity */
  /* Fused Ulysses input all-to-all (head_perm_cp > 1, num_peers == head_perm_cp): instead of one local send buffer,
   * the slice of head group c is stored straight into rank c's RECEIVE buffer over NVLink peer memory,
   *   peer_q_out[c]  + (peer_row_offset + r) * q_out_token_stride  + (h % hn) * q_out_head_stride
   *   peer_kv_out[c] + (peer_row_offset + r) * kv_out_token_stride + (h % hn) * kv_out_head_stride (+128 for V)
   * i.e. where all_to_all_input_split (context_parallel.py:381-401) would have delivered it; peer_row_offset is
   * the sum of the split sizes of the ranks before this one. q_out / kv_out are then unused (may be NULL). The
   * caller owns the cross-GPU synchronisation (all ranks' launches complete -> barrier -> consumers start). */
  int32_t num_peers; /* 0 = off */
  void* peer_q_out[MAGI_MAX_PEERS];
  void* peer_kv_out[MAGI_MAX_PEERS];
  int64_t peer_row_offset;
  /* Fused K/V all-gather for the sequence-sharded strategy cp_ring_kv (gather_peers > 0; head_perm_cp <= 1, no cache,
   * num_peers == 0): q goes to q_out as usual; every K/V row r of this rank's stripe is stored into ALL
   * gather_peers ranks' gathered in-flight K/V buffers (this rank's own included), packed (rows, Hkv, 256) with the
   * kv_out strides, at row
   *   (r / gather_rows_per_chunk) * gather_chunk_stride + gather_row_offset + (r % gather_rows_per_chunk)
   * i.e. chunk-major: local rows are (chunk, stripe row), gather_chunk_stride = tokens per chunk, gather_row_offset =
   * first row of this rank's stripe inside a chunk. The stores are the transfer (NVLink peer memory); kv_out is
   * unused (may be NULL). Replaces the cp - 1 ring steps of K/V blocks (new capability: the reference has no
   * sequence-sharded attention). The caller owns the cross-GPU synchronisation. */
  int32_t gather_peers; /* 0 = off */
  void* gather_kv_out[MAGI_MAX_PEERS];
  int64_t gather_rows_per_chunk, gather_chunk_stride, gather_row_offset;
  /* gather_multicast != 0: gather_peers == 1 and gather_kv_out[0] is the NVSwitch MULTICAST mapping of the symmetric
   * buffer (one store, replicated to every rank by the switch). The kernel then stores with multimem.st - the PTX ISA
   * leaves ordinary ld/st to a multimem address undefined. */
  int32_t gather_multicast;
  /* Further tensors of the SAME launch - the cross-attention's half of a layer (get_xqkv, dit_module.py:1005-1021) - so
   * that a layer's prologue is one launch instead of three. Always stored in place whatever the routing of q / k / v
   * above; LayerNorm parameters bf16 (the `_xattn` norms are not promoted to fp32, dit_model.py:621-633), no RoPE.
   *   x_qx  (rows, num_heads_q, 128) view       -> qx_out  (rows, num_heads_q, 128)
   *   x_tk / x_tv (rows_t, num_heads_kv, 128) views (normally the two halves of the packed text projection,
   *   head stride 256)                          -> tkv_out (rows_t, num_heads_kv, 256) packed [K | V]
   * All NULL / 0 = off. */
  const void* x_qx;
  int64_t xqx_row_stride, xqx_head_stride; /* elements; head stride 0 selects 128 */
  const void* ln_qx_w; /* bf16 */
  const void* ln_qx_b;
  void* qx_out;
  int64_t qx_out_token_stride, qx_out_head_stride;
  const void* x_tk;
  const void* x_tv;
  int64_t rows_t;
  int64_t xt_row_stride, xt_head_stride;
  const void* ln_tk_w; /* bf16 */
  const void* ln_tk_b;
  void* tkv_out;
  int64_t tkv_out_token_stride, tkv_out_head_stride;
} magi_prologue_args;

MAGI_API int magi_qk_norm_rope_fwd(const magi_prologue_args* args, void* stream);

/*
 * Copy packed KV rows into the cache slot: cache[row_offset : row_offset + rows] = kv[0 : rows].
 * Replaces dit_module.py:929 when the new KV arrives from the Ulysses all-to-all rather than the prologue.
 */
MAGI_API int magi_kv_cache_append(void* cache, int64_t cache_token_stride, const void* kv, int64_t kv_token_stride,
                         int64_t row_offset, int64_t rows, int64_t row_elems, void* stream);

/*
 * Merge two partial attention results over disjoint key sets (ring-KV exchange step):
 *   lse = log(exp(lse_a) + exp(lse_b)); out = out_a * exp(lse_a - lse) + out_b * exp(lse_b - lse)
 * out_a/lse_a are updated in place. (rows, heads, 128) bf16 contiguous-innermost with token/head strides.
 * New capability (no reference call site; SURVEY.md section 2.2 "ring attention: no").
 */
MAGI_API int magi_lse_merge(void* out_a, float* lse_a, const void* out_b, const float* lse_b, int64_t rows, int32_t heads,
                   int64_t a_token_stride, int64_t a_head_stride, int64_t b_token_stride, int64_t b_head_stride,
                   void* stream);

/*
 * Head<->sequence re-layout used around the Ulysses all-to-all.
 *   mode 0: "seq (cp hn) hd -> (cp seq) hn hd"  (context_parallel.py:396, send side of all_to_all_input_split)
 *           kv_repeat > 1 replicates each source head kv_repeat times first (context_parallel.py:393-394);
 *           `heads` = source heads.
 *   mode 1: "(cp seq) hn hd -> seq (cp hn) hd"  (context_parallel.py:596, after all_to_all_output_split);
 *           `heads` = hn of src. The destination may be wider: dst row = (cp, dst_heads_total, hd) and this call
 *           fills heads [dst_head_offset, dst_head_offset + hn) of every cp block, so the per-head-group outputs of
 *           the overlapped pipeline (context_parallel.py:584-594) land in place without a torch.cat.
 *           dst_heads_total = 0 selects hn (dense).
 */
MAGI_API int magi_ulysses_relayout(void* dst, const void* src, int64_t seq, int32_t heads, int32_t head_elems,
                                   int32_t cp, int32_t kv_repeat, int32_t mode, int32_t dst_heads_total,
                                   int32_t dst_head_offset, void* stream);

/*
 * CSO (cp_shuffle_overlap) receive side: un-shuffle and un-pad the key/value sequence after its all-to-all,
 *   "(cp dn spb) hn nhd -> dn (cp spb) hn nhd" [:, :clip] flattened   (dit_module.py:1215-1224: rearrange + slice +
 *   .contiguous()), as one gather that can write straight into the KV-cache slot:
 *   dst[d * clip + t, :] = src[((t / spb) * dn + d) * spb + t % spb, :]   for d < dn, t < clip <= cp * spb.
 * Rows are row_elems bf16 (multiple of 8), token strides in elements.
 */
MAGI_API int magi_cso_unshuffle(void* dst, int64_t dst_token_stride, const void* src, int64_t src_token_stride,
                                int32_t cp, int32_t dn, int64_t spb, int64_t clip, int64_t row_elems, void* stream);

/*
 * Fused adaLN gating chain behind the attention path's output projection (and behind the MLP):
 *   out[row, :] = bf16( LayerNorm_hidden( float(x[row, :]) * float(gate[cond_map[row], :]) ) * (ln_w + 1) + ln_b
 *                       + float(residual[row, :]) )
 * Replaces bias_modulate_add (dit_module.py:292-307): x.float(), residual.float(), gate.float(), the Triton
 * range_mod kernel (:201-289), FusedLayerNorm post_norm (:355-357, fp32 parameters per dit_model.py:621-638), the
 * residual add and the cast back, seven eager kernels in the reference.
 * Rows are "(sq b)"-flattened; cond_map[row] indexes the rows of gate = gate.flatten(0, 1) (dit_model.py:204-206).
 * x_dtype / gate_dtype: 0 = fp32, 1 = bf16. residual / out: bf16. hidden must be a multiple of 256 and at most
 * 12288. Strides in elements; all pointers 16-byte aligned, row strides multiples of 8 elements.
 */
typedef struct MagiGatingArgs {
  const void* x; int32_t x_dtype; int64_t x_row_stride;
  const void* residual; int64_t residual_row_stride;
  const void* cond_map; int32_t cond_map_is_int64; int64_t cond_map_stride;
  const void* gate; int32_t gate_dtype; int64_t gate_row_stride; int64_t num_gates;
  const float* ln_w; const float* ln_b;
  float eps; int32_t zero_centered_gamma;
  void* out; int64_t out_row_stride;
  int64_t rows; int32_t hidden;
} MagiGatingArgs;
MAGI_API int magi_bias_modulate_add(const MagiGatingArgs* args, void* stream);

/*
 * Stand-alone non-interleaved (NeoX) rotary embedding, out of place (or in place: out == x with equal strides):
 *   out[..., :r/2] = x1 cos - x2 sin,  out[..., r/2:r] = x1 sin + x2 cos,  out[..., r:] = x[..., r:]
 * with x1 = x[..., :r/2], x2 = x[..., r/2:r], computed in fp32, one rounding to the storage type.
 * Replaces flash_apply_rotary_emb(q, cos_emb, sin_emb) = flash_attn.layers.rotary.apply_rotary_emb (Triton),
 * dit_module.py:961, 978, for callers that keep the reference's unfused get_q / get_k; the fused prologue
 * (magi_qk_norm_rope_fwd) does not need it.
 * x / out: (batch, seqlen, heads, head_dim), fp32 (dtype 0; what the reference passes: the transposed view of the fp32
 * LayerNorm output) or bf16 (dtype 1), explicit batch / sequence / head strides in elements, innermost dim contiguous.
 * cos / sin: (seqlen, rotary_dim / 2) fp32 contiguous; row s is applied to every batch row and head.
 * rotary_dim <= head_dim; rotary_dim / 2 and head_dim - rotary_dim multiples of 4; strides multiples of 4 (fp32) or
 * 8 (bf16) elements; all pointers 16-byte aligned.
 */
typedef struct magi_rotary_args {
  const void* x;
  void* out;
  int32_t dtype; /* 0 = fp32, 1 = bf16 (x and out) */
  int64_t batch, seqlen;
  int32_t heads, head_dim, rotary_dim;
  int64_t x_batch_stride, x_seq_stride, x_head_stride;
  int64_t out_batch_stride, out_seq_stride, out_head_stride;
  const float* cos;
  const float* sin;
} magi_rotary_args;
MAGI_API int magi_apply_rotary_emb(const magi_rotary_args* args, void* stream);

/*
 * q / k / v (/ qx) projections fused with the QK-LayerNorm + RoPE prologue: ONE tensor-core GEMM whose epilogue is
 * magi_qk_norm_rope_fwd.
 * Replaces: CustomLayerNormLinear.forward_q / forward_qx / forward_k / forward_v   dit_module.py:405-425 (4 GEMMs)
 *           + get_q / get_k / get_v / get_kv (:953-996) + the q half of get_xqkv (:1005-1009)
 *
 * x: (rows, hidden) bf16 = the LayerNorm'ed hidden states (forward_ln, dit_module.py:399-403), rows in "(sq b)" order.
 * w: (heads_total * 128, hidden) bf16, the projection weights stacked in the order [Wq; Wk; Wv; Wqx]
 *    (heads_total = Hq + 2 Hkv, + Hq when num_heads_qx == num_heads_q; must be even). nn.Linear layout: y = x w^T.
 * Every 128-wide head of y is rounded to bf16 (the Linear's output dtype) and then treated exactly like the
 * corresponding input of magi_qk_norm_rope_fwd: q and k heads LayerNorm (ln_dtype 0 = fp32 / 1 = bf16 parameters) +
 * RoPE, v heads pass-through into the packed kv_out (and the KV-cache slot, rows [0, cache_rows)), qx heads LayerNorm
 * with bf16 parameters and no RoPE into qx_out. hidden must be a multiple of 64; strides in elements (multiples of 8).
 */
typedef struct magi_qkv_proj_args {
  const void* x;
  int64_t rows, x_row_stride;
  int32_t hidden;
  const void* w;
  int64_t w_row_stride;
  int32_t num_heads_q, num_heads_kv, num_heads_qx; /* num_heads_qx: 0 or num_heads_q */
  int32_t batch;
  const void* ln_q_w;
  const void* ln_q_b;
  const void* ln_k_w;
  const void* ln_k_b;
  int32_t ln_dtype;
  const void* ln_qx_w; /* bf16 */
  const void* ln_qx_b;
  float eps;
  int32_t zero_centered_gamma;
  const float* sin;
  const float* cos;
  void* q_out;
  int64_t q_out_token_stride, q_out_head_stride;
  void* kv_out;
  int64_t kv_out_token_stride, kv_out_head_stride;
  void* qx_out;
  int64_t qx_out_token_stride, qx_out_head_stride;
  void* cache;
  int64_t cache_token_stride, cache_head_stride, cache_row_offset, cache_rows;
  int32_t num_sms; /* 0 = all SMs of the current device */
  /* Fused Ulysses input all-to-all (num_peers = cp > 1), the routing of magi_prologue_args.peer_*: head slice c of q and
   * of the packed [K|V] is stored straight into rank c's receive buffer - (all rows, num_heads_q / cp, 128) and
   * (all rows, max(1, num_heads_kv / cp), 256) - at row peer_row_offset + r, with KV-head replication when
   * cp > num_heads_kv (NVLink peer memory; the caller owns the cross-GPU barrier). q_out / kv_out / cache are then
   * unused (may be NULL); qx_out stays local. */
  int32_t num_peers;
  void* peer_q_out[MAGI_MAX_PEERS];
  void* peer_kv_out[MAGI_MAX_PEERS];
  int64_t peer_row_offset;
  /* fp8_quant (PerTensorQuantizedFp8Linear, dit_module.py:431-457: the q / qx / k / v projections of every layer but
   * the first and the last when engine_config.fp8_quant, e.g. example/4.5B/4.5B_distill_quant_config.json): fp8 != 0
   * means x and w hold float8_e4m3fn bytes - x as written by magi_fp8_quantize, w the stacked fp8 weights - strides
   * are in bytes = elements, hidden and both row strides must be multiples of 128 / 16, and every 128-wide head of
   * the fp32 accumulator is multiplied by out_scale[t] before the bf16 rounding, t = 0 q, 1 k, 2 v, 3 qx:
   * out_scale (DEVICE pointer to 4 floats) holds input_scale[0] * weight_scale[0] of each projection, the two scalars
   * bmm_fp8 hands to cuBLASLt (:450-457). The four projections must share one quantised x (equal input_scale). */
  int32_t fp8;
  const float* out_scale;
} magi_qkv_proj_args;
MAGI_API int magi_qkv_proj_norm_rope_fwd(const magi_qkv_proj_args* args, void* stream);

/*
 * Activation quantiser in front of the fp8 linears: `div_clamp_to` (dit_module.py:364-384), one memory-bound launch.
 *   out[r, c] = float8_e4m3fn( bfloat16( clamp( float(x[r, c]) / scale[scale_len == 1 ? 0 : c], -448, 448 ) ) )
 * with IEEE division and both roundings to nearest even, i.e. byte-identical to the reference's eager chain.
 * x (rows, cols) bf16, out (rows, cols) bytes, scale fp32 on the device with scale_len == 1 (scalar) or == cols
 * (per input channel: PerTensorQuantizedFp8Linear.input_scale is a vector of in_features, :443;
 * PerChannelQuantizedFp8Linear.smooth_scale is (1, in_features), :474). cols a multiple of 16, row strides in
 * elements (x: multiple of 8, out: multiple of 16), x / out / a per-channel scale 16-byte aligned.
 */
MAGI_API int magi_fp8_quantize(const void* x, int64_t x_row_stride, const float* scale, int32_t scale_len, void* out,
                               int64_t out_row_stride, int64_t rows, int32_t cols, void* stream);

/*
 * MLP activation between fc1 and fc2 (CustomMLP.forward, dit_module.py:541-549), one memory-bound launch:
 *   gated != 0 (24B, gated_linear_unit): out[r, c] = bf16(silu(float(x[r, c])) * float(x[r, f + c])), x = (rows, 2 f)
 *                 = flashinfer.activation.silu_and_mul (dit_module.py:546)
 *   gated == 0 (4.5B): out[r, c] = bf16(gelu(float(x[r, c]))) with the erf form of torch.nn.functional.gelu
 *                 (dit_module.py:548), x = (rows, f)
 * x / out bf16, f a multiple of 8, row strides in elements (multiples of 8), pointers 16-byte aligned; out may not
 * alias x when gated.
 */
MAGI_API int magi_mlp_activation(const void* x, int64_t x_row_stride, void* out, int64_t out_row_stride, int64_t rows,
                                 int32_t f, int32_t gated, void* stream);

/* Diagnostics */
MAGI_API const char* magi_last_error(void);
MAGI_API const char* magi_build_info(void);
/* Number of kernel launches issued by this library in this process (all threads). */
MAGI_API uint64_t magi_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MAGI_ATTN_H_ */
