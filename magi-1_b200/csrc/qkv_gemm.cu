// q / k / v / qx projections fused with the QK-LayerNorm + RoPE prologue:
//   Y = X * W^T            X (rows, hidden) bf16 = LayerNorm(hidden_states), W = [Wq; Wk; Wv; Wqx] (heads * 128, hidden)
//   per 128-wide head of Y: round to bf16 (the Linear's output dtype), then
//     q  heads: fp32 LayerNorm(+1p) over 128 -> NeoX RoPE-96 -> bf16 -> q_out (rows, Hq, 128)
//     k  heads: the same with the k parameters          -> kv_out[..., :128] (+ KV-cache slot)
//     v  heads: pass-through                            -> kv_out[..., 128:] (+ KV-cache slot)
//     qx heads: LayerNorm with bf16 parameters, no RoPE -> qx_out (rows, Hq, 128)
// Replaces CustomLayerNormLinear.forward_q/qx/k/v (inference/model/dit/dit_module.py:390-425) + get_q / get_k / get_v /
// get_kv (:953-996) + the q half of get_xqkv (:1005-1009): four cuBLAS GEMMs whose bf16 outputs (1.07 GB per layer at
// cfg2) make a round trip through HBM into the prologue kernel become ONE tcgen05 GEMM whose epilogue IS the prologue.
// With the TMEM fragment shape 32x32b one thread owns one row and a 128-column accumulator tile is exactly one head:
// LayerNorm statistics, the rotation partners (i, i + 48) and the 256-byte output row are all thread-local.
//
// CTA pair (cluster of 2, tcgen05.mma.cta_group::2, M = 256, N = 256, K = 16 per instruction): each CTA stages its 128
// rows of X and its 128 rows of W per 64-wide k block (32 KB per stage, 6 stages), and keeps two 256-column fp32
// accumulators in TMEM so that the epilogue of tile i overlaps the main loop of tile i + 1.
//   warp 0      TMA producer (both CTAs)
//   warp 1      MMA issuer (leader CTA) + TMEM allocator (both)
//   warps 2-5   epilogue: TMEM lane quarter = warp % 4, one thread per row, two heads per tile
//
// fp8_quant (PerTensorQuantizedFp8Linear, dit_module.py:431-457, the q / qx / k / v of every layer but the first and
// the last in the *_distill_quant configs): the same kernel instantiated with kFp8 = true. X and W are e4m3 bytes, a k
// block is 128 elements (still one 128-byte swizzle row, so staging, descriptors and the pipeline are unchanged), the
// instruction is tcgen05.mma.kind::f8f6f4 (K = 32 per instruction, twice the bf16 rate), and the epilogue multiplies
// the fp32 accumulator by input_scale * weight_scale of the head's projection (what bmm_fp8 / cuBLASLt does with its
// A / B scale pointers) before the bf16 rounding.
#include <math.h>
#include <stdio.h>

#include "common.cuh"
#include "qkv_gemm.cuh"

namespace magi {

namespace {

constexpr int kBM = 128;                 // rows of X per CTA (256 per pair)
constexpr int kBN = 256;                 // columns of Y per tile = 2 heads; each CTA stages 128 rows of W
constexpr int kBKBytes = 128;            // k block: one 128-byte swizzle row = 64 bf16 or 128 e4m3
constexpr int kABytes = kBM * kBKBytes;  // 16 KB
constexpr int kBBytes = 128 * kBKBytes;  // 16 KB
constexpr int kStageBytes = kABytes + kBBytes;
#ifndef MAGI_GEMM_STAGES
#define MAGI_GEMM_STAGES 6
#endif
constexpr int kStages = MAGI_GEMM_STAGES;
constexpr int kThreads = 192;
constexpr int kEpiWarps = 4;
constexpr uint32_t kIdescBf16 = make_idesc_bf16(256, 256, 0, 0);
constexpr uint32_t kIdescE4m3 = make_idesc_e4m3(256, 256);
constexpr int kSmemBytes = kStages * kStageBytes + 1024;

struct GemmBarriers {
  uint64_t empty[kStages];     // both CTAs (multicast commit): the MMAs that read the stage have completed
  uint64_t acc_full[2];        // both CTAs (multicast commit): accumulator b holds a finished tile
  uint64_t full[kStages];      // leader: 1 arrival (expect_tx) + the bytes of both CTAs
  uint64_t acc_empty[2];       // leader: 2 x 4 epilogue warps have accumulator b in registers
  uint32_t tmem_base;
  float gamma[3][128], beta[3][128];   // LayerNorm scale (w + 1 where zero-centred) and bias of q, k, qx
};

__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}

// Tile order: bands of kBandM row-tiles, inside a band column-tile major, row-tile minor - so that the tiles the CTA
// pairs work on at the same time (consecutive indices) share a few row-tiles of X and a few column-tiles of W that fit
// L2 together. Row-major order made every wave of pairs stream the WHOLE weight: at 24B shapes (W = 176 MB > L2) that
// was m_tiles x |W| = 18.7 GB of HBM reads per launch, more than the GEMM's own time.
constexpr int kBandM = 8;
__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int& mt, int& nt) {
  const int band = tile / (kBandM * n_tiles);
  const int m0 = band * kBandM;
  const int bh = min(kBandM, m_tiles - m0);          // height of this (possibly last, shorter) band
  const int local = tile - band * kBandM * n_tiles;
  nt = local / bh;
  mt = m0 + (local - nt * bh);
}

__device__ __forceinline__ float bf16_round_f(float x) {
  return __uint_as_float(static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(x))) << 16);
}

}  // namespace

template <bool kFp8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
qkv_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                const QkvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ GemmBarriers bar;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int pairs = static_cast<int>(gridDim.x) >> 1;
  const int n_tiles = p.heads_total / 2;                           // tiles along N (2 heads each)
  const int m_tiles = static_cast<int>((p.rows + 2 * kBM - 1) / (2 * kBM));
  const int total_tiles = m_tiles * n_tiles;
  constexpr int kBK = kFp8 ? kBKBytes : kBKBytes / 2;             // elements per k block
  const int k_blocks = p.hidden / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar.full[s], 1);
      mbar_init(&bar.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar.acc_full[b], 1);
      mbar_init(&bar.acc_empty[b], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(&bar.tmem_base, 512);
    tmem_relinquish_pair();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
  }
  // LayerNorm parameters as fp32 scale / bias in shared memory: q and k (fp32 or bf16 parameters), qx (bf16 parameters:
  // `weight + 1` is itself rounded to bf16 there, like ATen's bf16 add)
  for (int i = threadIdx.x; i < 3 * 128; i += kThreads) {
    const int which = i >> 7, c = i & 127;
    const void* w = which == 0 ? p.ln_q_w : which == 1 ? p.ln_k_w : p.ln_qx_w;
    const void* b = which == 0 ? p.ln_q_b : which == 1 ? p.ln_k_b : p.ln_qx_b;
    float g = 1.f, be = 0.f;
    if (w != nullptr) {
      const bool is_bf16 = which == 2 ? true : p.ln_bf16 != 0;
      if (is_bf16) {
        g = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(w)[c]);
        be = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(b)[c]);
        if (p.zero_centered_gamma) g = bf16_round_f(g + 1.f);
      } else {
        g = reinterpret_cast<const float*>(w)[c];
        be = reinterpret_cast<const float*>(b)[c];
        if (p.zero_centered_gamma) g += 1.f;
      }
    }
    bar.gamma[which][c] = g;
    bar.beta[which][c] = be;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bar.tmem_base;

  if (warp == 0) {
    // ====================================================================== TMA producer (both CTAs)
    uint32_t stage = 0, phase = 0;
    for (int tile = pair; tile < total_tiles; tile += pairs) {
      int mt, nt;
      tile_coords(tile, m_tiles, n_tiles, mt, nt);
      const int row0 = mt * 2 * kBM + rank * kBM;          // this CTA's 128 rows of X
      const int wrow0 = nt * kBN + rank * 128;             // this CTA's 128 rows of W (its half of the tile's N)
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait_warp(&bar.empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* dst = smem + stage * kStageBytes;
          if (rank == 0) mbar_expect_tx(&bar.full[stage], 2 * kStageBytes);
          tma_load_2d_pair(dst, &tm_x, &bar.full[stage], kb * kBK, row0);
          tma_load_2d_pair(dst + kABytes, &tm_w, &bar.full[stage], kb * kBK, wrow0);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ====================================================================== MMA issuer (leader CTA)
    if (rank == 0) {
      const uint32_t tb = uniform(tmem_base);
      const uint32_t base = smem_u32(smem);
      uint32_t stage = 0, phase = 0, t = 0;
      for (int tile = pair; tile < total_tiles; tile += pairs, ++t) {
        const uint32_t b = t & 1;
        mbar_wait_cluster(&bar.acc_empty[b], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait_cluster(&bar.full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = make_smem_desc_sw128(base + stage * kStageBytes, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(base + stage * kStageBytes + kABytes, 16, 1024);
            // four instructions per k block, each 32 bytes of K further into the swizzle row (K = 16 bf16 / 32 e4m3)
#pragma unroll
            for (int ks = 0; ks < kBKBytes / 32; ++ks) {
              if constexpr (kFp8)
                mma_ss_pair_f8(tb + b * 256, ad + static_cast<uint64_t>(ks * 2), bd + static_cast<uint64_t>(ks * 2),
                               kIdescE4m3, (kb > 0 || ks > 0) ? 1u : 0u);
              else
                mma_ss_pair(tb + b * 256, ad + static_cast<uint64_t>(ks * 2), bd + static_cast<uint64_t>(ks * 2),
                            kIdescBf16, (kb > 0 || ks > 0) ? 1u : 0u);
            }
            tc_commit_pair(&bar.empty[stage]);
            if (kb == k_blocks - 1) tc_commit_pair(&bar.acc_full[b]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ====================================================================== epilogue warps = the prologue
    const int quarter = warp & 3;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t acc_empty_remote = mapa_u32(&bar.acc_empty[0], 0);
    const int Hq = p.Hq, Hkv = p.Hkv;
    uint32_t t = 0;
    for (int tile = pair; tile < total_tiles; tile += pairs, ++t) {
      int mt, nt;
      tile_coords(tile, m_tiles, n_tiles, mt, nt);
      const uint32_t b = t & 1;
      const int64_t row = static_cast<int64_t>(mt) * 2 * kBM + rank * kBM + quarter * 32 + lane;
      const bool row_ok = row < p.rows;
      mbar_wait(&bar.acc_full[b], (t >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t u[128];
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) MAGI_TMEM_LD32(tmem_base + lane_sel + b * 256 + hh * 128 + c0, u, c0);
        tmem_wait_ld();
        if (hh == 1) {
          // both heads of this accumulator are in registers (or stored): hand it back to the issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_remote + 8 * b);
        }
        // head of Y -> which tensor. Order of W's rows: q heads, k heads, v heads, qx heads.
        const int head = nt * 2 + hh;
        int type, h;
        if (head < Hq) { type = 0; h = head; }
        else if (head < Hq + Hkv) { type = 1; h = head - Hq; }
        else if (head < Hq + 2 * Hkv) { type = 2; h = head - Hq - Hkv; }
        else { type = 3; h = head - Hq - 2 * Hkv; }
        float v[128];
        if constexpr (kFp8) {
          const float sc = __ldg(p.out_scale + type);   // input_scale * weight_scale of this head's projection
#pragma unroll
          for (int i = 0; i < 128; ++i) v[i] = bf16_round_f(__uint_as_float(u[i]) * sc);
        } else {
#pragma unroll
          for (int i = 0; i < 128; ++i) v[i] = bf16_round_f(__uint_as_float(u[i]));   // the Linear's bf16 output
        }
        if (type != 2) {
          const int set = type == 3 ? 2 : type;
          // fp32 LayerNorm over the head (two-pass, biased variance)
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int i = 0; i < 128; i += 4) { s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3]; }
          const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / 128.0f);
          s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
          for (int i = 0; i < 128; i += 4) {
            v[i] -= mean; v[i + 1] -= mean; v[i + 2] -= mean; v[i + 3] -= mean;
            s0 = fmaf(v[i], v[i], s0); s1 = fmaf(v[i + 1], v[i + 1], s1);
            s2 = fmaf(v[i + 2], v[i + 2], s2); s3 = fmaf(v[i + 3], v[i + 3], s3);
          }
          const float rstd = rsqrtf(((s0 + s1) + (s2 + s3)) * (1.0f / 128.0f) + p.eps);
#pragma unroll
          for (int i = 0; i < 128; ++i) v[i] = fmaf(v[i] * rstd, bar.gamma[set][i], bar.beta[set][i]);
          if (type != 3 && p.sin != nullptr && row_ok) {
            // NeoX rotation of dims [0, 48) with [48, 96) by this row's angles; dims [96, 128) pass through
            const int64_t trow = p.batch == 1 ? row : row / p.batch;
            const float4* cs = reinterpret_cast<const float4*>(p.cos + trow * 48);
            const float4* sn = reinterpret_cast<const float4*>(p.sin + trow * 48);
#pragma unroll
            for (int g = 0; g < 12; ++g) {
              const float4 c = __ldg(cs + g), s = __ldg(sn + g);
              const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x1 = v[4 * g + e], x2 = v[48 + 4 * g + e];
                v[4 * g + e] = fmaf(x1, cc[e], x2 * -ss[e]);
                v[48 + 4 * g + e] = fmaf(x2, cc[e], x1 * ss[e]);
              }
            }
          }
        }
        if (!row_ok) continue;
        // destinations of this head vector: one, plus the KV-cache slot, or - context parallel with fewer kv heads than
        // ranks - the receive buffers of every rank that holds a replica of the head
        __nv_bfloat16* dst[9];
        int nd = 1;
        const bool peers = p.cp > 1 && type != 3;
        if (peers) {
          const int64_t prow = p.peer_row_offset + row;
          if (type == 0) {
            dst[0] = reinterpret_cast<__nv_bfloat16*>(p.peer_q[h / p.hq_loc]) + prow * (p.hq_loc * 128) +
                     (h % p.hq_loc) * 128;
          } else {
            const int off = type == 2 ? 128 : 0;
            if (p.cp > Hkv) {       // repeat_interleave (context_parallel.py:393-394): rank c holds kv head c / rep
              const int rep = p.cp / Hkv;
              nd = rep;
              for (int c = 0; c < rep; ++c)
                dst[c] = reinterpret_cast<__nv_bfloat16*>(p.peer_kv[h * rep + c]) + prow * 256 + off;
            } else {
              dst[0] = reinterpret_cast<__nv_bfloat16*>(p.peer_kv[h / p.hkv_loc]) + prow * (p.hkv_loc * 256) +
                       (h % p.hkv_loc) * 256 + off;
            }
          }
        } else if (type == 0) {
          dst[0] = reinterpret_cast<__nv_bfloat16*>(p.q_out) + row * p.q_out_token_stride + h * p.q_out_head_stride;
        } else if (type == 3) {
          dst[0] = reinterpret_cast<__nv_bfloat16*>(p.qx_out) + row * p.qx_out_token_stride + h * p.qx_out_head_stride;
        } else {
          const int off = type == 2 ? 128 : 0;
          dst[0] = reinterpret_cast<__nv_bfloat16*>(p.kv_out) + row * p.kv_out_token_stride + h * p.kv_out_head_stride +
                   off;
          if (p.cache != nullptr && row < p.cache_rows)
            dst[nd++] = reinterpret_cast<__nv_bfloat16*>(p.cache) + (p.cache_row_offset + row) * p.cache_token_stride +
                        h * p.cache_head_stride + off;
        }
        // one thread owns one row: 32-byte stores write whole sectors (16-byte ones half a sector each, see the
        // attention epilogue); they need 32-byte aligned destination rows (checked on the host)
#pragma unroll
        for (int i = 0; i < 128; i += 16) {
          uint32_t o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = pack_bf16x2(v[i + 2 * k], v[i + 2 * k + 1]);
          for (int d = 0; d < nd; ++d) {
            if (p.align32) {
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst[d] + i), "r"(o[0]),
                           "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                           : "memory");
            } else {
              *reinterpret_cast<uint4*>(dst[d] + i) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(dst[d] + i + 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
        }
      }
    }
  }

  // stores into other GPUs' memory must be visible system-wide before the kernel (and the device barrier that follows
  // it in stream order) completes - see the note at the end of the prologue kernel
  if (p.cp > 1 && warp >= 2) __threadfence_system();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

cudaError_t qkv_gemm_launch(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const QkvGemmParams& p, int num_ctas,
                            cudaStream_t stream) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(qkv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(qkv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    if (dev < 64) attr_set[dev] = true;
  }
  num_ctas &= ~1;
  if (num_ctas < 2) num_ctas = 2;
  if (p.fp8)
    qkv_gemm_kernel<true><<<num_ctas, kThreads, kSmemBytes, stream>>>(tm_x, tm_w, p);
  else
    qkv_gemm_kernel<false><<<num_ctas, kThreads, kSmemBytes, stream>>>(tm_x, tm_w, p);
  return cudaGetLastError();
}

}  // namespace magi
