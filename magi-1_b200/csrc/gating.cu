// adaLN gating chain of the DiT block, fused: the consumer of the attention path's output projection
// (SURVEY.md section 8(f)-1).
//
// Reference semantics being replaced (samiazirar/MAGI-1, inference/model/dit/dit_module.py):
//   bias_modulate_add                         :292-307   x.float() -> range_mod -> post_norm -> + residual.float() -> cast
//   range_mod_triton / range_mod_kernel_fwd   :201-289   y[row, :] = x[row, :] * gatings[map[row], :]
//   FusedLayerNorm.forward (w + 1, bias)      :355-357   fp32 parameters (dit_model.py:621-638)
// i.e. out[row] = bf16( LayerNorm_h(x[row] * gate[map[row]]) * (w + 1) + b + residual[row] ), seven eager kernels and
// ~44 B/element of HBM traffic in the reference, one kernel and 8 B/element (fp32 x) or 6 B/element (bf16 x) here.
//
// HBM-bound. A row of h elements is owned by kWarpsPerRow warps and lives in registers between the three LayerNorm
// passes (product, mean, variance), so x and the residual are read exactly once: lane l of warp w holds the 8-element
// groups v * 32 * W + w * 32 + l (v < 6), every global access is a 16-byte vector and a warp instruction covers whole
// 512-byte (bf16) / 1 KB (fp32, two instructions) segments. gate / w / b are re-read per row from L1 (<= 8 gate rows).
// At most 6 groups per lane (3072 -> two warps per row, 6144 -> four, joined by a named barrier): 126 registers, two
// CTAs per SM. With 12 groups per lane (one warp per 3072-wide row, 253 registers, one CTA per SM) the same kernel
// ran at 4.5 instead of 5.4 TB/s: a warp has nothing in flight while it works on its row (profiles/README.md).
#include "common.cuh"
#include "gating.cuh"

namespace magi {

namespace {

#ifndef MAGI_GATE_MAXVEC
#define MAGI_GATE_MAXVEC 6
#endif
#ifndef MAGI_GATE_BLOCKS
#define MAGI_GATE_BLOCKS (MAGI_GATE_MAXVEC <= 6 ? 2 : 1)
#endif
constexpr int kMaxVec = MAGI_GATE_MAXVEC;   // 8-element groups per lane: h <= kMaxVec * 256 * kWarpsPerRow

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint64_t fmul2g(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t bf2f2(uint32_t u) {
  return pack_f2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2bf2(uint64_t v) {
  float lo, hi;
  unpack_f2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
// 8 consecutive values starting at element `col` of a row stored as fp32 or bf16 -> 4 packed fp32 pairs
template <bool kBf16, bool kStream>
__device__ __forceinline__ void load8(const void* row, int col, uint64_t* v) {
  if constexpr (kBf16) {
    const void* p = reinterpret_cast<const __nv_bfloat16*>(row) + col;
    const uint4 u = kStream ? ldg_stream16(p) : __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = bf2f2(u.x); v[1] = bf2f2(u.y); v[2] = bf2f2(u.z); v[3] = bf2f2(u.w);
  } else {
    const float* p = reinterpret_cast<const float*>(row) + col;
    uint4 a, b;
    if constexpr (kStream) { a = ldg_stream16(p); b = ldg_stream16(p + 4); }
    else { a = __ldg(reinterpret_cast<const uint4*>(p)); b = __ldg(reinterpret_cast<const uint4*>(p + 4)); }
    v[0] = (static_cast<uint64_t>(a.y) << 32) | a.x; v[1] = (static_cast<uint64_t>(a.w) << 32) | a.z;
    v[2] = (static_cast<uint64_t>(b.y) << 32) | b.x; v[3] = (static_cast<uint64_t>(b.w) << 32) | b.z;
  }
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

}  // namespace

template <bool kXBf16, bool kGateBf16, int kWarpsPerRow>
__global__ void __launch_bounds__(256, MAGI_GATE_BLOCKS) gate_ln_residual_kernel(const GatingParams p) {
  constexpr int W = kWarpsPerRow;
  constexpr int kRowsPerCta = 8 / W;
  __shared__ float red[2][kRowsPerCta][W];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = warp / W;          // which of the CTA's concurrent rows
  const int wr = warp % W;            // which column slice of the row
  const int nvec = p.hidden / (256 * W);   // 8-element groups per lane (host guarantees <= kMaxVec, exact division)
  const float inv_h = 1.0f / static_cast<float>(p.hidden);
  const uint64_t one2 = pack_f2(1.f, 1.f);

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kRowsPerCta + slot; ; row += static_cast<int64_t>(gridDim.x) * kRowsPerCta) {
    // all warps of a row group leave together (the named barrier below counts W * 32 threads)
    if (row >= p.rows) break;
    const char* xrow = reinterpret_cast<const char*>(p.x) + row * p.x_row_stride * (kXBf16 ? 2 : 4);
    const __nv_bfloat16* rrow = reinterpret_cast<const __nv_bfloat16*>(p.residual) + row * p.res_row_stride;
    const int64_t gi = p.map_is_i64 ? reinterpret_cast<const int64_t*>(p.cond_map)[row * p.map_stride]
                                    : static_cast<int64_t>(reinterpret_cast<const int32_t*>(p.cond_map)[row * p.map_stride]);
    const char* grow = reinterpret_cast<const char*>(p.gate) + gi * p.gate_row_stride * (kGateBf16 ? 2 : 4);

    // ---- pass 0: t = x * gate (fp32), kept in registers; residual loads issued alongside
    uint64_t t[kMaxVec][4];
    uint4 res[kMaxVec];
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      if (v < nvec) {
        const int col = ((v * W + wr) * 32 + lane) * 8;
        load8<kXBf16, true>(xrow, col, t[v]);
        res[v] = ldg_stream16(rrow + col);
      }
    }
    uint64_t acc_a = 0, acc_b = 0;   // +0.0f pairs
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      if (v < nvec) {
        const int col = ((v * W + wr) * 32 + lane) * 8;
        uint64_t g[4];
        load8<kGateBf16, false>(grow, col, g);
#pragma unroll
        for (int i = 0; i < 4; ++i) t[v][i] = fmul2g(t[v][i], g[i]);
        acc_a = fadd2(acc_a, fadd2(t[v][0], t[v][1]));
        acc_b = fadd2(acc_b, fadd2(t[v][2], t[v][3]));
      }
    }
    float lo, hi;
    unpack_f2(fadd2(acc_a, acc_b), lo, hi);
    float s = warp_sum(lo + hi);
    if constexpr (W > 1) {
      if (lane == 0) red[0][slot][wr] = s;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(W * 32) : "memory");
      s = 0.f;
#pragma unroll
      for (int i = 0; i < W; ++i) s += red[0][slot][i];
    }
    const float mean = s * inv_h;
    // ---- pass 1: biased variance around the mean (two-pass, like ATen's layer_norm)
    const uint64_t mean2 = pack_f2(mean, mean);
    acc_a = 0; acc_b = 0;
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      if (v < nvec) {
#pragma unroll
        for (int i = 0; i < 4; ++i) t[v][i] = fsub2(t[v][i], mean2);
        acc_a = ffma2(t[v][0], t[v][0], acc_a); acc_b = ffma2(t[v][1], t[v][1], acc_b);
        acc_a = ffma2(t[v][2], t[v][2], acc_a); acc_b = ffma2(t[v][3], t[v][3], acc_b);
      }
    }
    unpack_f2(fadd2(acc_a, acc_b), lo, hi);
    float q = warp_sum(lo + hi);
    if constexpr (W > 1) {
      if (lane == 0) red[1][slot][wr] = q;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(W * 32) : "memory");
      q = 0.f;
#pragma unroll
      for (int i = 0; i < W; ++i) q += red[1][slot][i];
    }
    const float rstd = rsqrtf(q * inv_h + p.eps);
    const uint64_t rstd2 = pack_f2(rstd, rstd);
    // ---- pass 2: affine, residual, one rounding to bf16
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.out_row_stride;
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      if (v < nvec) {
        const int col = ((v * W + wr) * 32 + lane) * 8;
        uint64_t w[4], b[4];
        load8<false, false>(p.ln_w, col, w);
        load8<false, false>(p.ln_b, col, b);
        const uint64_t r[4] = {bf2f2(res[v].x), bf2f2(res[v].y), bf2f2(res[v].z), bf2f2(res[v].w)};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (p.zero_centered_gamma != 0) w[i] = fadd2(w[i], one2);
          o[i] = f2bf2(fadd2(ffma2(fmul2g(t[v][i], rstd2), w[i], b[i]), r[i]));
        }
        *reinterpret_cast<uint4*>(orow + col) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

template <bool kXBf16, bool kGateBf16>
static cudaError_t launch_w(const GatingParams& p, int warps_per_row, int grid, cudaStream_t stream) {
  switch (warps_per_row) {
    case 1: gate_ln_residual_kernel<kXBf16, kGateBf16, 1><<<grid, 256, 0, stream>>>(p); break;
    case 2: gate_ln_residual_kernel<kXBf16, kGateBf16, 2><<<grid, 256, 0, stream>>>(p); break;
    case 4: gate_ln_residual_kernel<kXBf16, kGateBf16, 4><<<grid, 256, 0, stream>>>(p); break;
    default: gate_ln_residual_kernel<kXBf16, kGateBf16, 8><<<grid, 256, 0, stream>>>(p); break;
  }
  return cudaGetLastError();
}

int gating_warps_per_row(int hidden) {
  if (hidden <= 0 || hidden % 256 != 0) return 0;
  for (int w = 1; w <= 8; w *= 2)
    if (hidden % (256 * w) == 0 && hidden / (256 * w) <= kMaxVec) return w;
  return 0;
}

cudaError_t gating_launch(const GatingParams& p, bool x_bf16, bool gate_bf16, int sms, cudaStream_t stream) {
  const int w = gating_warps_per_row(p.hidden);
  const int rows_per_cta = 8 / w;
  int64_t blocks = (p.rows + rows_per_cta - 1) / rows_per_cta;
  if (blocks > static_cast<int64_t>(sms) * MAGI_GATE_BLOCKS) blocks = static_cast<int64_t>(sms) * MAGI_GATE_BLOCKS;   // persistent
  if (blocks < 1) blocks = 1;
  const int grid = static_cast<int>(blocks);
  if (x_bf16) return gate_bf16 ? launch_w<true, true>(p, w, grid, stream) : launch_w<true, false>(p, w, grid, stream);
  return gate_bf16 ? launch_w<false, true>(p, w, grid, stream) : launch_w<false, false>(p, w, grid, stream);
}

}  // namespace magi
