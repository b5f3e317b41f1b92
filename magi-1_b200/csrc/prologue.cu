// Fused QK-LayerNorm(+1p, bias) + NeoX RoPE-96 + layout prologue, and the small memory-bound helpers around the
// attention kernel (KV-cache append, LSE merge for ring-KV, Ulysses head<->sequence re-layout).
//
// Reference semantics being replaced (samiazirar/MAGI-1):
//   get_q / get_k / get_v / get_kv          inference/model/dit/dit_module.py:953-996
//   FusedLayerNorm.forward (w+1, bias)      dit_module.py:355-357
//   get_xqkv's two LayerNorms (no RoPE)     dit_module.py:1005-1021
//   KV-cache slice-assign                   dit_module.py:918-929
//   all_to_all_input_split's re-layout      inference/infra/parallelism/context_parallel.py:393-396
//
// HBM-bound: every element is read once (2 B) and written once (2 B); one half-warp owns one (row, head) vector of
// 128 bf16 as 16 x 16-byte lanes, so every global access is a full 256-byte coalesced segment.
#include "common.cuh"
#include "prologue.cuh"

namespace magi {

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  return u;
}
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float half_warp_sum(float x) {
  x += __shfl_xor_sync(0xffffffffu, x, 8);
  x += __shfl_xor_sync(0xffffffffu, x, 4);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  return x;
}

}  // namespace

#ifndef MAGI_PRO_UNROLL
#define MAGI_PRO_UNROLL 4
#endif
#ifndef MAGI_PRO_BLOCKS
#define MAGI_PRO_BLOCKS 3
#endif

__device__ __forceinline__ uint2 ld_8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void unpack4(const uint2& u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <bool kBf16>
__device__ __forceinline__ void load4(const void* base, int off, float* v) {
  if constexpr (kBf16) {
    unpack4(*reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off), v);
  } else {
    const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
}

// One warp owns one row (token): it streams that row's q heads, then k heads, then v heads, two heads per step (one
// per half-warp) and MAGI_PRO_UNROLL steps in flight. Inside a half-warp, lane s < 12 holds elements [4s, 4s+4) and
// [48+4s, 48+4s+4) of the 128-element head vector - the two RoPE partners of the non-interleaved rotation sit in the
// same lane, so the rotation needs no shuffles - and lanes 12..15 hold the pass-through dims [96, 128).
// sin/cos of the row and the LayerNorm affine of the current tensor stay in registers for the whole row.
template <bool kLnBf16>
__global__ void __launch_bounds__(256, MAGI_PRO_BLOCKS) prologue_kernel(const PrologueParams p) {
  constexpr int kUnroll = MAGI_PRO_UNROLL;
  const int lane = threadIdx.x & 31;
  const int sub = lane & 15;
  const int half = lane >> 4;  // which of the two heads of a step
  const int off0 = sub < 12 ? 4 * sub : 96 + 8 * (sub - 12);
  const int off1 = sub < 12 ? 48 + 4 * sub : off0 + 4;
  const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t warp_id = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int cp = p.head_perm_cp > 1 ? p.head_perm_cp : 1;
  const bool rope = p.sin != nullptr && sub < 12;

  for (int64_t r = warp_id; r < p.rows; r += warps_total) {
    float cs[4], sn[4];
    if (rope) {
      const int64_t trow = r / p.batch;
      const float4 c = *reinterpret_cast<const float4*>(p.cos + trow * 48 + 4 * sub);
      const float4 s4 = *reinterpret_cast<const float4*>(p.sin + trow * 48 + 4 * sub);
      cs[0] = c.x; cs[1] = c.y; cs[2] = c.z; cs[3] = c.w;
      sn[0] = s4.x; sn[1] = s4.y; sn[2] = s4.z; sn[3] = s4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { cs[i] = 1.f; sn[i] = 0.f; }
    }

#pragma unroll 1
    for (int type = 0; type < 3; ++type) {  // 0 = q, 1 = k, 2 = v
      const void* xbase = type == 0 ? p.x_q : type == 1 ? p.x_k : p.x_v;
      if (xbase == nullptr) continue;
      const int H = type == 0 ? p.Hq : p.Hkv;
      const int64_t rstride = type == 0 ? p.xq_row_stride : type == 1 ? p.xk_row_stride : p.xv_row_stride;
      const int64_t hstride = type == 0 ? p.xq_head_stride : type == 1 ? p.xk_head_stride : p.xv_head_stride;
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(xbase) + r * rstride;
      float wf[8], bf[8];
      if (type != 2) {
        const void* w = type == 0 ? p.ln_q_w : p.ln_k_w;
        const void* b = type == 0 ? p.ln_q_b : p.ln_k_b;
        load4<kLnBf16>(w, off0, wf); load4<kLnBf16>(w, off1, wf + 4);
        load4<kLnBf16>(b, off0, bf); load4<kLnBf16>(b, off1, bf + 4);
        if (p.zero_centered_gamma != 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // bf16 parameters: the reference computes `self.weight + 1` in bf16, i.e. rounds the sum
            wf[i] = kLnBf16 ? __bfloat162float(__float2bfloat16_rn(wf[i] + 1.0f)) : wf[i] + 1.0f;
          }
        }
      }

      // destination bases of this (row, tensor): q_out, or the K / V half of the packed kv row (and of the cache row)
      __nv_bfloat16* out_base = type == 0 ? reinterpret_cast<__nv_bfloat16*>(p.q_out)
                                          : reinterpret_cast<__nv_bfloat16*>(p.kv_out) + (type == 2 ? 128 : 0);
      const int64_t out_tstride = type == 0 ? p.q_out_token_stride : p.kv_out_token_stride;
      const int64_t out_hstride = type == 0 ? p.q_out_head_stride : p.kv_out_head_stride;
      __nv_bfloat16* out_row = out_base + r * out_tstride;
      __nv_bfloat16* cache_row =
          (type != 0 && p.cache != nullptr && r < p.cache_rows)
              ? reinterpret_cast<__nv_bfloat16*>(p.cache) + (p.cache_row_offset + r) * p.cache_token_stride +
                    (type == 2 ? 128 : 0)
              : nullptr;

      for (int h0 = 0; h0 < H; h0 += 2 * kUnroll) {
        uint2 ra[kUnroll], rb[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int h = h0 + 2 * u + half;
          ra[u] = make_uint2(0, 0);
          rb[u] = make_uint2(0, 0);
          if (h < H) {
            ra[u] = ld_8(src + h * hstride + off0);
            rb[u] = ld_8(src + h * hstride + off1);
          }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int h = h0 + 2 * u + half;
          if (h0 + 2 * u >= H) break;   // warp-uniform: both halves are past the last head
          const bool valid = h < H;
          uint2 pa = ra[u], pb = rb[u];
          if (type != 2) {
            float v[8];
            unpack4(ra[u], v);
            unpack4(rb[u], v + 4);
            // ---- LayerNorm over 128 in fp32 (two-pass, biased variance)
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[i];
            const float mean = half_warp_sum(s) * (1.0f / 128.0f);
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] -= mean; sq += v[i] * v[i]; }
            const float var = half_warp_sum(sq) * (1.0f / 128.0f);
            const float rstd = rsqrtf(var + p.eps);   // MUFU.RSQ (2 ulp in fp32: invisible after the bf16 rounding)
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = v[i] * rstd * wf[i] + bf[i];
            // ---- NeoX RoPE: (x1, x2) -> (x1 cos - x2 sin, x2 cos + x1 sin); identity on lanes 12..15 (cos 1, sin 0)
            if (rope) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float x1 = y[i], x2 = y[4 + i];
                y[i] = x1 * cs[i] - x2 * sn[i];
                y[4 + i] = x2 * cs[i] + x1 * sn[i];
              }
            }
            pa = make_uint2(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]));
            pb = make_uint2(pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
          }
          if (!valid) continue;

          // ---- stores: attention (or all-to-all send) layout, plus the fused KV-cache append.
          // Row bases were hoisted out of the head loop; without head permutation (cp == 1) a head costs one
          // multiply-add per target - no integer divisions on the common path.
          auto store = [&](__nv_bfloat16* dst) {
            *reinterpret_cast<uint2*>(dst + off0) = pa;
            *reinterpret_cast<uint2*>(dst + off1) = pb;
          };
          if (cp == 1) {
            store(out_row + h * out_hstride);
            if (cache_row != nullptr) store(cache_row + h * p.cache_head_stride);
          } else if (type == 0) {
            const int hn = p.Hq / cp;
            store(out_base + (static_cast<int64_t>(h / hn) * p.rows + r) * out_tstride +
                  static_cast<int64_t>(h % hn) * out_hstride);
          } else if (cp > p.Hkv) {
            // fewer kv heads than ranks: replicate (repeat_interleave, context_parallel.py:393-394)
            const int rep = cp / p.Hkv;
            for (int c = h * rep; c < (h + 1) * rep; ++c)
              store(out_base + (static_cast<int64_t>(c) * p.rows + r) * out_tstride);
          } else {
            const int hn = p.Hkv / cp;
            store(out_base + (static_cast<int64_t>(h / hn) * p.rows + r) * out_tstride +
                  static_cast<int64_t>(h % hn) * out_hstride);
          }
        }
      }
    }
  }
}

// cache[row_offset + r, :] = kv[r, :]   (row_elems bf16 per row, multiple of 8)
__global__ void __launch_bounds__(256) kv_append_kernel(__nv_bfloat16* __restrict__ cache, int64_t cache_stride,
                                                        const __nv_bfloat16* __restrict__ kv, int64_t kv_stride,
                                                        int64_t row_offset, int64_t rows, int row_vecs) {
  const int64_t total = rows * row_vecs;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / row_vecs;
    const int c = static_cast<int>(i - r * row_vecs);
    const uint4 v = ld_stream(kv + r * kv_stride + c * 8);
    *reinterpret_cast<uint4*>(cache + (row_offset + r) * cache_stride + c * 8) = v;
  }
}

// out_a = out_a * exp(lse_a - lse) + out_b * exp(lse_b - lse), lse_a = lse = logaddexp(lse_a, lse_b)
// one half-warp per (row, head); lse tensors are (rows, heads) fp32.
__global__ void __launch_bounds__(256) lse_merge_kernel(__nv_bfloat16* __restrict__ out_a, float* __restrict__ lse_a,
                                                        const __nv_bfloat16* __restrict__ out_b,
                                                        const float* __restrict__ lse_b, int64_t rows, int heads,
                                                        int64_t a_ts, int64_t a_hs, int64_t b_ts, int64_t b_hs) {
  const int sub = threadIdx.x & 15;
  const int64_t units = rows * heads;
  const int64_t hw_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 4;
  for (int64_t u = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 4; u < units; u += hw_total) {
    const int64_t r = u / heads;
    const int h = static_cast<int>(u - r * heads);
    const float la = lse_a[u], lb = lse_b[u];
    const float mx = fmaxf(la, lb);
    float wa, wb, lse;
    if (mx == -INFINITY) {
      wa = 0.f; wb = 0.f; lse = -INFINITY;
    } else {
      const float ea = __expf(la - mx), eb = __expf(lb - mx);
      const float inv = 1.0f / (ea + eb);
      wa = ea * inv; wb = eb * inv;
      lse = mx + __logf(ea + eb);
    }
    __nv_bfloat16* pa = out_a + r * a_ts + static_cast<int64_t>(h) * a_hs + sub * 8;
    const __nv_bfloat16* pb = out_b + r * b_ts + static_cast<int64_t>(h) * b_hs + sub * 8;
    float fa[8], fb[8];
    unpack8(*reinterpret_cast<const uint4*>(pa), fa);
    unpack8(ld_stream(pb), fb);
#pragma unroll
    for (int i = 0; i < 8; ++i) fa[i] = fa[i] * wa + fb[i] * wb;
    *reinterpret_cast<uint4*>(pa) = pack8(fa);
    if (sub == 0) lse_a[u] = lse;
  }
}

// mode 0: dst[(c, s), hl, :] = src[s, (c*hn + hl) / rep, :]   "seq (cp hn) hd -> (cp seq) hn hd" (+ head replication)
// mode 1: dst[s, c*hn + hl, :] = src[(c, s), hl, :]           "(cp seq) hn hd -> seq (cp hn) hd"
__global__ void __launch_bounds__(256) ulysses_relayout_kernel(__nv_bfloat16* __restrict__ dst,
                                                               const __nv_bfloat16* __restrict__ src, int64_t seq,
                                                               int hn, int head_vecs, int cp, int rep, int mode,
                                                               int dst_hn, int dst_off) {
  const int64_t total = seq * cp * hn * head_vecs;
  const int src_heads = cp * hn / rep;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % head_vecs);
    int64_t t = i / head_vecs;
    if (mode == 0) {  // iterate over dst: (c, s, hl)
      const int hl = static_cast<int>(t % hn); t /= hn;
      const int64_t s = t % seq;
      const int c = static_cast<int>(t / seq);
      const int hs = (c * hn + hl) / rep;
      const uint4 v = ld_stream(src + ((s * src_heads + hs) * head_vecs + e) * 8);
      *reinterpret_cast<uint4*>(dst + i * 8) = v;
    } else {          // iterate over dst: (s, c, hl)
      const int hl = static_cast<int>(t % hn); t /= hn;
      const int c = static_cast<int>(t % cp);
      const int64_t s = t / cp;
      const uint4 v = ld_stream(src + (((static_cast<int64_t>(c) * seq + s) * hn + hl) * head_vecs + e) * 8);
      *reinterpret_cast<uint4*>(dst + (((s * cp + c) * dst_hn + dst_off + hl) * head_vecs + e) * 8) = v;
    }
  }
}

// CSO receive side: dst[d * clip + t, :] = src[((t / spb) * dn + d) * spb + t % spb, :] for t < clip, i.e.
// "(cp dn spb) -> dn (cp spb)" followed by dropping every chunk's zero padding (dit_module.py:1215-1224).
__global__ void __launch_bounds__(256) cso_unshuffle_kernel(__nv_bfloat16* __restrict__ dst, int64_t dst_stride,
                                                            const __nv_bfloat16* __restrict__ src, int64_t src_stride,
                                                            int dn, int64_t spb, int64_t clip, int row_vecs) {
  const int64_t total = static_cast<int64_t>(dn) * clip * row_vecs;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / row_vecs;
    const int c8 = static_cast<int>(i - r * row_vecs);
    const int64_t d = r / clip, t = r - d * clip;
    const int64_t c = t / spb, s = t - c * spb;
    const uint4 v = ld_stream(src + ((c * dn + d) * spb + s) * src_stride + c8 * 8);
    *reinterpret_cast<uint4*>(dst + r * dst_stride + c8 * 8) = v;
  }
}

cudaError_t cso_unshuffle_launch(void* dst, int64_t dst_stride, const void* src, int64_t src_stride, int dn,
                                 int64_t spb, int64_t clip, int row_vecs, int grid, cudaStream_t stream) {
  cso_unshuffle_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(dst), dst_stride,
                                                 reinterpret_cast<const __nv_bfloat16*>(src), src_stride, dn, spb,
                                                 clip, row_vecs);
  return cudaGetLastError();
}

cudaError_t prologue_launch(const PrologueParams& p, bool ln_bf16, int grid, cudaStream_t stream) {
  if (ln_bf16) prologue_kernel<true><<<grid, 256, 0, stream>>>(p);
  else prologue_kernel<false><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}
cudaError_t kv_append_launch(void* cache, int64_t cache_stride, const void* kv, int64_t kv_stride, int64_t row_offset,
                             int64_t rows, int row_vecs, int grid, cudaStream_t stream) {
  kv_append_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(cache), cache_stride,
                                             reinterpret_cast<const __nv_bfloat16*>(kv), kv_stride, row_offset, rows,
                                             row_vecs);
  return cudaGetLastError();
}
cudaError_t lse_merge_launch(void* out_a, float* lse_a, const void* out_b, const float* lse_b, int64_t rows, int heads,
                             int64_t a_ts, int64_t a_hs, int64_t b_ts, int64_t b_hs, int grid, cudaStream_t stream) {
  lse_merge_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(out_a), lse_a,
                                             reinterpret_cast<const __nv_bfloat16*>(out_b), lse_b, rows, heads, a_ts,
                                             a_hs, b_ts, b_hs);
  return cudaGetLastError();
}
cudaError_t ulysses_relayout_launch(void* dst, const void* src, int64_t seq, int hn, int head_vecs, int cp, int rep,
                                    int mode, int dst_hn, int dst_off, int grid, cudaStream_t stream) {
  ulysses_relayout_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(dst),
                                                    reinterpret_cast<const __nv_bfloat16*>(src), seq, hn, head_vecs,
                                                    cp, rep, mode, dst_hn, dst_off);
  return cudaGetLastError();
}

}  // namespace magi
