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

template <bool kLnBf16>
__device__ __forceinline__ void load_affine(const void* w, const void* b, int sub, bool plus1, float* wf, float* bf) {
  if constexpr (kLnBf16) {
    uint4 uw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(w) + sub * 8);
    uint4 ub = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(b) + sub * 8);
    unpack8(uw, wf);
    unpack8(ub, bf);
    if (plus1) {
      // the reference computes `self.weight + 1` in the parameter dtype: round the sum back to bf16
#pragma unroll
      for (int i = 0; i < 8; ++i) wf[i] = __bfloat162float(__float2bfloat16_rn(wf[i] + 1.0f));
    }
  } else {
    const float4* pw = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(w) + sub * 8);
    const float4* pb = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(b) + sub * 8);
    const float4 w0 = pw[0], w1 = pw[1], b0 = pb[0], b1 = pb[1];
    wf[0] = w0.x; wf[1] = w0.y; wf[2] = w0.z; wf[3] = w0.w; wf[4] = w1.x; wf[5] = w1.y; wf[6] = w1.z; wf[7] = w1.w;
    bf[0] = b0.x; bf[1] = b0.y; bf[2] = b0.z; bf[3] = b0.w; bf[4] = b1.x; bf[5] = b1.y; bf[6] = b1.z; bf[7] = b1.w;
    if (plus1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) wf[i] += 1.0f;
    }
  }
}

}  // namespace

// One warp owns one row (token): it streams that row's q heads, then k heads, then v heads, two heads per step (one
// per half-warp, 16 lanes x 16 B = one 256-byte vector) and kUnroll steps in flight, so every lane keeps kUnroll
// independent 16-byte loads outstanding. sin/cos of the row and the LayerNorm affine of the current tensor live in
// registers for the whole row.
template <bool kLnBf16>
__global__ void __launch_bounds__(256, 3) prologue_kernel(const PrologueParams p) {
  constexpr int kUnroll = 4;
  const int lane = threadIdx.x & 31;
  const int sub = lane & 15;   // 16-byte slot inside the 128-element vector
  const int half = lane >> 4;  // which of the two heads of a step
  const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t warp_id = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int cp = p.head_perm_cp > 1 ? p.head_perm_cp : 1;
  const bool rope = p.sin != nullptr;
  const int partner = sub < 6 ? lane + 6 : (sub < 12 ? lane - 6 : lane);
  const int dsub = sub < 6 ? sub : sub - 6;

  for (int64_t r = warp_id; r < p.rows; r += warps_total) {
    float cs[8], sn[8];
    if (rope && sub < 12) {
      const int64_t trow = r / p.batch;
      const float4* pc = reinterpret_cast<const float4*>(p.cos + trow * 48 + dsub * 8);
      const float4* ps = reinterpret_cast<const float4*>(p.sin + trow * 48 + dsub * 8);
      const float4 c0 = pc[0], c1 = pc[1], s0 = ps[0], s1 = ps[1];
      cs[0] = c0.x; cs[1] = c0.y; cs[2] = c0.z; cs[3] = c0.w; cs[4] = c1.x; cs[5] = c1.y; cs[6] = c1.z; cs[7] = c1.w;
      sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
      if (sub < 6) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sn[i] = -sn[i];   // first half: x1*cos - x2*sin ; second half: x2*cos + x1*sin
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { cs[i] = 1.f; sn[i] = 0.f; }
    }

#pragma unroll 1
    for (int type = 0; type < 3; ++type) {  // 0 = q, 1 = k, 2 = v
      const void* xbase = type == 0 ? p.x_q : type == 1 ? p.x_k : p.x_v;
      if (xbase == nullptr) continue;
      const int H = type == 0 ? p.Hq : p.Hkv;
      const int64_t rstride = type == 0 ? p.xq_row_stride : type == 1 ? p.xk_row_stride : p.xv_row_stride;
      const int64_t hstride = type == 0 ? p.xq_head_stride : type == 1 ? p.xk_head_stride : p.xv_head_stride;
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(xbase) + r * rstride + sub * 8;
      float wf[8], bf[8];
      if (type != 2)
        load_affine<kLnBf16>(type == 0 ? p.ln_q_w : p.ln_k_w, type == 0 ? p.ln_q_b : p.ln_k_b, sub,
                             p.zero_centered_gamma != 0, wf, bf);

      for (int h0 = 0; h0 < H; h0 += 2 * kUnroll) {
        uint4 raw[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int h = h0 + 2 * u + half;
          raw[u] = make_uint4(0, 0, 0, 0);
          if (h < H) raw[u] = ld_stream(src + h * hstride);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int h = h0 + 2 * u + half;
          if (h0 + 2 * u >= H) break;   // warp-uniform: both halves are past the last head
          const bool valid = h < H;
          uint4 packed = raw[u];
          if (type != 2) {
            float v[8];
            unpack8(raw[u], v);
            // ---- LayerNorm over 128 in fp32
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[i];
            const float mean = half_warp_sum(s) * (1.0f / 128.0f);
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] -= mean; sq += v[i] * v[i]; }
            const float var = half_warp_sum(sq) * (1.0f / 128.0f);
            const float rstd = 1.0f / sqrtf(var + p.eps);
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = v[i] * rstd * wf[i] + bf[i];
            // ---- NeoX RoPE on dims [0,48) x [48,96): slots 0-5 pair with slots 6-11 of the same half-warp
            if (rope) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float o = __shfl_sync(0xffffffffu, y[i], partner);
                y[i] = y[i] * cs[i] + o * sn[i];
              }
            }
            packed = pack8(y);
          }
          if (!valid) continue;

          // ---- stores: attention (or all-to-all send) layout, plus the fused KV-cache append
          if (type == 0) {
            const int hn = p.Hq / cp;
            const int64_t drow = static_cast<int64_t>(h / hn) * p.rows + r;
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.q_out) + drow * p.q_out_token_stride +
                                 static_cast<int64_t>(h % hn) * p.q_out_head_stride + sub * 8;
            *reinterpret_cast<uint4*>(dst) = packed;
          } else {
            const int col = (type == 2 ? 128 : 0) + sub * 8;
            if (cp > p.Hkv) {
              // fewer kv heads than ranks: replicate (repeat_interleave, context_parallel.py:393-394)
              const int rep = cp / p.Hkv;
              for (int c = h * rep; c < (h + 1) * rep; ++c) {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.kv_out) +
                                     (static_cast<int64_t>(c) * p.rows + r) * p.kv_out_token_stride + col;
                *reinterpret_cast<uint4*>(dst) = packed;
              }
            } else {
              const int hn = p.Hkv / cp;
              const int64_t drow = static_cast<int64_t>(h / hn) * p.rows + r;
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.kv_out) + drow * p.kv_out_token_stride +
                                   static_cast<int64_t>(h % hn) * p.kv_out_head_stride + col;
              *reinterpret_cast<uint4*>(dst) = packed;
            }
            if (p.cache != nullptr && r < p.cache_rows) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.cache) +
                                   (p.cache_row_offset + r) * p.cache_token_stride +
                                   static_cast<int64_t>(h) * p.cache_head_stride + col;
              *reinterpret_cast<uint4*>(dst) = packed;
            }
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
