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

template <bool kLnBf16>
__global__ void __launch_bounds__(256) prologue_kernel(const PrologueParams p) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & 15;   // 16-byte slot inside the 128-element vector
  const int half = lane >> 4;  // which of the warp's two units
  const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t warp_id = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int nq = p.x_q ? p.Hq : 0, nk = p.x_k ? p.Hkv : 0, nv = p.x_v ? p.Hkv : 0;
  const int per_row = nq + nk + nv;
  const int64_t units = p.rows * per_row;
  const int cp = p.head_perm_cp > 1 ? p.head_perm_cp : 1;

  for (int64_t base = warp_id * 2; base < units; base += warps_total * 2) {
    const int64_t u = base + half;
    const bool valid = u < units;
    const int64_t r = valid ? u / per_row : 0;
    const int hh = valid ? static_cast<int>(u - r * per_row) : 0;
    int type, h;  // 0 = q, 1 = k, 2 = v
    if (hh < nq) { type = 0; h = hh; }
    else if (hh < nq + nk) { type = 1; h = hh - nq; }
    else { type = 2; h = hh - nq - nk; }

    const __nv_bfloat16* src =
        type == 0 ? reinterpret_cast<const __nv_bfloat16*>(p.x_q) + r * p.xq_row_stride
                  : type == 1 ? reinterpret_cast<const __nv_bfloat16*>(p.x_k) + r * p.xk_row_stride
                              : reinterpret_cast<const __nv_bfloat16*>(p.x_v) + r * p.xv_row_stride;
    const int64_t hstride = type == 0 ? p.xq_head_stride : type == 1 ? p.xk_head_stride : p.xv_head_stride;
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (valid) raw = ld_stream(src + h * hstride + sub * 8);
    float v[8];
    unpack8(raw, v);

    // ---- LayerNorm over 128 in fp32 (all lanes take part in the shuffles; V lanes discard the result)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mean = half_warp_sum(s) * (1.0f / 128.0f);
    float d[8], sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = v[i] - mean; sq += d[i] * d[i]; }
    const float var = half_warp_sum(sq) * (1.0f / 128.0f);
    const float rstd = 1.0f / sqrtf(var + p.eps);
    float y[8];
    if (type != 2) {
      float wf[8], bf[8];
      load_affine<kLnBf16>(type == 0 ? p.ln_q_w : p.ln_k_w, type == 0 ? p.ln_q_b : p.ln_k_b, sub,
                           p.zero_centered_gamma != 0, wf, bf);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = d[i] * rstd * wf[i] + bf[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = v[i];
    }

    // ---- NeoX RoPE on dims [0,48) x [48,96): slots 0-5 pair with slots 6-11 (same half-warp)
    if (p.sin != nullptr) {
      const int partner = sub < 6 ? lane + 6 : (sub < 12 ? lane - 6 : lane);
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = __shfl_sync(0xffffffffu, y[i], partner);
      if (type != 2 && sub < 12) {
        const int dsub = sub < 6 ? sub : sub - 6;
        const int64_t trow = r / p.batch;
        const float4* pc = reinterpret_cast<const float4*>(p.cos + trow * 48 + dsub * 8);
        const float4* ps = reinterpret_cast<const float4*>(p.sin + trow * 48 + dsub * 8);
        const float4 c0 = pc[0], c1 = pc[1], s0 = ps[0], s1 = ps[1];
        const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        if (sub < 6) {
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = y[i] * cs[i] - o[i] * sn[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = y[i] * cs[i] + o[i] * sn[i];
        }
      }
    }
    if (!valid) continue;
    const uint4 packed = type == 2 ? raw : pack8(y);

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
