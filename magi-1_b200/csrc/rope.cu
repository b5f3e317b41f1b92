// Stand-alone non-interleaved (NeoX) rotary embedding, out of place or in place:
//   out[..., :r/2]  = x1 * cos - x2 * sin      x1 = x[..., :r/2], x2 = x[..., r/2:r]
//   out[..., r/2:r] = x1 * sin + x2 * cos      dims [r, d) are copied
// in fp32 whatever the storage type, one rounding at the end.
//
// Reference call site being replaced (samiazirar/MAGI-1): flash_apply_rotary_emb(q, cos_emb, sin_emb) in get_q / get_k,
// inference/model/dit/dit_module.py:961, 978 (flash_attn.layers.rotary.apply_rotary_emb, a Triton kernel), on the fp32
// (b, sq, hn, hd) TRANSPOSED VIEW of the LayerNorm output, hence the explicit batch / sequence / head strides.
// The product's own forward never launches this kernel: its prologue (prologue.cu) fuses the rotation with the
// LayerNorm, the packing and the KV-cache append. HBM-bound: 4 elements per thread, 16-byte (fp32) or 8-byte (bf16)
// accesses, the two partners of a rotation in the same thread.
#include <cuda_bf16.h>

#include "rope.cuh"

namespace magi {

namespace {

template <bool kBf16>
__device__ __forceinline__ float4 load4(const void* base, int64_t off) {
  if constexpr (kBf16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
  }
}

template <bool kBf16>
__device__ __forceinline__ void store4(void* base, int64_t off, const float4& v) {
  if constexpr (kBf16) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = u;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v;
  }
}

}  // namespace

template <bool kBf16>
__global__ void __launch_bounds__(256) rotary_neox_kernel(const RotaryParams p) {
  const int half = p.rotary_dim / 2;
  const int rot_items = half / 4;                               // each covers 4 (x1, x2) pairs
  const int items = rot_items + (p.head_dim - p.rotary_dim) / 4;   // + pass-through groups of 4
  const int64_t total = p.batch * p.seqlen * p.heads * items;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int it = static_cast<int>(i % items);
    int64_t t = i / items;
    const int h = static_cast<int>(t % p.heads);
    t /= p.heads;
    const int64_t s = t % p.seqlen;
    const int64_t b = t / p.seqlen;
    const int64_t xo = b * p.x_bs + s * p.x_ss + static_cast<int64_t>(h) * p.x_hs;
    const int64_t oo = b * p.o_bs + s * p.o_ss + static_cast<int64_t>(h) * p.o_hs;
    if (it < rot_items) {
      const int c = 4 * it;
      const float4 x1 = load4<kBf16>(p.x, xo + c), x2 = load4<kBf16>(p.x, xo + half + c);
      const float4 cs = __ldg(reinterpret_cast<const float4*>(p.cos + s * half + c));
      const float4 sn = __ldg(reinterpret_cast<const float4*>(p.sin + s * half + c));
      float4 o1, o2;
      o1.x = x1.x * cs.x - x2.x * sn.x; o2.x = x1.x * sn.x + x2.x * cs.x;
      o1.y = x1.y * cs.y - x2.y * sn.y; o2.y = x1.y * sn.y + x2.y * cs.y;
      o1.z = x1.z * cs.z - x2.z * sn.z; o2.z = x1.z * sn.z + x2.z * cs.z;
      o1.w = x1.w * cs.w - x2.w * sn.w; o2.w = x1.w * sn.w + x2.w * cs.w;
      store4<kBf16>(p.out, oo + c, o1);
      store4<kBf16>(p.out, oo + half + c, o2);
    } else {
      const int c = p.rotary_dim + 4 * (it - rot_items);
      store4<kBf16>(p.out, oo + c, load4<kBf16>(p.x, xo + c));
    }
  }
}

cudaError_t rotary_launch(const RotaryParams& p, bool bf16, int grid, cudaStream_t stream) {
  if (bf16) rotary_neox_kernel<true><<<grid, 256, 0, stream>>>(p);
  else rotary_neox_kernel<false><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace magi
