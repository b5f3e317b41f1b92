// MLP activation between fc1 and fc2 (CustomMLP.forward, inference/model/dit/dit_module.py:541-549), HBM-bound:
//   gated (SwiGLU, 24B): out[r, c] = bf16( silu(float(x[r, c])) * float(x[r, f + c]) )     x: (rows, 2 f)
//                        = flashinfer.activation.silu_and_mul (dit_module.py:546): ONE rounding, where the eager
//                        chunk + silu + mul chain rounds twice and moves 10 bytes per output element instead of 6
//   plain (GELU, 4.5B):  out[r, c] = bf16( gelu_erf(float(x[r, c])) )                       x: (rows, f)
// 16-byte vectors, grid-stride over (row, vector) pairs, grid sized from the SM count.
#include "activation.cuh"
#include "common.cuh"

namespace magi {

namespace {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ uint4 ld_nc(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

template <bool kGated>
__global__ void __launch_bounds__(256) mlp_act_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_row_stride,
                                                      __nv_bfloat16* __restrict__ out, int64_t out_row_stride,
                                                      int64_t rows, int f) {
  const int vecs = f >> 3;
  const int64_t total = rows * vecs;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vecs;
    const int c = static_cast<int>(i - r * vecs) << 3;
    const __nv_bfloat16* xr = x + r * x_row_stride + c;
    float a[8], y[8];
    unpack8(ld_nc(xr), a);
    if (kGated) {
      float b[8];
      unpack8(ld_nc(xr + f), b);
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = silu_f(a[k]) * b[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = gelu_f(a[k]);
    }
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
    *reinterpret_cast<uint4*>(out + r * out_row_stride + c) = o;
  }
}

}  // namespace

cudaError_t mlp_act_launch(const void* x, int64_t x_row_stride, void* out, int64_t out_row_stride, int64_t rows, int f,
                           bool gated, int sms, cudaStream_t stream) {
  const int64_t total = rows * (f >> 3);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sms) * 8;   // 8 resident 256-thread CTAs per SM; grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out);
  if (gated) mlp_act_kernel<true><<<static_cast<int>(blocks), 256, 0, stream>>>(xp, x_row_stride, op, out_row_stride, rows, f);
  else mlp_act_kernel<false><<<static_cast<int>(blocks), 256, 0, stream>>>(xp, x_row_stride, op, out_row_stride, rows, f);
  return cudaGetLastError();
}

}  // namespace magi
