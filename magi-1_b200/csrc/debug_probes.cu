// Bring-up probes, compiled only into debug variants (-DMAGI_DEBUG_PROBES=1, tools/build_variants.sh); the product
// library exports none of this.
#if MAGI_DEBUG_PROBES
#include "common.cuh"
#include "tmem_shapes.cuh"

namespace magi {

// Checks the register <-> (lane, column) maps of the 16-lane TMEM shapes that attn_pair.cu relies on, against the
// row-per-thread 32x32b shape attn_fwd.cu has been validated with. err[0] = mismatches of the 16x256b.x16 load
// (at lane offsets 0 and 16 of each warp's quarter), err[1] = of the 16x128b.x16 store, err[2] = of the 16x256b.x4
// load / store round trip used by the O rescale.
__global__ void __launch_bounds__(256, 1) tmem_layout_probe_kernel(int* err) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const int quarter = warp & 3, hh = warp >> 2;
  const uint32_t qlane = static_cast<uint32_t>(quarter * 32) << 16;
  const uint32_t flane = static_cast<uint32_t>(quarter * 32 + hh * 16) << 16;
  // (1) pattern through the known-good shape: value(row, col) = row * 1024 + col, 128 columns
  if (hh == 0) {
    uint32_t r[32];
    for (int c0 = 0; c0 < 128; c0 += 32) {
#pragma unroll
      for (int k = 0; k < 32; ++k) r[k] = static_cast<uint32_t>((quarter * 32 + lane) * 1024 + c0 + k);
      MAGI_TMEM_ST32(tb + qlane + c0, r, 0);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  int bad = 0;
  {
    uint32_t s[64];
    MAGI_TMEM_LD_16x256b_X16(tb + flane, s, 0);
    tmem_wait_ld();
#pragma unroll
    for (int n = 0; n < 16; ++n)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const uint32_t want = static_cast<uint32_t>((quarter * 32 + hh * 16 + (lane >> 2) + 8 * h) * 1024 + 8 * n +
                                                      2 * (lane & 3) + e);
          bad += s[4 * n + 2 * h + e] != want;
        }
  }
  if (bad) atomicAdd(&err[0], bad);
  // (2) 16x128b.x16 store of "packed" words into columns [128, 192): word(row, col) = row * 1024 + col, read back
  // through 32x32b
  {
    uint32_t pk[32];
#pragma unroll
    for (int n = 0; n < 16; ++n)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        pk[2 * n + h] = static_cast<uint32_t>((quarter * 32 + hh * 16 + (lane >> 2) + 8 * h) * 1024 + 4 * n + (lane & 3));
    MAGI_TMEM_ST_16x128b_X16(tb + flane + 128, pk, 0);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  bad = 0;
  if (hh == 0) {
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
      MAGI_TMEM_LD32(tb + qlane + 128 + c0, r, 0);
      tmem_wait_ld();
#pragma unroll
      for (int k = 0; k < 32; ++k) bad += r[k] != static_cast<uint32_t>((quarter * 32 + lane) * 1024 + c0 + k);
    }
  }
  if (bad) atomicAdd(&err[1], bad);
  // (3) 16x256b.x4 round trip on columns [32, 64): +1 to every element, check through 32x32b
  {
    uint32_t o[16];
    MAGI_TMEM_LD_16x256b_X4(tb + flane + 32, o, 0);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] += 1;
    MAGI_TMEM_ST_16x256b_X4(tb + flane + 32, o, 0);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  bad = 0;
  if (hh == 0) {
    uint32_t r[32];
    MAGI_TMEM_LD32(tb + qlane + 32, r, 0);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < 32; ++k) bad += r[k] != static_cast<uint32_t>((quarter * 32 + lane) * 1024 + 32 + k + 1);
  }
  if (bad) atomicAdd(&err[2], bad);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

}  // namespace magi

extern "C" __attribute__((visibility("default"))) int magi_debug_tmem_layout(int* host_err3) {
  int* d = nullptr;
  if (cudaMalloc(&d, 3 * sizeof(int)) != cudaSuccess) return 1;
  cudaMemset(d, 0, 3 * sizeof(int));
  magi::tmem_layout_probe_kernel<<<1, 256>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(host_err3, d, 3 * sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : 2;
}
#endif
