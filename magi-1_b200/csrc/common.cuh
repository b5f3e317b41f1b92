// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Everything here is written against the PTX ISA directly; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#ifndef MAGI_WATCHDOG_NS
#define MAGI_WATCHDOG_NS 4000000000ull  // an mbarrier wait longer than 4 s traps instead of hanging the GPU
#endif

namespace magi {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One lane of a fully converged warp (the same lane every time for the same mask). Single-lane tcgen05 / TMA issue
// must sit behind this (not behind `lane == 0` around a whole loop): operands computed in convergent code stay in
// uniform registers, otherwise ptxas wraps every UTCHMMA / UTMALDG in a per-lane "waterfall" loop of R2UR moves.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---------------------------------------------------------------- programmatic dependent launch (griddepcontrol)
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become resident while the kernel
// before it in the stream is still running: as soon as every CTA of that kernel has executed launch_dependents (or
// exited). It must execute pdl_wait() before it touches anything the earlier kernel reads or writes; the wait returns
// when the earlier grid has completed and its memory is visible. Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch with (pdl = true) or without the programmatic-serialization attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream,
                                 bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// MAGI_PDL=0 turns the attribute off everywhere (plain stream-ordered launches).
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MAGI_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive that hands back the barrier's state token (only used to give the arrive a consumer, MAGI_EARLY_PUBLISH)
__device__ __forceinline__ uint64_t mbar_arrive_token(uint64_t* bar) {
  uint64_t state;
  asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(smem_u32(bar)) : "memory");
  return state;
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Blocking wait with a watchdog: a protocol bug traps (launch failure) instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0xfffu) == 0) {
      uint64_t t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > MAGI_WATCHDOG_NS) {
        printf("[magi] mbarrier watchdog: block %d thread %d bar@%u parity %u\n", (int)blockIdx.x,
               (int)threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// Whole-warp wait where only one lane polls (the issuer / producer warps share their scheduler with softmax warps:
// 32 polling lanes cost the same issue slots as one, but a failed try_wait re-issues less often from one lane).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
#ifdef MAGI_POLL_ONE_LANE
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
#else
  mbar_wait(bar, parity);
#endif
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 3-D tiled load: coordinates are (c0 = innermost element, c1, c2); completes `bytes` on the mbarrier.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, "
      "%5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mbarrier arrives once every tcgen05 op previously issued by this thread has completed.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------- CTA pair (cluster of 2, tcgen05 cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// arrive on an mbarrier of another CTA of the cluster (address from mapa_u32). Default semantics on purpose (what the
// vendor's cluster barrier does): `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR in front of every arrive
// (measured: the pair kernel ran at half speed with it). What the arrivals order here are TMEM accesses, which
// tcgen05.wait / tcgen05.fence::before_thread_sync already cover.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// wait on a barrier of THIS CTA whose arrivals come from both CTAs of the pair (cluster-scope acquire), with watchdog
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (((++spins) & 0xfffu) == 0) {
      uint64_t t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > MAGI_WATCHDOG_NS) {
        printf("[magi] mbarrier watchdog (cluster): block %d thread %d bar@%u parity %u\n", (int)blockIdx.x,
               (int)threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}
// In a CTA pair the shared::cluster addresses of the two CTAs differ in one bit; clearing it addresses the even
// (leader) CTA's copy of the same object (what the vendor's 2-SM TMA atoms do for the mbarrier operand).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
// 3-D tiled load issued by either CTA of a pair into ITS OWN shared memory, completing bytes on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// the mbarrier at the same shared-memory offset in BOTH CTAs of the pair arrives once every tcgen05 op previously
// issued by this thread has completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// D[tmem, 256 rows over the pair] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 columns per CTA]
__device__ __forceinline__ void mma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same with 8-bit floating point operands (formats in the instruction descriptor), K = 32 per instruction
__device__ __forceinline__ void mma_ss_pair_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem, each CTA its own 128 rows] * B[smem, N/2 columns per CTA]
__device__ __forceinline__ void mma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 format, version = 1), 128-byte swizzle.
//   lbo/sbo are byte offsets (multiples of 16); addr is a shared-space byte address.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // layout type: SWIZZLE_128B
  return d;
}

// Same, no swizzle (canonical 8-row x 16-byte core matrices; lbo / sbo = byte strides between core matrices along
// K / along M-N). Used for the all-ones operand of the row-sum MMA, where the exact element placement is irrelevant.
__device__ __forceinline__ uint64_t make_smem_desc_noswizzle(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
//   b_mn_major = 1 when B is stored N-contiguous (used for P*V with V kept [key][head_dim]).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                      // D format: F32
         | (1u << 7)                    // A format: BF16
         | (1u << 10)                   // B format: BF16
         | (uint32_t(a_mn_major) << 15) // A major
         | (uint32_t(b_mn_major) << 16) // B major
         | (uint32_t(N >> 3) << 17)     // N / 8
         | (uint32_t(M >> 4) << 24);    // M / 16
}

// Instruction descriptor for kind::f8f6f4, e4m3 x e4m3 -> fp32, both operands K-major (format code 0 = E4M3 in bits
// [7, 10) and [10, 13); K = 32 per instruction).
__host__ __device__ constexpr uint32_t make_idesc_e4m3(int M, int N) {
  return (1u << 4)                      // D format: F32
         | (uint32_t(N >> 3) << 17)     // N / 8
         | (uint32_t(M >> 4) << 24);    // M / 16
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives columns [c, c+32) of TMEM lane (base_lane + i).
#define MAGI_R32(r, o)                                                                                        \
  r[o + 0], r[o + 1], r[o + 2], r[o + 3], r[o + 4], r[o + 5], r[o + 6], r[o + 7], r[o + 8], r[o + 9],         \
      r[o + 10], r[o + 11], r[o + 12], r[o + 13], r[o + 14], r[o + 15], r[o + 16], r[o + 17], r[o + 18],      \
      r[o + 19], r[o + 20], r[o + 21], r[o + 22], r[o + 23], r[o + 24], r[o + 25], r[o + 26], r[o + 27],      \
      r[o + 28], r[o + 29], r[o + 30], r[o + 31]

#define MAGI_TMEM_LD32(taddr, r, o)                                                                           \
  asm volatile(                                                                                               \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
      : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),        \
        "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]),      \
        "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]), "=r"(r[o + 16]), "=r"(r[o + 17]),  \
        "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]),  \
        "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]),  \
        "=r"(r[o + 30]), "=r"(r[o + 31])                                                                      \
      : "r"(taddr)                                                                                            \
      : "memory")

// one column: thread i receives column c of TMEM lane (base_lane + i)
__device__ __forceinline__ uint32_t tmem_ld_32x32b_x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}

#define MAGI_TMEM_ST32(taddr, r, o)                                                                           \
  asm volatile(                                                                                               \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                         \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                              \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),        \
      "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), \
      "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]),             \
      "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15]), "r"(r[o + 16]), "r"(r[o + 17]), "r"(r[o + 18]),          \
      "r"(r[o + 19]), "r"(r[o + 20]), "r"(r[o + 21]), "r"(r[o + 22]), "r"(r[o + 23]), "r"(r[o + 24]),          \
      "r"(r[o + 25]), "r"(r[o + 26]), "r"(r[o + 27]), "r"(r[o + 28]), "r"(r[o + 29]), "r"(r[o + 30]),          \
      "r"(r[o + 31])                                                                                          \
      : "memory")

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------- packed fp32x2 math (sm_100: FFMA2 / FADD2 / FMNMX3)
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fadd2_rm(uint64_t a, uint64_t b) {  // round toward -inf
  uint64_t d;
  asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// 2^x for a pair on the FMA/ALU pipes instead of the MUFU (Cody-Waite split + degree-3 polynomial on [0,1);
// max relative error ~9e-5, well below the bf16 rounding of P). Valid for x in [-127, 127]; smaller x clamp to 2^-127.
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& y0, float& y1) {
  const float kMagic = 12582912.f;  // 2^23 + 2^22: adding it leaves floor(x) in the low mantissa bits
  x0 = fmaxf(x0, -127.f);
  x1 = fmaxf(x1, -127.f);
  const uint64_t x = pack_f2(x0, x1);
  const uint64_t magic = pack_f2(kMagic, kMagic);
  const uint64_t xr = fadd2_rm(x, magic);
  const uint64_t fl = fsub2(xr, magic);        // floor(x)
  const uint64_t f = fsub2(x, fl);             // x - floor(x) in [0, 1)
  uint64_t pz = ffma2(pack_f2(0.077119089663028717f, 0.077119089663028717f), f,
                      pack_f2(0.227564394474029541f, 0.227564394474029541f));
  pz = ffma2(pz, f, pack_f2(0.695146143436431885f, 0.695146143436431885f));
  pz = ffma2(pz, f, pack_f2(1.0f, 1.0f));
  float p0, p1, r0, r1;
  unpack_f2(pz, p0, p1);
  unpack_f2(xr, r0, r1);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// pack two fp32 into bf16x2 (lo -> bits [0,16), hi -> bits [16,32)), round-to-nearest-even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

}  // namespace magi
