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
// HBM-bound: every element is read once (2 B) and written once (2 B); in the prologue 8 lanes own one (row, head)
// vector of 128 bf16 (two 16-byte pieces per lane), the helper kernels move 16 bytes per thread.
#include "common.cuh"
#include "prologue.cuh"

namespace magi {

// 16-byte store to an NVSwitch multicast (multimem) address: the switch replicates it to every GPU of the group.
__device__ __forceinline__ void multimem_st_v4(void* mc_ptr, const uint4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

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

}  // namespace

__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two bf16 in one 32-bit word -> packed (lo, hi) fp32 pair: one shift, one mask
__device__ __forceinline__ uint64_t bf2_to_f2(uint32_t u) {
  return pack_f2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(uint64_t v) {
  float lo, hi;
  unpack_f2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float quarter_warp_sum(float x) {   // over the 8 lanes that share one head vector
  x += __shfl_xor_sync(0xffffffffu, x, 4);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  return x;
}
// 8 consecutive parameters (fp32 or bf16) starting at element `off` -> 4 packed fp32 pairs
template <bool kBf16>
__device__ __forceinline__ void load_param8(const void* base, int off, uint64_t* v) {
  if constexpr (kBf16) {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    v[0] = bf2_to_f2(u.x); v[1] = bf2_to_f2(u.y); v[2] = bf2_to_f2(u.z); v[3] = bf2_to_f2(u.w);
  } else {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off + 4);
    v[0] = pack_f2(a.x, a.y); v[1] = pack_f2(a.z, a.w); v[2] = pack_f2(b.x, b.y); v[3] = pack_f2(b.z, b.w);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused prologue. The loads are taken off the compute warps: one producer warp per
// CTA streams "super-units" (kGroup consecutive work units of one tensor; a unit = (row, group of 8 heads) = 2 KB)
// into a shared-memory ring with cp.async.bulk - ONE 8 KB bulk copy per super-unit when the projection output is
// dense, one copy per head otherwise - completing on an mbarrier per stage. kConsumers warps pick the units up
// (warp w takes unit w % kGroup of every (w / kGroup)-th super-unit), read them with conflict-free LDS.128, normalise /
// rotate in packed fp32 and store straight to global memory. The bytes in flight are set by the ring depth (stages x
// 8 KB x CTAs per SM), not by the register budget of the compute warps: a register-only version of the same math
// topped out at 16 warps x 3 KB per SM (124 registers) and 78 % of the measured copy bandwidth, and a first ring with
// one 2 KB copy per barrier was producer-bound (~50 cycles of serial issue latency per copy); profiles/README.md.
// All fp32 arithmetic is packed (add/mul/fma.f32x2), 8 lanes own one head vector.
#ifndef MAGI_PRO_STAGES
#define MAGI_PRO_STAGES 10
#endif
#ifndef MAGI_PRO_CONSUMERS
#define MAGI_PRO_CONSUMERS 8
#endif
#ifndef MAGI_PRO_RING_BLOCKS
#define MAGI_PRO_RING_BLOCKS 2
#endif
constexpr int kRingUnitHeads = 8;
constexpr int kRingUnitBytes = kRingUnitHeads * 256;
constexpr int kRingGroup = 4;                                  // units per super-unit (4 x 8 heads = 32 producer lanes)
constexpr int kRingStageBytes = kRingGroup * kRingUnitBytes;
constexpr int kRingStages = MAGI_PRO_STAGES;
constexpr int kRingConsumers = MAGI_PRO_CONSUMERS;
constexpr int kRingConsumerGroups = kRingConsumers / kRingGroup;
static_assert(kRingConsumers % kRingGroup == 0, "consumer warps come in groups of kRingGroup");
// Every use of a stage must be consumed by the SAME group of warps: an mbarrier parity wait can only tell the current
// phase from the one before it, so a warp of another group arriving one use early would see "complete" spuriously
// (found with 3 groups x 10 stages: unspecified launch failure from the watchdog trap).
static_assert(kRingStages % kRingConsumerGroups == 0, "stages must be a multiple of the consumer groups");
constexpr int kRingThreads = (kRingConsumers + 1) * 32;
constexpr int kRingBarBytes = ((2 * kRingStages * 8 + 127) / 128) * 128;
constexpr int kRingSmemBytes = kRingBarBytes + kRingStages * kRingStageBytes;

__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

struct RingSuper {
  int type;        // 0 = q, 1 = k, 2 = v, 3 = qx (cross-attention query), 4 = text k, 5 = text v
  uint32_t u0;     // first unit (index inside the tensor's unit list)
  int cnt;         // valid units (1..kRingGroup)
};

// kMode 0: q / packed kv (+ KV-cache append) in place; 1: Ulysses send layout or peer receive buffers (head_perm_cp);
// 2: q in place, every K/V row stored into ALL ranks' gathered in-flight K/V buffers (fused all-gather, cp_ring_kv)
template <bool kLnBf16, int kMode>
__global__ void __launch_bounds__(kRingThreads, MAGI_PRO_RING_BLOCKS) prologue_ring_kernel(const PrologueParams p) {
  extern __shared__ __align__(128) uint8_t ring_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_smem);
  uint64_t* empty = full + kRingStages;
  const uint32_t ring = smem_u32(ring_smem + kRingBarBytes);
  const int lane = threadIdx.x & 31;
  const int warp = uniform(static_cast<int>(threadIdx.x >> 5));

  if (threadIdx.x == 0) {
    pdl_launch_dependents();
    for (int i = 0; i < kRingStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kRingGroup); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();   // programmatic dependent launch (common.cuh): nothing before this line touches global memory

  const uint32_t rows = static_cast<uint32_t>(p.rows);
  const uint32_t rows_t = static_cast<uint32_t>(p.rows_t);
  const uint32_t upr_q = (p.Hq + kRingUnitHeads - 1) / kRingUnitHeads;
  const uint32_t upr_kv = (p.Hkv + kRingUnitHeads - 1) / kRingUnitHeads;
  // units per tensor, tensor-major: q, k, v, then the cross-attention's qx, text k, text v
  const uint32_t n_q = p.x_q != nullptr ? rows * upr_q : 0u;
  const uint32_t n_k = p.x_k != nullptr ? rows * upr_kv : 0u;
  const uint32_t n_v = p.x_v != nullptr ? rows * upr_kv : 0u;
  const uint32_t n_x = p.x_qx != nullptr ? rows * upr_q : 0u;
  const uint32_t n_tk = p.x_tk != nullptr ? rows_t * upr_kv : 0u;
  const uint32_t n_tv = p.x_tv != nullptr ? rows_t * upr_kv : 0u;
  auto sus = [](uint32_t n) { return (n + kRingGroup - 1) / kRingGroup; };
  const uint32_t e_q = sus(n_q), e_k = e_q + sus(n_k), e_v = e_k + sus(n_v), e_x = e_v + sus(n_x);
  const uint32_t e_tk = e_x + sus(n_tk), total = e_tk + sus(n_tv);
  auto decode = [&](uint32_t su) {
    RingSuper d;
    d.type = su < e_q ? 0 : su < e_k ? 1 : su < e_v ? 2 : su < e_x ? 3 : su < e_tk ? 4 : 5;
    const uint32_t first = d.type == 0 ? 0u : d.type == 1 ? e_q : d.type == 2 ? e_k : d.type == 3 ? e_v
                                                                              : d.type == 4 ? e_x : e_tk;
    const uint32_t n = d.type == 0 ? n_q : d.type == 1 ? n_k : d.type == 2 ? n_v : d.type == 3 ? n_x
                                                                          : d.type == 4 ? n_tk : n_tv;
    d.u0 = (su - first) * kRingGroup;
    d.cnt = static_cast<int>(min(static_cast<uint32_t>(kRingGroup), n - d.u0));
    return d;
  };

  if (warp == kRingConsumers) {
    // ------------------------------------------------------------------ producer
    uint32_t m = 0;
#pragma unroll 1
    for (uint32_t su = blockIdx.x; su < total; su += gridDim.x, ++m) {
      const uint32_t stage = m % kRingStages;
      const uint32_t phase = (m / kRingStages) & 1u;
      if (m >= kRingStages) mbar_wait(&empty[stage], phase ^ 1u);
      const RingSuper d = decode(su);
      const bool qlike = d.type == 0 || d.type == 3;
      const int H = qlike ? p.Hq : p.Hkv;
      const uint32_t upr = qlike ? upr_q : upr_kv;
      const int64_t rstride = d.type == 0 ? p.xq_row_stride : d.type == 1 ? p.xk_row_stride
                              : d.type == 2 ? p.xv_row_stride : d.type == 3 ? p.xqx_row_stride : p.xt_row_stride;
      const int64_t hstride = d.type == 0 ? p.xq_head_stride : d.type == 1 ? p.xk_head_stride
                              : d.type == 2 ? p.xv_head_stride : d.type == 3 ? p.xqx_head_stride : p.xt_head_stride;
      const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(
          d.type == 0 ? p.x_q : d.type == 1 ? p.x_k : d.type == 2 ? p.x_v : d.type == 3 ? p.x_qx
                                                                  : d.type == 4 ? p.x_tk : p.x_tv);
      const uint32_t dst = ring + stage * kRingStageBytes;
      if (hstride == 128 && rstride == static_cast<int64_t>(H) * 128 && (H % kRingUnitHeads) == 0) {
        // dense projection output: the units of a super-unit are consecutive 2 KB pieces of one array
        if (elect_one()) {
          mbar_expect_tx(&full[stage], d.cnt * kRingUnitBytes);
          bulk_g2s(dst, x + static_cast<int64_t>(d.u0) * (kRingUnitHeads * 128), d.cnt * kRingUnitBytes, &full[stage]);
        }
      } else {
        // strided rows / heads (views, packed [K|V] projection of the cross-attention): lane = (unit, head)
        const int sub = lane >> 3, hh = lane & 7;
        const uint32_t u = d.u0 + sub;
        const uint32_t r = upr == 1 ? u : u / upr;
        const int head = static_cast<int>(u - r * upr) * kRingUnitHeads + hh;
        const bool valid = sub < d.cnt && head < H;
        const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, valid));
        if (elect_one()) mbar_expect_tx(&full[stage], nvalid * 256);
        __syncwarp();
        if (valid) bulk_g2s(dst + lane * 256, x + r * rstride + head * hstride, 256, &full[stage]);
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    const int s = lane & 7;
    const int j = lane >> 3;
    const int sub = warp % kRingGroup;       // which unit of a super-unit this warp takes
    const int grp = warp / kRingGroup;
    // lane s < 6: elements [8s, 8s+8) and [48+8s, 48+8s+8) (RoPE partners in the same lane); lanes 6, 7: the
    // pass-through dims, ordered so that both LDS.128 of a quarter-warp are bank-conflict free and every STG.128
    // writes whole 32-byte sectors
    const int e0 = s < 6 ? 8 * s : 112 + 8 * (s - 6);
    const int e1 = s < 6 ? 48 + 8 * s : 96 + 8 * (s - 6);
    const bool rot = p.sin != nullptr && s < 6;
    const uint64_t one2 = pack_f2(1.f, 1.f);
    const uint32_t opaque_zero = static_cast<uint32_t>(p.batch) >> 30;   // 0 at run time, unknown at compile time
    int cur_type = -1;
    uint64_t w2[8], b2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { w2[i] = one2; b2[i] = 0; }

#pragma unroll 1
    for (uint32_t m = grp;; m += kRingConsumerGroups) {
      const uint64_t su64 = static_cast<uint64_t>(blockIdx.x) + static_cast<uint64_t>(m) * gridDim.x;
      if (su64 >= total) break;
      const RingSuper d = decode(static_cast<uint32_t>(su64));
      const int type = d.type;
      const uint32_t stage = m % kRingStages;
      const uint32_t phase = (m / kRingStages) & 1u;
      if (sub >= d.cnt) {                      // tail super-unit: nothing for this warp, keep the ring protocol going
        mbar_wait(&full[stage], phase);
        if (lane == 0) mbar_arrive(&empty[stage]);
        continue;
      }
      const bool qlike = type == 0 || type == 3;
      const bool is_v = type == 2 || type == 5;          // pass-through tensors
      const bool extra = type >= 3;                      // the cross-attention's tensors: in place, bf16 LN, no RoPE
      const uint32_t upr = qlike ? upr_q : upr_kv;
      const uint32_t unit = d.u0 + sub;
      const uint32_t r = upr == 1 ? unit : unit / upr;
      const int h0 = static_cast<int>(unit - r * upr) * kRingUnitHeads;
      const int H = qlike ? p.Hq : p.Hkv;

      if (type != cur_type) {
        cur_type = type;
        if (!is_v) {
          const void* w = type == 0 ? p.ln_q_w : type == 1 ? p.ln_k_w : type == 3 ? p.ln_qx_w : p.ln_tk_w;
          const void* b = type == 0 ? p.ln_q_b : type == 1 ? p.ln_k_b : type == 3 ? p.ln_qx_b : p.ln_tk_b;
          const bool bf16_params = kLnBf16 || extra;
          if (bf16_params) {
            load_param8<true>(w, e0, w2); load_param8<true>(w, e1, w2 + 4);
            load_param8<true>(b, e0, b2); load_param8<true>(b, e1, b2 + 4);
          } else {
            load_param8<false>(w, e0, w2); load_param8<false>(w, e1, w2 + 4);
            load_param8<false>(b, e0, b2); load_param8<false>(b, e1, b2 + 4);
          }
          if (p.zero_centered_gamma != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              w2[i] = fadd2(w2[i], one2);
              // bf16 parameters: the reference computes `self.weight + 1` in bf16, i.e. rounds the sum
              if (bf16_params) w2[i] = bf2_to_f2(f2_to_bf2(w2[i]));
            }
          }
        }
      }

      // sin / cos of the row: issued here, consumed only after the LayerNorm of the first step (the packing below
      // sits after the ring wait in program order, so the loads have the wait + LDS + LayerNorm time to land;
      // ncu showed the warp parked on these loads when they were unpacked right away)
      float4 c0 = make_float4(1.f, 1.f, 1.f, 1.f), c1 = c0, s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
      const bool rope = rot && type <= 1;
      if (rope) {
        const uint32_t trow = p.batch == 1 ? r : r / static_cast<uint32_t>(p.batch);
        const float4* c = reinterpret_cast<const float4*>(p.cos + static_cast<int64_t>(trow) * 48 + 8 * s);
        const float4* sp = reinterpret_cast<const float4*>(p.sin + static_cast<int64_t>(trow) * 48 + 8 * s);
        c0 = __ldg(c); c1 = __ldg(c + 1); s0 = __ldg(sp); s1 = __ldg(sp + 1);
      }

      __nv_bfloat16* out_base =
          type == 0 ? reinterpret_cast<__nv_bfloat16*>(p.q_out)
          : type == 3 ? reinterpret_cast<__nv_bfloat16*>(p.qx_out)
          : reinterpret_cast<__nv_bfloat16*>(extra ? p.tkv_out : p.kv_out) + (is_v ? 128 : 0);
      const int64_t out_tstride = type == 0 ? p.q_out_token_stride : type == 3 ? p.qx_out_token_stride
                                  : extra ? p.tkv_out_token_stride : p.kv_out_token_stride;
      const int64_t out_hstride = type == 0 ? p.q_out_head_stride : type == 3 ? p.qx_out_head_stride
                                  : extra ? p.tkv_out_head_stride : p.kv_out_head_stride;
      __nv_bfloat16* cache_row = nullptr;
      int64_t grow = 0;                          // kMode 2: row of this unit in the gathered (chunk-major) buffers
      if constexpr (kMode == 2) {
        const uint32_t rpc = static_cast<uint32_t>(p.gather_rows_per_chunk);
        const uint32_t chunk = r / rpc;
        grow = static_cast<int64_t>(chunk) * p.gather_chunk_stride + p.gather_row_offset + (r - chunk * rpc);
      }
      if constexpr (kMode == 0) {
        if ((type == 1 || type == 2) && p.cache != nullptr && static_cast<int64_t>(r) < p.cache_rows)
          cache_row = reinterpret_cast<__nv_bfloat16*>(p.cache) + (p.cache_row_offset + r) * p.cache_token_stride +
                      (type == 2 ? 128 : 0);
      }

      mbar_wait(&full[stage], phase);
      const uint32_t sbase = ring + stage * kRingStageBytes + sub * kRingUnitBytes + j * 256;
      uint4 ra[2], rb[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        ra[u] = lds128(sbase + u * 1024 + e0 * 2);
        rb[u] = lds128(sbase + u * 1024 + e1 * 2);
      }
      // Hand the stage back to the producer - but only once the LDS above have RETURNED. mbarrier.arrive carries no
      // register dependency on them and ptxas issued it right behind the loads (no scoreboard wait): with the
      // load/store unit backed up by NVLink peer stores (fused all-to-all / all-gather modes) the producer's next bulk
      // copy into this stage could land before the reads had happened, and a unit came out holding the data of the
      // super-unit one ring later (round 2: 1-3 K/V rows of 37 440 per call on 2 x B200, found by bench.py's cp_parity
      // leg). The arrive's address now depends on every loaded register through a zero the compiler cannot see.
      const uint32_t landed = (ra[0].x ^ ra[0].w ^ ra[1].x ^ ra[1].w ^ rb[0].x ^ rb[0].w ^ rb[1].x ^ rb[1].w) & opaque_zero;
      __syncwarp();
      if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(&empty[stage]) + landed));

#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int h = h0 + 4 * u + j;
        if (h0 + 4 * u >= H) break;                // warp-uniform: the whole step is past the last head
        uint4 pa = ra[u], pb = rb[u];
        if (!is_v) {
          uint64_t v[8];
          v[0] = bf2_to_f2(pa.x); v[1] = bf2_to_f2(pa.y); v[2] = bf2_to_f2(pa.z); v[3] = bf2_to_f2(pa.w);
          v[4] = bf2_to_f2(pb.x); v[5] = bf2_to_f2(pb.y); v[6] = bf2_to_f2(pb.z); v[7] = bf2_to_f2(pb.w);
          // ---- LayerNorm over 128 in fp32 (two-pass, biased variance)
          float lo, hi;
          unpack_f2(fadd2(fadd2(fadd2(v[0], v[1]), fadd2(v[2], v[3])), fadd2(fadd2(v[4], v[5]), fadd2(v[6], v[7]))),
                    lo, hi);
          const float mean = quarter_warp_sum(lo + hi) * (1.0f / 128.0f);
          const uint64_t mean2 = pack_f2(mean, mean);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fsub2(v[i], mean2);
          uint64_t sq_a = fmul2(v[0], v[0]), sq_b = fmul2(v[1], v[1]);
#pragma unroll
          for (int i = 2; i < 8; i += 2) { sq_a = ffma2(v[i], v[i], sq_a); sq_b = ffma2(v[i + 1], v[i + 1], sq_b); }
          unpack_f2(fadd2(sq_a, sq_b), lo, hi);
          const float var = quarter_warp_sum(lo + hi) * (1.0f / 128.0f);
          const float rstd = rsqrtf(var + p.eps);   // MUFU.RSQ (2 ulp in fp32: invisible after the bf16 rounding)
          const uint64_t rstd2 = pack_f2(rstd, rstd);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = ffma2(fmul2(v[i], rstd2), w2[i], b2[i]);
          // ---- NeoX RoPE: (x1, x2) -> (x1 cos - x2 sin, x2 cos + x1 sin); identity on lanes 6, 7 (cos 1, sin 0)
          if (rope) {
            const uint64_t cs2[4] = {pack_f2(c0.x, c0.y), pack_f2(c0.z, c0.w), pack_f2(c1.x, c1.y), pack_f2(c1.z, c1.w)};
            const uint64_t sn2[4] = {pack_f2(s0.x, s0.y), pack_f2(s0.z, s0.w), pack_f2(s1.x, s1.y), pack_f2(s1.z, s1.w)};
            const uint64_t nsn2[4] = {pack_f2(-s0.x, -s0.y), pack_f2(-s0.z, -s0.w), pack_f2(-s1.x, -s1.y),
                                      pack_f2(-s1.z, -s1.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint64_t x1 = v[i], x2 = v[4 + i];
              v[i] = ffma2(x1, cs2[i], fmul2(x2, nsn2[i]));
              v[4 + i] = ffma2(x2, cs2[i], fmul2(x1, sn2[i]));
            }
          }
          pa = make_uint4(f2_to_bf2(v[0]), f2_to_bf2(v[1]), f2_to_bf2(v[2]), f2_to_bf2(v[3]));
          pb = make_uint4(f2_to_bf2(v[4]), f2_to_bf2(v[5]), f2_to_bf2(v[6]), f2_to_bf2(v[7]));
        }
        if (h >= H) continue;

        auto store = [&](__nv_bfloat16* dst) {
          *reinterpret_cast<uint4*>(dst + e0) = pa;
          *reinterpret_cast<uint4*>(dst + e1) = pb;
        };
        if (kMode == 0 || extra) {
          store(out_base + r * out_tstride + h * out_hstride);
          if (cache_row != nullptr) store(cache_row + h * p.cache_head_stride);
        } else if constexpr (kMode == 2) {
          if (type == 0) {
            store(out_base + r * out_tstride + h * out_hstride);
          } else {
            // fused K/V all-gather: the same 16-byte stores, once per rank of the group (own buffer included), over
            // NVLink peer memory; row = (chunk, this rank's stripe offset + row in stripe)
            const int64_t off = grow * out_tstride + h * out_hstride + (type == 2 ? 128 : 0);
            if (p.gather_multicast) {
              // NVSwitch multicast mapping: multimem.st (an ordinary st to a multimem address is undefined in the PTX ISA)
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.peer_kv[0]) + off;
              multimem_st_v4(dst + e0, pa);
              multimem_st_v4(dst + e1, pb);
            } else {
              for (int c = 0; c < p.gather_peers; ++c) store(reinterpret_cast<__nv_bfloat16*>(p.peer_kv[c]) + off);
            }
          }
        } else {
          // all-to-all SEND layout "(cp seq) hn hd" in one local buffer, or - fused all-to-all - straight into the
          // receive buffer of the rank that owns the head slice, at this rank's row offset (NVLink peer stores)
          const int cp = p.head_perm_cp;
          auto dest = [&](int c) -> __nv_bfloat16* {
            if (p.use_peers != 0) {
              __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(type == 0 ? p.peer_q[c] : p.peer_kv[c]) +
                                    (type == 2 ? 128 : 0);
              return base + (p.peer_row_offset + r) * out_tstride;
            }
            return out_base + (static_cast<int64_t>(c) * rows + r) * out_tstride;
          };
          if (type == 0) {
            const int hn = p.Hq / cp;
            store(dest(h / hn) + static_cast<int64_t>(h % hn) * out_hstride);
          } else if (cp > p.Hkv) {
            // fewer kv heads than ranks: replicate (repeat_interleave, context_parallel.py:393-394)
            const int rep = cp / p.Hkv;
            for (int c = h * rep; c < (h + 1) * rep; ++c) store(dest(c));
          } else {
            const int hn = p.Hkv / cp;
            store(dest(h / hn) + static_cast<int64_t>(h % hn) * out_hstride);
          }
        }
      }
    }
    // Stores into OTHER GPUs' memory (fused all-to-all / all-gather over NVLink, unicast or NVSwitch multicast) are
    // consumed by the peers after a device-side barrier that follows this kernel in stream order. The barrier's signal
    // and these data stores may take different routes: make every store of this thread visible system-wide before the
    // kernel (and hence the signal) can complete. Found in round 2: once the per-call work-list kernel no longer sat
    // between the barrier and the attention launch, the peers occasionally read rows that had not landed yet.
    if (kMode != 0 && (p.use_peers != 0 || p.gather_peers > 0)) __threadfence_system();
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

int prologue_units(const PrologueParams& p, int64_t* units) {
  const int64_t upr_q = (p.Hq + kRingUnitHeads - 1) / kRingUnitHeads, upr_kv = (p.Hkv + kRingUnitHeads - 1) / kRingUnitHeads;
  *units = p.rows * ((p.x_q ? upr_q : 0) + (p.x_k ? upr_kv : 0) + (p.x_v ? upr_kv : 0) + (p.x_qx ? upr_q : 0)) +
           p.rows_t * ((p.x_tk ? upr_kv : 0) + (p.x_tv ? upr_kv : 0));
  return MAGI_PRO_RING_BLOCKS;
}
template <bool kLnBf16, int kMode>
static cudaError_t prologue_launch_t(const PrologueParams& p, int64_t units, int sms, cudaStream_t stream) {
  static bool attr_set = false;   // per template instance; idempotent
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(prologue_ring_kernel<kLnBf16, kMode>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  // persistent: one CTA per resident slot; a CTA needs at least one super-unit per consumer group
  int64_t blocks = static_cast<int64_t>(sms) * MAGI_PRO_RING_BLOCKS;
  const int64_t need = (units / kRingGroup + kRingConsumerGroups) / kRingConsumerGroups;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  return launch_kernel(prologue_ring_kernel<kLnBf16, kMode>, static_cast<int>(blocks), kRingThreads, kRingSmemBytes,
                       stream, pdl_enabled(), p);
}
cudaError_t prologue_launch(const PrologueParams& p, bool ln_bf16, int sms, cudaStream_t stream) {
  int64_t units = 0;
  prologue_units(p, &units);
  const int mode = p.gather_peers > 0 ? 2 : (p.head_perm_cp > 1 ? 1 : 0);
  if (ln_bf16)
    return mode == 2 ? prologue_launch_t<true, 2>(p, units, sms, stream)
                     : mode == 1 ? prologue_launch_t<true, 1>(p, units, sms, stream)
                                 : prologue_launch_t<true, 0>(p, units, sms, stream);
  return mode == 2 ? prologue_launch_t<false, 2>(p, units, sms, stream)
                   : mode == 1 ? prologue_launch_t<false, 1>(p, units, sms, stream)
                               : prologue_launch_t<false, 0>(p, units, sms, stream);
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
