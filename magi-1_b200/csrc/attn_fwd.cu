// Range-masked (chunk-causal) attention forward for sm_100a: tcgen05.mma with accumulators in TMEM,
// TMA-staged Q/K/V tiles, register online softmax, block-sparse skipping of masked chunk tiles by construction
// (only the (q-range x k-range) rectangles are ever enumerated), two-segment K/V source (KV cache ++ new tokens).
//
// Reference semantics being replaced (samiazirar/MAGI-1):
//   core_attention / full_attention / cross_attention, inference/model/dit/dit_module.py:1023-1136
//   _full_adjust_key_and_value's torch.cat + chunk().contiguous(), dit_module.py:914-946
//
// CTA = 384 threads, persistent (one per SM), one work item = 256 query rows x 1 query head x one range:
//   warps 0-3   softmax warpgroup 0 : rows [0,128)   of the item, owns S0/P0/O0 in TMEM
//   warps 4-7   softmax warpgroup 1 : rows [128,256) of the item, owns S1/P1/O1 in TMEM
//   warp  8     tcgen05.mma issuer (warp-uniform loop, elect.sync around the issue) + TMEM allocator
//   warp  9     TMA producer (same structure)
//   warps 10-11 idle (keep warpgroup granularity)
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_t (bf16) is written in place over S_t in two
// 64-key halves (columns [0,32) and [64,96) of S_t).
// One thread owns one query row (tcgen05.ld 32x32b), so row max / row sum need no cross-thread reduction.
#include <math.h>
#include <stdio.h>

#include "attn_fwd.cuh"
#include "common.cuh"

namespace magi {

namespace {

constexpr int kTileM = 128;               // query rows per softmax warpgroup
constexpr int kPairM = 2 * kTileM;        // query rows per work item
constexpr int kTileN = 128;               // keys per K/V tile
// Measured interleaved in one process (tools/ab_variants.py, profiles/r02o_ab.log, cfg2 self-attention at the power
// cap): Q single-buffered + 4 K/V stages 9.096 ms, + 5 stages 9.106, Q double-buffered + 3 stages 9.387 (9.185 with the
// predicated P V issue path): the K/V ring depth is worth more than prefetching the next item's Q, also for the
// 4-tile cross-attention items (364 vs 374 us).
#ifndef MAGI_Q_STAGES
#define MAGI_Q_STAGES 1     // 2: the next item's Q tiles are prefetched while the current item runs (3 K/V stages)
#endif
#ifndef MAGI_KV_STAGES
#define MAGI_KV_STAGES (MAGI_Q_STAGES == 1 ? 4 : 3)   // smem: (2 * Q stages + KV stages) * 32 KB <= 224 KB
#endif
// MAGI_DYNAMIC_SCHED=1: work items are handed out by an atomic counter (the TMA warp fetches the CTA's next item and
// publishes it to the other roles through a two-slot shared-memory mailbox) instead of the static boustrophedon walk.
// Items are sorted longest-first, so this is LPT list scheduling; it is what lets a launch that overlaps the tail of
// the kernel before it (launch_dependency = 1) do ALL of its work on the SMs that are already free.
#ifndef MAGI_DYNAMIC_SCHED
#define MAGI_DYNAMIC_SCHED 1
#endif
// kDecoded (template parameter): the mailbox carries the DECODED item (ranges, tile counts, head), not its index: the binary
// search over the plan's prefix sums and the range loads (three or four dependent global loads, ~2000 cycles) are done
// once, by the scheduler, ahead of time, instead of by every role at the start of every item.
// Two instantiations, chosen per launch (attn_fwd_launch): interleaved A/B (profiles/r02za_ab.log) has the decoded
// mailbox 9 % faster on the 4-tile cross-attention items (305 vs 335 us) and 1.2 % slower on the 73..293-tile
// self-attention items (ptxas allocates the tile loop differently), so launches whose whole key sequence is short take
// it and the others keep the index mailbox.
constexpr int kDecodedMaxKeys = 8192;
constexpr int kQStages = MAGI_Q_STAGES;
constexpr int kStages = MAGI_KV_STAGES;   // K/V smem ring depth (tiles; K and V alternate)
constexpr int kBlkBytes = 128 * 64 * 2;   // one 64-column (128 B wide) swizzle block of a 128-row tile
constexpr int kTileBytes = 2 * kBlkBytes; // 128 x 128 bf16
#ifndef MAGI_COL_SPLIT
#define MAGI_COL_SPLIT 1   // 1: one thread owns a whole 128-column row of S_t; 2: two threads (two warps) share it
#endif
constexpr int kSplit = MAGI_COL_SPLIT;
static_assert(kSplit == 1 || kSplit == 2, "MAGI_COL_SPLIT must be 1 or 2");
constexpr int kCols = 128 / kSplit;              // S columns per softmax thread
constexpr int kSoftmaxWarps = 8 * kSplit;
constexpr int kMmaWarp = kSoftmaxWarps;          // tcgen05 issuer + TMEM allocator
constexpr int kTmaWarp = kSoftmaxWarps + 1;
constexpr int kThreads = kSplit == 1 ? 384 : 576;   // 12 warps (3 warpgroups, setmaxnreg) or 18 warps (112 regs each)
constexpr uint32_t kIdescQK = make_idesc_bf16(128, 128, 0, 0);
constexpr uint32_t kIdescPV = make_idesc_bf16(128, 128, 0, 1);
constexpr float kRescaleThreshold = 8.0f; // log2 units: lazily rescale O only when the row max grew by > 2^8
#ifndef MAGI_EMU_MOD
#define MAGI_EMU_MOD 3                    // every MAGI_EMU_MOD-th column pair takes the polynomial exp2 (0 = never)
#endif
#ifndef MAGI_SOFTMAX_REGS
#define MAGI_SOFTMAX_REGS 208   // 256 softmax threads x 208 + 128 issuer/producer/idle threads x 88 = 384 x 168
#endif
#ifndef MAGI_AUX_REGS
#define MAGI_AUX_REGS 88
#endif
// setmaxnreg only moves registers inside the pool the CTA was launched with (384 threads x 168 registers): asking
// for more blocks the second softmax warpgroup forever.
static_assert(kSplit != 1 || 256 * MAGI_SOFTMAX_REGS + 128 * MAGI_AUX_REGS <= 384 * 168,
              "register pool of the CTA exceeded");
static_assert(MAGI_SOFTMAX_REGS % 8 == 0 && MAGI_AUX_REGS % 8 == 0, "setmaxnreg takes multiples of 8");
constexpr int kEmuMod = MAGI_EMU_MOD;
// Found by reading the SASS with the single-warp issue model, measured in round 2:
//   MAGI_STATIC_BARRIERS=1  the mbarriers live in static shared memory, so their addresses are link-time constants. With
//                           the struct behind the hand-aligned dynamic base, ptxas re-derives that base (S2R
//                           SR_CgaCtaId / SR_SWINHI, LEA, LOP3, ~15 instructions with long stalls) inside the divergent
//                           region of EVERY lane-0 mbarrier.arrive of the softmax warps instead of keeping it in a register.
// Also seen there: ptxas sinks the arrive that publishes the FIRST half of P below 100-200 instructions of the second
// half's arithmetic (nothing depends on it), so both halves become visible to the MMA issuer 90 instructions apart
// and P*V starts later than the source suggests. An empty `asm volatile` with "+r" operands on the second half's
// scores does not move it (it vanishes in PTX; the sinking is ptxas's): that needs a real dependency and a B200 to
// measure it.
#ifndef MAGI_STATIC_BARRIERS
#define MAGI_STATIC_BARRIERS 1   // measured (profiles/r02_variants.log): 1180 vs 1135 TFLOP/s sustained, 1373 vs 1312 burst
#endif
//   MAGI_EARLY_PUBLISH=1    that real dependency: the first publish's mbarrier.arrive returns its state token, and the
//                           second half's `-m * scale` operand is OR-ed with (token & 0) where the 0 is opaque to the
//                           compiler (Hq >> 30), one LOP3 per tile. Every FFMA2 of the second half then depends on the
//                           arrive, so ptxas has to issue it first.
#ifndef MAGI_EARLY_PUBLISH
#define MAGI_EARLY_PUBLISH 0
#endif
//   MAGI_GUARD_BY_SUM=1     steady-state tiles (j > 0) no longer track the row max over all 128 columns. The running max
//                           m is only a REFERENCE POINT for the exponentials (softmax is shift-invariant and fp32 / bf16
//                           share an 8-bit exponent), so all that has to be guaranteed before a half of P is published
//                           is that no 2^x overflowed: for the first half that is read off its row sum (already
//                           computed: sum >= max, and an inf / NaN fails the comparison too), for the second half off
//                           a max over ITS 64 columns only (32 FMNMX3 instead of 108 FMNMX3 / FMNMX per tile). Either
//                           guard failing takes the same exact path as before (reload S, true tile max, rescale O).
//                           Measured interleaved (tools/ab_variants.py, profiles/r02r_ab.log, r02s_ab.log): 9.01-9.07 ms
//                           against 9.30-9.31 for the full max tracking on the cfg2 self-attention launch at the power
//                           cap (+3 %) once the K/V ring is 4 deep; neutral with the 3-deep ring of round 1.
#ifndef MAGI_GUARD_BY_SUM
#define MAGI_GUARD_BY_SUM 1
#endif
#ifndef MAGI_EPI_UNROLL
#define MAGI_EPI_UNROLL 2     // 32-column chunks of O per epilogue loop iteration (interleaved A/B r02zl_ab.log: 2 beats 1 and 4)
#endif
constexpr int kEpiUnroll = MAGI_EPI_UNROLL;
constexpr float kSumGuard = 32768.f;      // 2^15: a 64-column half of P summing to more than this re-references the row

#ifdef MAGI_TRACE
// Debug build only: clock64 stamps of CTA 0 (warp 0 of each softmax warpgroup, and the MMA issuer), first 512 tiles.
__device__ long long g_trace[3][512][8];
#define TRACE(who, idx, slot) \
  do { if (blockIdx.x == 0 && (idx) < 512) g_trace[who][idx][slot] = clock64(); } while (0)
// per-ITEM stamps of CTA 0 (softmax warp 0, MMA issuer, TMA warp): where the time between two items goes
__device__ long long g_itrace[3][128][8];
#define ITRACE(who, idx, slot) \
  do { if (blockIdx.x == 0 && (idx) < 128 && (threadIdx.x & 31) == 0) g_itrace[who][idx][slot] = clock64(); } while (0)
#else
#define TRACE(who, idx, slot) do {} while (0)
#define ITRACE(who, idx, slot) do {} while (0)
#endif

struct alignas(16) Item {
  int q_begin, q_end;  // query rows of this item's range
  int qrow0;           // first query row of the 256-row item
  int h, hkv;
  int ka0, ka1;        // key rows in segment A (cache), [ka0, ka1)
  int kb0, kb1;        // key rows in segment B (new tokens), segment-local
  int nt_a, nt;        // K/V tiles in segment A, total
  int index;           // position in the work list; -1 = no more work (mailbox)
};
static_assert(sizeof(Item) == 48, "the mailbox moves an Item as three 16-byte words");

struct Barriers {
  uint64_t q_full[kQStages], q_empty[kQStages];
  uint64_t kv_full[kStages], kv_empty[kStages];
  uint64_t s_full[2], p_half[2][2], o_full[2], o_free[2];
  uint64_t item_full[2], item_empty[2];   // MAGI_DYNAMIC_SCHED mailbox
  int sched[2];
  Item mail[2];
  uint32_t tmem_base;
  // exchange buffers of the two threads that share a row (only with MAGI_COL_SPLIT == 2)
  float xch[kSplit == 2 ? 2 : 1][2][2][kSplit == 2 ? 128 : 1];   // [parity][q tile][column half][row]: row max
  float xl[2][2][kSplit == 2 ? 128 : 1];                         // [q tile][column half][row]: row sum
};

constexpr int kSmemBytes = (2 * kQStages + kStages) * kTileBytes + 1024 /*align slack*/ + sizeof(Barriers);
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");


__device__ __forceinline__ bool decode_item(const int* __restrict__ ws, int item, int Hq, int group, int len_a,
                                            Item& it) {
  const int n = ws[1];
  const int* prefix = ws + 4;
  const int* rq0 = prefix + (n + 1);
  const int* rq1 = rq0 + n;
  const int* rk0 = rq1 + n;
  const int* rk1 = rk0 + n;
  // largest r with prefix[r] <= item
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= item) lo = mid; else hi = mid - 1;
  }
  const int r = lo;
  const int local = item - prefix[r];
  it.q_begin = rq0[r];
  it.q_end = rq1[r];
  const int nqt = (it.q_end - it.q_begin + kPairM - 1) / kPairM;
  // order inside a range: kv head (outer), q tile, q head within the GQA group (inner) -> neighbours share K/V in L2
  const int per_kv = nqt * group;
  it.hkv = local / per_kv;
  const int rem = local - it.hkv * per_kv;
  const int qt = rem / group;
  it.h = it.hkv * group + (rem - qt * group);
  it.qrow0 = it.q_begin + qt * kPairM;
  const int k0 = rk0[r], k1 = rk1[r];
  it.ka0 = min(k0, len_a);
  it.ka1 = min(k1, len_a);
  it.kb0 = max(k0, len_a) - len_a;
  it.kb1 = max(k1, len_a) - len_a;
  it.nt_a = (it.ka1 - it.ka0 + kTileN - 1) / kTileN;
  it.nt = it.nt_a + (it.kb1 - it.kb0 + kTileN - 1) / kTileN;
  it.index = item;
  return true;
}

__device__ __forceinline__ void tile_info(const Item& it, int j, int& seg, int& row0, int& nvalid) {
  if (j < it.nt_a) {
    seg = 0;
    row0 = it.ka0 + j * kTileN;
    nvalid = min(kTileN, it.ka1 - row0);
  } else {
    seg = 1;
    row0 = it.kb0 + (j - it.nt_a) * kTileN;
    nvalid = min(kTileN, it.kb1 - row0);
  }
}

// Descriptors differ only in their 14-bit start-address field, so the issuer keeps one base descriptor per operand
// and adds a compile-time constant (in 16-byte units) per MMA.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) { return make_smem_desc_sw128(addr, 16, 1024); }
__device__ __forceinline__ uint64_t desc_v(uint32_t addr) { return make_smem_desc_sw128(addr, kBlkBytes, 1024); }

// S[128 x 128] = Q[128 x 128(hd)] * K[128 x 128(hd)]^T ; both operands K-major, 128B-swizzled, two 64-wide blocks.
__device__ __forceinline__ void issue_qk(uint32_t d_tmem, uint64_t q_desc, uint64_t k_desc) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint64_t off = static_cast<uint64_t>(((ks >> 2) * kBlkBytes + (ks & 3) * 32) >> 4);
    mma_ss(d_tmem, q_desc + off, k_desc + off, kIdescQK, ks > 0 ? 1u : 0u);
  }
}

// O[128 x 128(hd)] (+)= P[128 x 128(keys)] * V[128(keys) x 128(hd)] ; P from TMEM (bf16, 2 per column),
// V is MN-major (head_dim contiguous): 8-key groups 1024 B apart (SBO), 64-wide hd blocks kBlkBytes apart (LBO).
// `half` selects keys [64 half, 64 half + 64) of the tile (k-steps 4 half .. 4 half + 3).
// `nks` = 16-key groups of the tile that hold at least one key of the range (8 for a full tile): the groups past the
// end of a partial last tile are not issued at all - their P is exactly 0, and skipping them keeps V rows that lie
// inside the tensor view but past the key range (possibly NaN / Inf padding) out of the accumulator.
__device__ __forceinline__ void issue_pv(uint32_t d_tmem, uint32_t p_tmem, uint64_t v_desc, uint32_t accumulate,
                                         int half, int nks) {
#pragma unroll
  for (int ks = 4 * half; ks < 4 * half + 4; ++ks) {
    // P half h (keys [64h, 64h+64), bf16) sits in columns [64h, 64h+32) of S_t: over the S columns its writer read
    if (ks < nks)
      mma_ts(d_tmem, p_tmem + 64 * half + (ks - 4 * half) * 8, v_desc + static_cast<uint64_t>((ks * 2048) >> 4),
             kIdescPV, (accumulate | (ks > 0)) ? 1u : 0u);   // only k-step 0 of a fresh accumulator overwrites
  }
}

// The common case of issue_pv - a full tile (nks == 8) - without the per-k-step predicates: the issuer's instruction
// count between "P half published" and the first UTCHMMA is on the critical path of every tile (the softmax
// warpgroup's next S_t can only follow P_t V), and the predicated form costs ~50 instructions there against ~15.
__device__ __forceinline__ void issue_pv_full(uint32_t d_tmem, uint32_t p_tmem, uint64_t v_desc, uint32_t accumulate,
                                              int half) {
#pragma unroll
  for (int ks = 4 * half; ks < 4 * half + 4; ++ks)
    mma_ts(d_tmem, p_tmem + 64 * half + (ks - 4 * half) * 8, v_desc + static_cast<uint64_t>((ks * 2048) >> 4), kIdescPV,
           ks > 0 ? 1u : accumulate);
}

__device__ __forceinline__ void stg256(void* ptr, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

}  // namespace

__global__ void attn_setup_kernel(const int* __restrict__ q_ranges, const int* __restrict__ k_ranges,
                                  const int* __restrict__ cu_q, const int* __restrict__ cu_k, int n, int Hq,
                                  int total_q, int total_k, int* __restrict__ ws) {
  int* prefix = ws + 4;
  int* rq0 = prefix + (n + 1);
  int* rq1 = rq0 + n;
  int* rk0 = rq1 + n;
  int* rk1 = rk0 + n;
  auto load = [&](int i, int& q0, int& q1, int& k0, int& k1) {
    if (q_ranges != nullptr) {
      q0 = q_ranges[2 * i]; q1 = q_ranges[2 * i + 1];
      k0 = k_ranges[2 * i]; k1 = k_ranges[2 * i + 1];
    } else {
      q0 = cu_q[i]; q1 = cu_q[i + 1];
      k0 = cu_k[i]; k1 = cu_k[i + 1];
    }
    q0 = max(0, min(q0, total_q)); q1 = max(q0, min(q1, total_q));
    k0 = max(0, min(k0, total_k)); k1 = max(k0, min(k1, total_k));
  };
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int q0, q1, k0, k1;
    load(i, q0, q1, k0, k1);
    const int lk = k1 - k0;
    int rank = 0;  // longest-key-range first (LPT order for the persistent CTAs)
    for (int j = 0; j < n; ++j) {
      int a0, a1, b0, b1;
      load(j, a0, a1, b0, b1);
      const int lj = b1 - b0;
      rank += (lj > lk) || (lj == lk && j < i);
    }
    rq0[rank] = q0; rq1[rank] = q1; rk0[rank] = k0; rk1[rank] = k1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < n; ++i) {
      prefix[i] = acc;
      acc += ((rq1[i] - rq0[i] + kPairM - 1) / kPairM) * Hq;
    }
    prefix[n] = acc;
    ws[0] = acc;
    ws[1] = n;
    int* counters = ws + attn_workspace_ints(n) - 2;   // {next item, CTAs finished}: MAGI_DYNAMIC_SCHED
    counters[0] = 0;
    counters[1] = 0;
  }
}

template <bool kDecoded>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_ka,
                const __grid_constant__ CUtensorMap tm_va, const __grid_constant__ CUtensorMap tm_kb,
                const __grid_constant__ CUtensorMap tm_vb, const AttnKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                               // kQStages x 2 tiles
  uint8_t* smem_kv = smem + 2 * kQStages * kTileBytes;  // kStages tiles
#if MAGI_STATIC_BARRIERS
  __shared__ Barriers bar_storage;
  Barriers* bar = &bar_storage;
#else
  Barriers* bar = reinterpret_cast<Barriers*>(smem + (2 * kQStages + kStages) * kTileBytes);
#endif

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = p.Hq / p.Hkv;
  const int* __restrict__ ws = p.work;

  if (threadIdx.x == 0) {
    pdl_launch_dependents();   // the next launch in the stream may take the SMs this grid leaves (common.cuh)
    for (int s = 0; s < kQStages; ++s) {
      mbar_init(&bar->q_full[s], 1);
      mbar_init(&bar->q_empty[s], 1);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar->kv_full[s], 1);
      mbar_init(&bar->kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bar->s_full[t], 1);
      mbar_init(&bar->p_half[t][0], kTileM / 32);   // one arrival per softmax warp, per 64-key half of P
      mbar_init(&bar->p_half[t][1], kTileM / 32);
      mbar_init(&bar->o_full[t], 1);
      mbar_init(&bar->o_free[t], kSplit * kTileM / 32);
      mbar_init(&bar->item_full[t], 1);
      mbar_init(&bar->item_empty[t], kSoftmaxWarps + 1);   // every softmax warp + the MMA issuer
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&bar->tmem_base, 512);
    tmem_relinquish();
  }
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_ka);
    tma_prefetch_desc(&tm_va);
    tma_prefetch_desc(&tm_kb);
    tma_prefetch_desc(&tm_vb);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!p.pdl_wait_at_end) pdl_wait();   // set-up above overlapped the tail of the kernel before; its results from here
  const uint32_t tmem_base = bar->tmem_base;
  const int total_items = ws[0];
#if MAGI_DYNAMIC_SCHED
  // consumer side of the item mailbox (softmax warps, MMA issuer): -1 = no more work
  uint32_t cslot = 0, cphase = 0;
  auto next_item = [&](Item& it) -> bool {
    mbar_wait(&bar->item_full[cslot], cphase);
    if constexpr (kDecoded) {
    // field by field into registers: writing `it` through an int4 pointer would pin it to local memory, and the tile
    // loop reads its fields once per tile
    const int4* src = reinterpret_cast<const int4*>(&bar->mail[cslot]);
    const int4 w0 = src[0], w1 = src[1], w2 = src[2];
    it.q_begin = w0.x; it.q_end = w0.y; it.qrow0 = w0.z; it.h = w0.w;
    it.hkv = w1.x; it.ka0 = w1.y; it.ka1 = w1.z; it.kb0 = w1.w;
    it.kb1 = w2.x; it.nt_a = w2.y; it.nt = w2.z; it.index = w2.w;
    } else {
      it.index = *reinterpret_cast<volatile int*>(&bar->sched[cslot]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar->item_empty[cslot]);
    if (++cslot == 2) { cslot = 0; cphase ^= 1; }
    if constexpr (!kDecoded) {
      if (it.index >= 0) decode_item(ws, it.index, p.Hq, group, p.len_a, it);
    }
    return it.index >= 0;
  };
#endif

  if (warp < kSoftmaxWarps) {
    // ====================================================================== softmax warpgroups
    if constexpr (kSplit == 1) setmaxnreg_inc<MAGI_SOFTMAX_REGS>();
    const int grp = warp >> 2;                 // warpgroup index
    const int t = grp / kSplit;                // which 128-row query tile of the item
    const int ch = grp % kSplit;               // which column half of S_t / O_t this thread owns (kSplit == 2)
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int rowq = quarter * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + t * 128;          // S_t (and P_t)
    const uint32_t tSc = tS + ch * kCols;                        // this thread's S columns
    const uint32_t tO = tmem_base + lane_sel + 256 + t * 128 + ch * kCols;   // this thread's O columns
    const int row_in_item = t * kTileM + rowq;
    const int pair_bar = 1 + t * 4 + quarter;  // named barrier shared by the two warps that split a row (kSplit == 2)
    uint32_t s_phase = 0, o_phase = 0;
#ifdef MAGI_TRACE
    int s_items = 0;
#endif
    const float scale_log2 = p.scale_log2;
#if MAGI_EARLY_PUBLISH
    const uint32_t opaque_zero = static_cast<uint32_t>(p.Hq) >> 30;   // 0 at run time, unknown at compile time
#endif
    __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(p.out);
    auto pair_max = [&](float v, int par) -> float {   // max over the two threads that share this row
      if constexpr (kSplit == 1) return v;
      bar->xch[par][t][ch][rowq] = v;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      return fmaxf(v, bar->xch[par][t][ch ^ 1][rowq]);
    };

#if MAGI_DYNAMIC_SCHED
    Item it;
    while (next_item(it)) {
      const int item = it.index;
#else
    for (int round = 0; round * static_cast<int>(gridDim.x) < total_items; ++round) {
      // items are sorted longest-first; CTAs walk them boustrophedon (0..G-1, G-1..0, ...) so that the CTA that drew
      // the longest item of one round draws the shortest of the next: static LPT-like balance without atomics
      const int item = round * static_cast<int>(gridDim.x) +
                       ((round & 1) ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x)
                                    : static_cast<int>(blockIdx.x));
      if (item >= total_items) continue;
      Item it;
      decode_item(ws, item, p.Hq, group, p.len_a, it);
#endif
      const int row = it.qrow0 + row_in_item;
      const bool row_valid = row < it.q_end;
      float m = -INFINITY, l = 0.f;   // running max of the row (shared by its threads), this thread's partial row sum
#ifdef MAGI_TRACE
      if (warp == 0) ITRACE(0, s_items, 0);
#endif

      for (int j = 0; j < it.nt; ++j) {
        int seg, krow0, nvalid;
        tile_info(it, j, seg, krow0, nvalid);
        const bool tr = quarter == 0 && lane == 0 && item == 0 && ch == 0;
        if (tr) TRACE(t, j, 0);
        mbar_wait(&bar->s_full[t], s_phase);
        s_phase ^= 1;
        tc_fence_after();
        if (tr) TRACE(t, j, 1);
#ifdef MAGI_TRACE
        if (warp == 0 && j == 0) ITRACE(0, s_items, 1);
#endif

        uint32_t sr[kCols];
        auto load_s = [&]() {
#pragma unroll
          for (int c0 = 0; c0 < kCols; c0 += 32) MAGI_TMEM_LD32(tSc + c0, sr, c0);
          tmem_wait_ld();
          if (nvalid < kTileN) {
            // Only the last key tile of a segment is partial. The empty volatile asm keeps ptxas from if-converting
            // this block into 128 ISETP+SEL pairs executed on EVERY tile (a third of the loop's instructions).
            asm volatile("" ::: "memory");
#pragma unroll
            for (int c = 0; c < kCols; ++c)
              if (ch * kCols + c >= nvalid) sr[c] = 0xff800000u;  // -inf
          }
        };
        load_s();
        if (tr) TRACE(t, j, 2);

        // Row max over this thread's scores: FMNMX3 in four chains. For j > 0 it is NOT on the critical path: P is
        // computed optimistically against the running max m while the chains run on the ALU pipe next to it; only
        // if some row's max grew by more than 2^8 (rare) is the tile redone with the new max.
        float mxa, mxb, mxc, mxd;
        auto max_step = [&](int k, bool first = false) {   // k = 0 .. kCols/4 - 1: consumes columns 4k .. 4k+3
          const float a0 = __uint_as_float(sr[4 * k]), a1 = __uint_as_float(sr[4 * k + 1]);
          const float a2 = __uint_as_float(sr[4 * k + 2]), a3 = __uint_as_float(sr[4 * k + 3]);
          if (k == 0 || first) { mxa = fmaxf(a0, a1); mxb = fmaxf(a2, a3); mxc = mxa; mxd = mxb; }
          else if (k & 1) { mxc = fmax3(mxc, a0, a1); mxd = fmax3(mxd, a2, a3); }
          else { mxa = fmax3(mxa, a0, a1); mxb = fmax3(mxb, a2, a3); }
        };
        uint64_t sum_a, sum_b;
        // local columns [64 hl, 64 hl + 64) -> packed bf16 P in sr[64 hl .. 64 hl + 32) (in place), row sums accumulated
        // with_max: 0 = no max chains, 1 = all kCols columns are covered by max steps during the first local half,
        //           2 = only the columns of the SECOND half are (MAGI_GUARD_BY_SUM)
        auto half = [&](const int hl, const uint64_t scale2, const uint64_t negm2, const int with_max) {
#pragma unroll
          for (int c = 64 * hl; c < 64 * hl + 64; c += 2) {
            const int pair = c >> 1;
            if (with_max == 1) {
              if constexpr (kSplit == 1) max_step(pair);
              else if ((pair & 1) == 0) max_step(pair >> 1);
            } else if (with_max == 2) {
              if ((pair & 1) == 0) max_step(16 + (pair >> 1), pair == 0);
            }
            const uint64_t x2 = ffma2(pack_f2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), scale2, negm2);
            float x0, x1, p0, p1;
            unpack_f2(x2, x0, x1);
            if (kEmuMod > 0 && (pair % (kEmuMod > 0 ? kEmuMod : 1)) == (kEmuMod > 0 ? kEmuMod : 1) - 1) {
              ex2_poly2(x0, x1, p0, p1);
            } else {
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            const uint64_t p2 = pack_f2(p0, p1);
            if (pair & 1) sum_b = fadd2(sum_b, p2); else sum_a = fadd2(sum_a, p2);
            sr[64 * hl + ((c - 64 * hl) >> 1)] = pack_bf16x2(p0, p1);
          }
        };
        auto local_max = [&]() { return fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd)); };
        // publish the P half whose keys are [64 hg, 64 hg + 64) from local registers sr[64 hl ..]
#if MAGI_EARLY_PUBLISH
        uint32_t publish_token = 0;
#endif
        auto publish = [&](int hl, int hg) {
          MAGI_TMEM_ST32(tS + 64 * hg, sr, 64 * hl);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
#if MAGI_EARLY_PUBLISH
          if (lane == 0) publish_token = static_cast<uint32_t>(mbar_arrive_token(&bar->p_half[t][hg]));
#else
          if (lane == 0) mbar_arrive(&bar->p_half[t][hg]);
#endif
        };

        float alpha = 1.f;
        bool need = false;
        const uint64_t scale2 = pack_f2(scale_log2, scale_log2);
        sum_a = pack_f2(0.f, 0.f);
        sum_b = sum_a;
        if (j == 0) {
#pragma unroll
          for (int k = 0; k < kCols / 4; ++k) max_step(k);
          m = pair_max(local_max(), j & 1);
          if (tr) TRACE(t, j, 3);
          const float neg_m = -m * scale_log2;
          half(0, scale2, pack_f2(neg_m, neg_m), 0);
        } else {
          if (tr) TRACE(t, j, 3);
          const float neg_m = -m * scale_log2;
#if MAGI_GUARD_BY_SUM
          static_assert(kSplit == 1, "MAGI_GUARD_BY_SUM is written for one thread per row");
          half(0, scale2, pack_f2(neg_m, neg_m), 2);
          float mx = local_max();                    // max over the SECOND half's columns only
          {
            float g0, g1;
            unpack_f2(fadd2(sum_a, sum_b), g0, g1);
            // first half: guarded by its row sum (negated comparison: inf / NaN take the exact path too)
            need = !(g0 + g1 <= kSumGuard) || (mx - m) * scale_log2 > kRescaleThreshold;
          }
#else
          half(0, scale2, pack_f2(neg_m, neg_m), 1);
          const float mx = pair_max(local_max(), j & 1);
          need = (mx - m) * scale_log2 > kRescaleThreshold;
#endif
          if (__any_sync(0xffffffffu, need)) {
            // rare: redo the first half against the new max and bring O back into range before P*V accumulates.
            // Both threads of a row see the same mx, hence take the same decision.
            load_s();
#if MAGI_GUARD_BY_SUM
            // exact tile max from the reloaded scores (the first half of sr was overwritten by P)
#pragma unroll
            for (int k = 0; k < kCols / 4; ++k) max_step(k);
            mx = local_max();
            need = mx > m;
#endif
            if (need) {
              alpha = ex2_approx((m - mx) * scale_log2);
              m = mx;
            }
            sum_a = pack_f2(0.f, 0.f);
            sum_b = sum_a;
            const float neg_m2 = -m * scale_log2;
            half(0, scale2, pack_f2(neg_m2, neg_m2), 0);
#pragma unroll 1
            for (int c = 0; c < kCols; c += 32) {
              uint32_t orr[32];
              MAGI_TMEM_LD32(tO + c, orr, 0);
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < 32; ++k) orr[k] = __float_as_uint(__uint_as_float(orr[k]) * alpha);
              MAGI_TMEM_ST32(tO + c, orr, 0);
            }
          }
        }
        // the first 64 keys this thread owns are final: hand them to the MMA issuer so that P*V starts early
        publish(0, kSplit == 1 ? 0 : ch);
        if (tr) TRACE(t, j, 4);
        if constexpr (kSplit == 1) {
#if MAGI_EARLY_PUBLISH
          const float neg_m = __uint_as_float(__float_as_uint(-m * scale_log2) | (publish_token & opaque_zero));
#else
          const float neg_m = -m * scale_log2;
#endif
          half(1, scale2, pack_f2(neg_m, neg_m), 0);
          if (tr) TRACE(t, j, 5);
          publish(1, 1);
        }
        float s0, s1;
        unpack_f2(fadd2(sum_a, sum_b), s0, s1);
        l = l * alpha + (s0 + s1);
        if (tr) TRACE(t, j, 6);
      }

      // ------------------------------------------------------------------ epilogue for this row
      const int hd = p.peer_head_offset + it.h;   // head index in the destination (peer_head_offset = 0 without peers)
      const int64_t head_off =
          p.out_group_heads > 0
              ? static_cast<int64_t>(hd / p.out_group_heads) * p.out_group_stride +
                    static_cast<int64_t>(hd % p.out_group_heads) * p.out_head_stride
              : static_cast<int64_t>(hd) * p.out_head_stride;
      __nv_bfloat16* orow = out + static_cast<int64_t>(row) * p.out_token_stride + head_off + ch * kCols;
      bool store_ok = row_valid;
      if (p.num_peers > 0) {
        // fused output all-to-all: this row belongs to rank rho; store into ITS output buffer (NVLink peer memory)
        // at the row's local index and this rank's head slot - the epilogue's stores are the transfer
        int rho = -1;
        for (int i = 0; i < p.num_peers; ++i)
          if (row >= p.peer_bounds[i] && row < p.peer_bounds[i + 1]) rho = i;
        store_ok = row_valid && rho >= 0;
        if (store_ok)
          orow = reinterpret_cast<__nv_bfloat16*>(p.peer_out[rho]) + (row - p.peer_bounds[rho]) * p.out_token_stride +
                 head_off + ch * kCols;
      }
      if (it.nt > 0) {
        if constexpr (kSplit == 2) {   // total row sum = sum of the two column halves
          bar->xl[t][ch][rowq] = l;
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          l += bar->xl[t][ch ^ 1][rowq];
        }
#ifdef MAGI_TRACE
        if (warp == 0) ITRACE(0, s_items, 2);
#endif
        mbar_wait(&bar->o_full[t], o_phase);
        o_phase ^= 1;
        tc_fence_after();
#ifdef MAGI_TRACE
        if (warp == 0) ITRACE(0, s_items, 3);
#endif
        const float inv_l = 1.f / l;
#pragma unroll kEpiUnroll
        for (int c = 0; c < kCols; c += 32) {
          uint32_t orr[32];
          MAGI_TMEM_LD32(tO + c, orr, 0);
          tmem_wait_ld();
          if (store_ok) {
            // One thread owns one row, so the 32 lanes of a store instruction hit 32 different rows: with 16-byte
            // stores every one of them is HALF a 32-byte sector, and the per-item trace (tools/trace_items.py) showed
            // the epilogue at 5 700 cycles, 27 % of a 4-tile cross-attention item. 32-byte stores (STG.256) halve
            // the instructions and write whole sectors; they need 32-byte aligned rows (checked on the host).
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 16; ++k)
              w[k] = pack_bf16x2(__uint_as_float(orr[2 * k]) * inv_l, __uint_as_float(orr[2 * k + 1]) * inv_l);
            if (p.out_align32) {
              stg256(orow + c, w);
              stg256(orow + c + 16, w + 8);
            } else {
#pragma unroll
              for (int k = 0; k < 16; k += 4)
                *reinterpret_cast<uint4*>(orow + c + 2 * k) = make_uint4(w[k], w[k + 1], w[k + 2], w[k + 3]);
            }
          }
        }
#ifdef MAGI_TRACE
        if (warp == 0) { ITRACE(0, s_items, 4); ++s_items; }
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->o_free[t]);
        if (row_valid && ch == 0 && p.lse != nullptr)
          p.lse[static_cast<int64_t>(row) * p.Hq + it.h] = (m * scale_log2 + log2f(l)) * 0.6931471805599453f;
      } else if (store_ok) {
        // empty key interval: zeros / -inf (flash-attn convention for seqlen_k == 0)
#pragma unroll
        for (int c = 0; c < kCols; c += 8) *reinterpret_cast<uint4*>(orow + c) = make_uint4(0, 0, 0, 0);
        if (ch == 0 && p.lse != nullptr) p.lse[static_cast<int64_t>(row) * p.Hq + it.h] = -INFINITY;
      }
    }
    // fused output all-to-all: the rows stored into other GPUs' buffers must be visible system-wide before this kernel
    // (and the device-side barrier that follows it in stream order) completes - see the note in prologue.cu
    if (p.num_peers > 0) __threadfence_system();
  } else if (warp == kMmaWarp) {
    // ====================================================================== MMA issuer
    // The whole warp runs the loop (waits included) so that everything feeding the MMAs is warp-uniform; only the
    // tcgen05.mma / tcgen05.commit instructions themselves are issued by the elected lane.
    if constexpr (kSplit == 1) setmaxnreg_dec<MAGI_AUX_REGS>();
    const uint32_t tb = uniform(tmem_base);
    const uint32_t tS0 = tb, tS1 = tb + 128, tO0 = tb + 256, tO1 = tb + 384;
    const uint64_t q_desc_base = desc_kmajor(smem_u32(smem_q));
    const uint32_t kv_addr = smem_u32(smem_kv);
    uint32_t stage = 0, phase = 0;  // kv ring (consumer side)
    uint32_t p_phase0 = 0, p_phase1 = 0;
    uint32_t item_cnt = 0;
    auto advance = [&]() {
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    };
    // O_t (+)= P_t V in two 64-key halves, each issued as soon as the softmax warpgroup has published that half of P
    auto pv = [&](uint32_t tO, uint32_t tS, int t, uint32_t& ph, uint32_t vst_, uint32_t acc, int nks) {
      const uint64_t vd = desc_v(kv_addr + vst_ * kTileBytes);
#ifdef MAGI_NO_PV_FAST
      const bool full = false;
#else
      const bool full = nks == 8;   // warp-uniform
#endif
      mbar_wait_warp(&bar->p_half[t][0], ph);
      tc_fence_after();
      if (full) { if (elect_one()) issue_pv_full(tO, tS, vd, acc, 0); }
      else if (elect_one()) issue_pv(tO, tS, vd, acc, 0, nks);
      __syncwarp();
      mbar_wait_warp(&bar->p_half[t][1], ph);
      tc_fence_after();
      if (full) { if (elect_one()) issue_pv_full(tO, tS, vd, acc, 1); }
      else if (elect_one()) issue_pv(tO, tS, vd, acc, 1, nks);
      __syncwarp();
      ph ^= 1;
    };
    // A partial last tile of a segment (nvalid < 128): the 16-key group that straddles the end of the range still
    // enters P V, so its rows past the range are zeroed in shared memory once the tile has landed (at most 15 rows x
    // 256 B, this warp's 32 lanes); the groups after it are skipped by `nks`. Returns nks.
    auto v_tail = [&](const Item& it_, int j_, uint32_t vst_) -> int {
      int seg, krow0, nvalid;
      tile_info(it_, j_, seg, krow0, nvalid);
      nvalid = uniform(nvalid);
#ifdef MAGI_NO_VTAIL
      return 8;   // debug variant: the round-1 behaviour (every k-step issued, nothing zeroed)
#endif
      if (nvalid >= kTileN) return 8;
      const int r1 = (nvalid + 15) & ~15;
      for (int idx = lane; idx < (r1 - nvalid) * 16; idx += 32) {
        const int row = nvalid + (idx >> 4), c = idx & 15;   // 16-byte pieces: 0-7 in hd block 0, 8-15 in block 1
        *reinterpret_cast<uint4*>(smem_kv + vst_ * kTileBytes + (c >> 3) * kBlkBytes + row * 128 + (c & 7) * 16) =
            make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
      __syncwarp();
      return r1 >> 4;
    };
#if MAGI_DYNAMIC_SCHED
    Item it;
    while (next_item(it)) {
      const int item = it.index;
#else
    for (int round = 0; round * static_cast<int>(gridDim.x) < total_items; ++round) {
      const int item = round * static_cast<int>(gridDim.x) +
                       ((round & 1) ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x)
                                    : static_cast<int>(blockIdx.x));
      if (item >= total_items) continue;
      Item it;
      decode_item(ws, item, p.Hq, group, p.len_a, it);
#endif
      const int nt = uniform(it.nt);
      if (nt == 0) continue;
      ITRACE(1, item_cnt, 0);
      const uint32_t o_free_parity = (item_cnt & 1) ^ 1;
      const uint32_t qs = item_cnt % kQStages;                       // Q stage of this item
      const uint64_t q0_desc = q_desc_base + static_cast<uint64_t>((qs * 2 * kTileBytes) >> 4);
      const uint64_t q1_desc = q0_desc + static_cast<uint64_t>(kTileBytes >> 4);
      mbar_wait_warp(&bar->q_full[qs], (item_cnt / kQStages) & 1);
      ITRACE(1, item_cnt, 1);
      // K_0 -> S0(0), S1(0)
      mbar_wait_warp(&bar->kv_full[stage], phase);
      tc_fence_after();
      ITRACE(1, item_cnt, 2);
      uint32_t kst = stage;
      advance();
      if (elect_one()) {
        const uint64_t kd = desc_kmajor(kv_addr + kst * kTileBytes);
        issue_qk(tS0, q0_desc, kd);
        tc_commit(&bar->s_full[0]);
        issue_qk(tS1, q1_desc, kd);
        tc_commit(&bar->s_full[1]);
        tc_commit(&bar->kv_empty[kst]);
        if (nt == 1) tc_commit(&bar->q_empty[qs]);
      }
      __syncwarp();
      ITRACE(1, item_cnt, 3);
      // V_0 -> O0 = P0(0) V_0
      mbar_wait_warp(&bar->kv_full[stage], phase);
      uint32_t vst = stage;
      advance();
      int nks = v_tail(it, 0, vst);   // k-step count of the V tile in `vst`
      mbar_wait_warp(&bar->o_free[0], o_free_parity);
      pv(tO0, tS0, 0, p_phase0, vst, 0, nks);
      for (int i = 1; i < nt; ++i) {
        // K_i -> S0(i)   (P0(i-1) already consumed by the in-order PV0(i-1))
        mbar_wait_warp(&bar->kv_full[stage], phase);
        tc_fence_after();
        kst = stage;
        advance();
        const bool tr = item == 0;
        if (elect_one()) {
          if (tr) TRACE(2, i, 0);
          issue_qk(tS0, q0_desc, desc_kmajor(kv_addr + kst * kTileBytes));
          tc_commit(&bar->s_full[0]);
          if (tr) TRACE(2, i, 1);
        }
        __syncwarp();
        // O1 (+)= P1(i-1) V_{i-1} ; S1(i)
        if (i == 1) mbar_wait_warp(&bar->o_free[1], o_free_parity);
        if (tr && lane == 0) TRACE(2, i, 2);
        pv(tO1, tS1, 1, p_phase1, vst, i > 1 ? 1u : 0u, nks);
        if (elect_one()) {
          tc_commit(&bar->kv_empty[vst]);
          issue_qk(tS1, q1_desc, desc_kmajor(kv_addr + kst * kTileBytes));
          tc_commit(&bar->s_full[1]);
          tc_commit(&bar->kv_empty[kst]);
          if (i == nt - 1) tc_commit(&bar->q_empty[qs]);
          if (tr) TRACE(2, i, 3);
        }
        __syncwarp();
        // V_i -> O0 += P0(i) V_i
        mbar_wait_warp(&bar->kv_full[stage], phase);
        vst = stage;
        advance();
        nks = v_tail(it, i, vst);
        if (tr && lane == 0) TRACE(2, i, 4);
        pv(tO0, tS0, 0, p_phase0, vst, 1, nks);
        if (tr && lane == 0) { TRACE(2, i, 5); TRACE(2, i, 6); }
      }
      // O1 (+)= P1(nt-1) V_{nt-1}
      if (nt == 1) mbar_wait_warp(&bar->o_free[1], o_free_parity);
      if (elect_one()) tc_commit(&bar->o_full[0]);   // everything up to PV0(nt-1)
      __syncwarp();
      pv(tO1, tS1, 1, p_phase1, vst, nt > 1 ? 1u : 0u, nks);
      if (elect_one()) {
        tc_commit(&bar->kv_empty[vst]);
        tc_commit(&bar->o_full[1]);
      }
      __syncwarp();
      ITRACE(1, item_cnt, 4);
      ++item_cnt;
    }
  } else if (warp == kTmaWarp) {
    // ====================================================================== TMA producer (same structure)
    if constexpr (kSplit == 1) setmaxnreg_dec<MAGI_AUX_REGS>();
    uint32_t stage = 0, phase = 0;  // kv ring (producer side)
    uint32_t item_cnt = 0;
#if MAGI_DYNAMIC_SCHED
    // scheduler: draws this CTA's next item from the global counter once all loads of the current one are issued
    // (a few tiles before the other roles finish it) and posts it in the mailbox
    uint32_t sslot = 0, sphase = 0;
    auto draw = [&](Item& it) -> bool {
      mbar_wait_warp(&bar->item_empty[sslot], sphase ^ 1);
      int item = 0;
      if (lane == 0) item = atomicAdd(p.counters, 1);
      item = uniform(item);
      if (item >= total_items) item = -1;
      it.index = item;
      if (item >= 0) decode_item(ws, item, p.Hq, group, p.len_a, it);
      if (lane == 0) {
        if constexpr (kDecoded) {
          int4* dst = reinterpret_cast<int4*>(&bar->mail[sslot]);
          dst[0] = make_int4(it.q_begin, it.q_end, it.qrow0, it.h);
          dst[1] = make_int4(it.hkv, it.ka0, it.ka1, it.kb0);
          dst[2] = make_int4(it.kb1, it.nt_a, it.nt, it.index);
        } else {
          *reinterpret_cast<volatile int*>(&bar->sched[sslot]) = item;
        }
        mbar_arrive(&bar->item_full[sslot]);
      }
      __syncwarp();
      if (++sslot == 2) { sslot = 0; sphase ^= 1; }
      return item >= 0;
    };
    Item it;
    while (draw(it)) {
      const int item = it.index;
#else
    for (int round = 0; round * static_cast<int>(gridDim.x) < total_items; ++round) {
      const int item = round * static_cast<int>(gridDim.x) +
                       ((round & 1) ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x)
                                    : static_cast<int>(blockIdx.x));
      if (item >= total_items) continue;
      Item it;
      decode_item(ws, item, p.Hq, group, p.len_a, it);
#endif
      it.nt = uniform(it.nt); it.nt_a = uniform(it.nt_a);
      it.h = uniform(it.h); it.hkv = uniform(it.hkv); it.qrow0 = uniform(it.qrow0);
      it.ka0 = uniform(it.ka0); it.ka1 = uniform(it.ka1); it.kb0 = uniform(it.kb0); it.kb1 = uniform(it.kb1);
      if (it.nt == 0) continue;
      ITRACE(2, item_cnt, 0);
      const uint32_t qs = item_cnt % kQStages;
      mbar_wait_warp(&bar->q_empty[qs], ((item_cnt / kQStages) & 1) ^ 1);
      ITRACE(2, item_cnt, 1);
      if (elect_one()) {
        uint8_t* qdst = smem_q + qs * 2 * kTileBytes;
        mbar_expect_tx(&bar->q_full[qs], 2 * kTileBytes);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          tma_load_3d(qdst + t * kTileBytes, &tm_q, &bar->q_full[qs], 0, it.h, it.qrow0 + t * kTileM);
          tma_load_3d(qdst + t * kTileBytes + kBlkBytes, &tm_q, &bar->q_full[qs], 64, it.h, it.qrow0 + t * kTileM);
        }
      }
      __syncwarp();
      ITRACE(2, item_cnt, 2);
      for (int j = 0; j < it.nt; ++j) {
        int seg, krow0, nvalid;
        tile_info(it, j, seg, krow0, nvalid);
        const CUtensorMap* mk = seg == 0 ? &tm_ka : &tm_kb;
        const CUtensorMap* mv = seg == 0 ? &tm_va : &tm_vb;
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
          const CUtensorMap* mm = kv == 0 ? mk : mv;
          mbar_wait_warp(&bar->kv_empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&bar->kv_full[stage], kTileBytes);
            uint8_t* dst = smem_kv + stage * kTileBytes;
            tma_load_3d(dst, mm, &bar->kv_full[stage], 0, it.hkv, krow0);
            tma_load_3d(dst + kBlkBytes, mm, &bar->kv_full[stage], 64, it.hkv, krow0);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      ITRACE(2, item_cnt, 3);
      ++item_cnt;
    }
  } else {
    if constexpr (kSplit == 1) setmaxnreg_dec<MAGI_AUX_REGS>();   // idle warps (setmaxnreg is warpgroup-collective)
    // fill_uncovered: query rows that no range covers come back as zeros / lse = -inf (what the drop-in wrappers of
    // flex_flash_attn_func / flash_attn_varlen_func promise) without a memset of the whole output before the launch:
    // the first otherwise idle warp of every CTA tests its share of the rows against the plan's ranges
#ifndef MAGI_NO_FILL   // (A/B builds only)
    if (p.fill_uncovered && warp == kTmaWarp + 1) {
      const int n = ws[1];
      const int* rq0 = ws + 4 + (n + 1);
      const int* rq1 = rq0 + n;
      const int per_cta = (p.total_q + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
      const int r_end = min(p.total_q, (static_cast<int>(blockIdx.x) + 1) * per_cta);
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
      for (int r = static_cast<int>(blockIdx.x) * per_cta + lane; r < r_end; r += 32) {
        bool covered = false;
        for (int i = 0; i < n; ++i) covered |= (r >= rq0[i]) & (r < rq1[i]);
        if (covered) continue;
        for (int h = 0; h < p.Hq; ++h) {
          __nv_bfloat16* orow = out + static_cast<int64_t>(r) * p.out_token_stride +
                                static_cast<int64_t>(h) * p.out_head_stride;
#pragma unroll
          for (int c = 0; c < 128; c += 8) *reinterpret_cast<uint4*>(orow + c) = make_uint4(0, 0, 0, 0);
          if (p.lse != nullptr) p.lse[static_cast<int64_t>(r) * p.Hq + h] = -INFINITY;
        }
      }
    }
#endif
  }

  if (p.pdl_wait_at_end) pdl_wait();    // complete in stream order even though nothing here depended on that kernel
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
#if MAGI_DYNAMIC_SCHED
  // the last CTA to leave rearms the counters: the plan is reused by the next layer's launch (stream-ordered after this)
  if (threadIdx.x == 0 && atomicAdd(p.counters + 1, 1) == static_cast<int>(gridDim.x) - 1) {
    p.counters[0] = 0;
    p.counters[1] = 0;
    __threadfence();
  }
#endif
}

#ifdef MAGI_TRACE
extern "C" __attribute__((visibility("default"))) int magi_debug_read_trace(long long* host, size_t bytes) {
  return (int)cudaMemcpyFromSymbol(host, g_trace, bytes < sizeof(g_trace) ? bytes : sizeof(g_trace));
}
extern "C" __attribute__((visibility("default"))) int magi_debug_read_itrace(long long* host, size_t bytes) {
  return (int)cudaMemcpyFromSymbol(host, g_itrace, bytes < sizeof(g_itrace) ? bytes : sizeof(g_itrace));
}
#endif

cudaError_t attn_setup_launch(const int* q_ranges, const int* k_ranges, const int* cu_q, const int* cu_k, int n,
                              int Hq, int total_q, int total_k, int* ws, cudaStream_t stream) {
  attn_setup_kernel<<<1, 128, 0, stream>>>(q_ranges, k_ranges, cu_q, cu_k, n, Hq, total_q, total_k, ws);
  return cudaGetLastError();
}

cudaError_t attn_fwd_launch(const CUtensorMap* maps, const AttnKernelParams& p, int num_ctas, int total_keys,
                            cudaStream_t stream) {
  static bool attr_set[64] = {false};   // per device; the attribute is sticky, so set it once instead of per call
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    if (dev < 64) attr_set[dev] = true;
  }
  static const int force = [] {   // MAGI_SCHED_DECODED=0 / 1 forces one instantiation (A/B runs)
    const char* v = getenv("MAGI_SCHED_DECODED");
    return v == nullptr ? -1 : (v[0] == '1' ? 1 : 0);
  }();
  const bool decoded = MAGI_DYNAMIC_SCHED && (force >= 0 ? force == 1 : total_keys <= kDecodedMaxKeys);
  if (decoded)
    return launch_kernel(attn_fwd_kernel<true>, num_ctas, kThreads, kSmemBytes, stream, pdl_enabled(), maps[0],
                         maps[1], maps[2], maps[3], maps[4], p);
  return launch_kernel(attn_fwd_kernel<false>, num_ctas, kThreads, kSmemBytes, stream, pdl_enabled(), maps[0], maps[1],
                       maps[2], maps[3], maps[4], p);
}

}  // namespace magi
