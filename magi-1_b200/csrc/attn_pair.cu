// Range-masked (chunk-causal) attention forward for sm_100a, CTA-PAIR version: the same work items, operands and results
// as attn_fwd.cu, re-pipelined so that the tensor pipe never waits for the softmax of the tile it has just produced.
//
// Reference semantics being replaced (samiazirar/MAGI-1):
//   core_attention / full_attention / cross_attention, inference/model/dit/dit_module.py:1023-1136
//   _full_adjust_key_and_value's torch.cat + chunk().contiguous(), dit_module.py:914-946
//
// Why a second kernel. In attn_fwd.cu a CTA owns two 128-row query tiles and TMEM is spent on S0 S1 O0 O1, so the bf16
// P tile has to alias the S tile it was computed from: S(i+1) cannot be issued before P(i) V(i) has drained, and every
// cycle of the softmax chain (S ready -> P published) beyond what the OTHER tile's two MMAs cover idles the tensor pipe
// (ncu: 68 % tensor-pipe active). Here two CTAs on the two SMs of a TPC form a cluster and issue
// `tcgen05.mma.cta_group::2` (M = 256): each CTA owns ONE 128-row query tile, so its 512 TMEM columns hold
//   S[0] [0,128)  S[1] [128,256)  O [256,384)  P[0] [384,448)  P[1] [448,512)
// S is double-buffered and P no longer aliases it: S(i+1) is issued while the softmax of tile i runs, P(i) V(i) as soon
// as P(i) is published. The B operands (K^T, V) are split between the two CTAs by the hardware - each CTA stages half
// of every K / V tile - which halves the shared-memory operand reads and the TMA traffic per SM that a one-tile CTA would
// otherwise double (B300_MICROARCH.md: 128 B/clk of shared memory per SM).
//
// CTA = 384 threads:
//   warps 0-7   softmax: warp w owns rows 32 (w % 4) + 16 (w / 4) + [0,16) of the CTA's query tile and ALL 128 key
//               columns of a tile, in the 16x256b TMEM fragment layout (a row's columns live in the 4 threads of a quad:
//               row max / row sum are two quad shuffles, nothing crosses a warp). Two warps per SM sub-partition.
//   warp  8     tcgen05.mma issuer (leader CTA only) + TMEM allocator (both CTAs)
//   warp  9     TMA producer (both CTAs: each loads its own Q tile and its half of every K / V tile)
//   warps 10-11 idle
// mbarriers signalled by tcgen05.commit are multicast to both CTAs; the ones the MMA issuer waits on live in the leader
// and are arrived on remotely by the peer CTA's warps / completed by the peer's TMA transactions.
#include <math.h>
#include <stdio.h>

#include "attn_fwd.cuh"
#include "common.cuh"
#include "tmem_shapes.cuh"

namespace magi {

namespace {

constexpr int kTileM = 128;              // query rows per CTA
constexpr int kPairM = 2 * kTileM;       // query rows per work item (one CTA pair)
constexpr int kTileN = 128;              // keys per K/V tile
constexpr int kQStages = 2;              // the next item's Q tile is prefetched while the current item runs
constexpr int kQBlk = 128 * 64 * 2;      // one 64-column (128 B wide) swizzle block of the 128-row Q tile
constexpr int kQBytes = 2 * kQBlk;       // 128 x 128 bf16
constexpr int kKBlk = 64 * 64 * 2;       // K half tile (this CTA's 64 keys): two blocks of 64 rows x 128 B
constexpr int kHalfBytes = 2 * kKBlk;    // = V half tile (128 keys x this CTA's 64 head dims): one block of 128 rows
#ifndef MAGI_PAIR_STAGES
#define MAGI_PAIR_STAGES 8               // ring of half tiles (K and V alternate): 4 K/V tiles in flight
#endif
constexpr int kStages = MAGI_PAIR_STAGES;
constexpr int kThreads = 384;
constexpr int kSoftmaxWarps = 8;
constexpr int kMmaWarp = 8;
constexpr int kTmaWarp = 9;
constexpr uint32_t kIdescQK = make_idesc_bf16(256, 128, 0, 0);
constexpr uint32_t kIdescPV = make_idesc_bf16(256, 128, 0, 1);
constexpr float kRescaleThreshold = 8.0f;  // log2 units: lazily rescale O only when the row max grew by > 2^8
#ifndef MAGI_PAIR_EMU_MOD
#define MAGI_PAIR_EMU_MOD 4                // every n-th column pair takes the polynomial exp2 (0 = never); 4 measured best
#endif
constexpr int kEmuMod = MAGI_PAIR_EMU_MOD;
// MAGI_PAIR_LSUM=1: the softmax row sums come off the TENSOR pipe instead of 64 FADD2 per thread and tile: one more
// (tiny) MMA per tile, L[256 x 16] += P[256 x 128 keys] * ones[128 keys x 16], accumulates sum_j P_ij (of the bf16 P
// the P V product uses, so numerator and denominator see the same rounding) into 16 TMEM columns; the epilogue reads
// its row's sum from there. Its columns come from single-buffering P (P V of tile g has always drained long before the
// softmax of tile g + 1 stores: checked with one mbarrier test per tile).
// Measured (profiles/r02_pair_variants2.log): correct, not faster (1136 vs 1141 TFLOP/s) - the exponentials of the two
// warps that share an SM sub-partition are MUFU-bound and run in lockstep - so it is OFF by default.
#ifndef MAGI_PAIR_LSUM
#define MAGI_PAIR_LSUM 0
#endif
constexpr bool kLsum = MAGI_PAIR_LSUM != 0;
constexpr uint32_t kColS = 0, kColO = 256, kColP = 384, kColL = 448;
constexpr uint32_t kIdescL = make_idesc_bf16(256, 16, 0, 0);
constexpr int kOnesBytes = 1024;

#ifdef MAGI_TRACE
// Debug build only: clock64 stamps of the leader CTA of pair 0 (softmax warps 0 and 4, which share SM sub-partition
// 0, and the MMA issuer), first 512 tiles. who: 0 = warp 0, 1 = warp 4, 2 = issuer (S side), 3 = issuer (P V side).
__device__ long long g_ptrace[4][512][8];
#define PTRACE(who, idx, slot) \
  do { if ((idx) < 512u) g_ptrace[who][idx][slot] = clock64(); } while (0)
#else
#define PTRACE(who, idx, slot) do {} while (0)
#endif

struct PairBarriers {
  // arrived on in BOTH CTAs by the leader's multicast tcgen05.commit
  uint64_t s_full[2];          // S[b] written
  uint64_t kv_empty[kStages];  // the MMAs that read ring stage s have completed
  uint64_t q_empty[kQStages];
  uint64_t o_full;             // last P V of the item has completed
  uint64_t pv_done;            // one phase per P V (only waited on by the rare rescale path)
  // waited on by the MMA issuer: only the LEADER's copies are used
  uint64_t kv_full[kStages];   // 1 arrival (leader's expect_tx) + the bytes of both CTAs' halves
  uint64_t q_full[kQStages];
  uint64_t s_free[2];          // 16 arrivals: every softmax warp of the pair has S[b] in registers
  uint64_t p_full[2];          // 16 arrivals: P[b] stored (one buffer, hence one barrier, with MAGI_PAIR_LSUM)
  uint64_t o_free;             // 16 arrivals: O read by the epilogue
  uint32_t tmem_base;
  float xl[kTileM], xm[kTileM];  // row sum / row max handed from the fragment layout to the row-per-thread epilogue
};

constexpr int kSmemBytes = kQStages * kQBytes + kStages * kHalfBytes + kOnesBytes + 1024 /*align slack*/;
static_assert(kSmemBytes + sizeof(PairBarriers) <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");

struct Item {
  int q_begin, q_end;  // query rows of this item's range
  int qrow0;           // first query row of the 256-row item
  int h, hkv;
  int ka0, ka1;        // key rows in segment A (cache), [ka0, ka1)
  int kb0, kb1;        // key rows in segment B (new tokens), segment-local
  int nt_a, nt;        // K/V tiles in segment A, total
};

// Same work list and the same in-range order as attn_fwd.cu (kv head outer, q tile, q head of the GQA group inner).
__device__ __forceinline__ void decode_item(const int* __restrict__ ws, int item, int group, int len_a, Item& it) {
  const int n = ws[1];
  const int* prefix = ws + 4;
  const int* rq0 = prefix + (n + 1);
  const int* rq1 = rq0 + n;
  const int* rk0 = rq1 + n;
  const int* rk1 = rk0 + n;
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

// The static schedule of attn_fwd.cu over CTA PAIRS: items are sorted longest-first and the pairs walk them
// boustrophedon. `Walker` enumerates this pair's items that have at least one K/V tile.
struct Walker {
  int round, item;
  Item it;
  __device__ __forceinline__ bool next(const int* ws, int total, int pairs, int pair, int group, int len_a) {
    for (;;) {
      if (round * pairs >= total) return false;
      item = round * pairs + ((round & 1) ? pairs - 1 - pair : pair);
      ++round;
      if (item >= total) continue;
      decode_item(ws, item, group, len_a, it);
      if (it.nt > 0) return true;
    }
  }
};

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) { return make_smem_desc_sw128(addr, 16, 1024); }
__device__ __forceinline__ uint64_t desc_v(uint32_t addr) { return make_smem_desc_sw128(addr, kHalfBytes, 1024); }

// S[256 x 128] = Q[256 x 128(hd)] * K[128 x 128(hd)]^T over the pair: A = this CTA's 128 Q rows (two 64-wide K-major
// blocks kQBlk apart), B = this CTA's 64 keys (two blocks kKBlk apart).
__device__ __forceinline__ void issue_qk(uint32_t d_tmem, uint64_t q_desc, uint64_t k_desc) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint64_t qoff = static_cast<uint64_t>(((ks >> 2) * kQBlk + (ks & 3) * 32) >> 4);
    const uint64_t koff = static_cast<uint64_t>(((ks >> 2) * kKBlk + (ks & 3) * 32) >> 4);
    mma_ss_pair(d_tmem, q_desc + qoff, k_desc + koff, kIdescQK, ks > 0 ? 1u : 0u);
  }
}

// O[256 x 128(hd)] (+)= P[256 x 128(keys)] * V[128(keys) x 128(hd)]: A = P from each CTA's TMEM (bf16, 2 keys per
// column), B = this CTA's 64 head dims of V, MN-major (head_dim contiguous): 8-key groups 1024 B apart.
__device__ __forceinline__ void issue_pv(uint32_t d_tmem, uint32_t p_tmem, uint64_t v_desc, uint32_t accumulate) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    mma_ts_pair(d_tmem, p_tmem + ks * 8, v_desc + static_cast<uint64_t>((ks * 2048) >> 4), kIdescPV,
                (accumulate | (ks > 0)) ? 1u : 0u);
}

// L[256 x 16] (+)= P[256 x 128(keys)] * ones[128(keys) x 16]: every column of L is the row sum of P. B is a block of
// bf16 ones (any in-range placement reads ones, so one descriptor serves all eight k-steps).
__device__ __forceinline__ void issue_l(uint32_t d_tmem, uint32_t p_tmem, uint64_t ones_desc, uint32_t accumulate) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    mma_ts_pair(d_tmem, p_tmem + ks * 8, ones_desc, kIdescL, (accumulate | (ks > 0)) ? 1u : 0u);
}

}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_ka,
                 const __grid_constant__ CUtensorMap tm_va, const __grid_constant__ CUtensorMap tm_kb,
                 const __grid_constant__ CUtensorMap tm_vb, const AttnKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                         // kQStages Q tiles
  uint8_t* smem_kv = smem + kQStages * kQBytes;   // kStages half tiles
  uint8_t* smem_ones = smem_kv + kStages * kHalfBytes;
  __shared__ PairBarriers bar;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());   // 0 = leader (issues the MMAs)
  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int pairs = static_cast<int>(gridDim.x) >> 1;
  const int group = p.Hq / p.Hkv;
  const int* __restrict__ ws = p.work;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar.kv_full[s], 1);
      mbar_init(&bar.kv_empty[s], 1);
    }
    for (int s = 0; s < kQStages; ++s) {
      mbar_init(&bar.q_full[s], 1);
      mbar_init(&bar.q_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar.s_full[b], 1);
      mbar_init(&bar.s_free[b], 2 * kSoftmaxWarps);
      mbar_init(&bar.p_full[b], 2 * kSoftmaxWarps);
    }
    mbar_init(&bar.o_full, 1);
    mbar_init(&bar.pv_done, 1);
    mbar_init(&bar.o_free, 2 * kSoftmaxWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc_pair(&bar.tmem_base, 512);
    tmem_relinquish_pair();
  }
  if (kLsum && threadIdx.x < kOnesBytes / 4) {
    reinterpret_cast<uint32_t*>(smem_ones)[threadIdx.x] = 0x3f803f80u;   // bf16 1.0 x 2
    fence_proxy_async_smem();   // generic-proxy writes, read by the tensor core (async proxy)
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
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit / peer TMA
  tc_fence_after();
  const uint32_t tmem_base = bar.tmem_base;
  const int total_items = ws[0];

  if (warp < kSoftmaxWarps) {
    // ====================================================================== softmax warps
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int hh = warp >> 2;                      // which 16 rows of the quarter
    const int qd = lane & 3;                       // position in the quad: columns 8n + 2 qd, + 1
    const int rl0 = quarter * 32 + hh * 16 + (lane >> 2);   // local row of registers with h = 0 (h = 1: + 8)
    const uint32_t frag_lane = static_cast<uint32_t>(quarter * 32 + hh * 16) << 16;
    const uint32_t quart_lane = static_cast<uint32_t>(quarter * 32) << 16;
    // the leader's copies of the barriers this warp arrives on (the two buffers' barriers are 8 bytes apart)
    const uint32_t s_free_remote = mapa_u32(&bar.s_free[0], 0);
    const uint32_t p_full_remote = mapa_u32(&bar.p_full[0], 0);
    const uint32_t o_free_remote = mapa_u32(&bar.o_free, 0);
    const float scale_log2 = p.scale_log2;
    const uint64_t scale2 = pack_f2(scale_log2, scale_log2);
    __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(p.out);
    uint32_t g = 0;        // tiles processed so far (all items): S / P buffer = g & 1, phase = (g >> 1) & 1
    uint32_t items = 0;    // items with tiles processed so far: o_full phase

    for (int round = 0; round * pairs < total_items; ++round) {
      const int item = round * pairs + ((round & 1) ? pairs - 1 - pair : pair);
      if (item >= total_items) continue;
      Item it;
      decode_item(ws, item, group, p.len_a, it);
      // epilogue geometry: one thread per row, warps w and w + 4 take the two 64-column halves of O
      const int erow_l = quarter * 32 + lane;
      const int erow = it.qrow0 + rank * kTileM + erow_l;
      const bool erow_valid = erow < it.q_end;
      const int64_t head_off =
          p.out_group_heads > 0
              ? static_cast<int64_t>(it.h / p.out_group_heads) * p.out_group_stride +
                    static_cast<int64_t>(it.h % p.out_group_heads) * p.out_head_stride
              : static_cast<int64_t>(it.h) * p.out_head_stride;
      __nv_bfloat16* orow = out + static_cast<int64_t>(erow) * p.out_token_stride + head_off + hh * 64;
      bool store_ok = erow_valid;
      if (p.num_peers > 0) {
        // fused output all-to-all: the row belongs to rank rho; store into ITS output buffer (NVLink peer memory)
        int rho = -1;
        for (int i = 0; i < p.num_peers; ++i)
          if (erow >= p.peer_bounds[i] && erow < p.peer_bounds[i + 1]) rho = i;
        store_ok = erow_valid && rho >= 0;
        if (store_ok)
          orow = reinterpret_cast<__nv_bfloat16*>(p.peer_out[rho]) + (erow - p.peer_bounds[rho]) * p.out_token_stride +
                 static_cast<int64_t>(p.peer_head_offset + it.h) * p.out_head_stride + hh * 64;
      }
      if (it.nt == 0) {
        // empty key interval: zeros / -inf (flash-attn convention for seqlen_k == 0)
        if (store_ok) {
#pragma unroll
          for (int c = 0; c < 64; c += 8) *reinterpret_cast<uint4*>(orow + c) = make_uint4(0, 0, 0, 0);
          if (hh == 0 && p.lse != nullptr) p.lse[static_cast<int64_t>(erow) * p.Hq + it.h] = -INFINITY;
        }
        continue;
      }

      float m[2] = {-INFINITY, -INFINITY};   // reference max of rows rl0, rl0 + 8 (replicated over the quad)
      float l[2] = {0.f, 0.f};               // this thread's partial row sums (its 32 columns of every tile)

      for (int j = 0; j < it.nt; ++j, ++g) {
        int seg, krow0, nvalid;
        tile_info(it, j, seg, krow0, nvalid);
        const uint32_t b = g & 1, ph = (g >> 1) & 1;
        const uint32_t pb = kLsum ? 0u : b;   // P buffer
#ifdef MAGI_TRACE
        const bool tr = blockIdx.x == 0 && quarter == 0 && lane == 0;
#endif
#ifdef MAGI_TRACE
        if (tr) PTRACE(hh, g, 0);
#endif
        mbar_wait(&bar.s_full[b], ph);
        tc_fence_after();
#ifdef MAGI_TRACE
        if (tr) PTRACE(hh, g, 1);
#endif
        uint32_t s[64];   // s[4n + 2h + e] = S[row rl0 + 8h][column 8n + 2 qd + e]
        MAGI_TMEM_LD_16x256b_X16(tmem_base + frag_lane + kColS + b * 128, s, 0);
        tmem_wait_ld();
        // S[b] is in registers: the issuer may overwrite it with the tile after next
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(s_free_remote + 8 * b);
#ifdef MAGI_TRACE
        if (tr) PTRACE(hh, g, 2);
#endif
        if (nvalid < kTileN) {
          // only the last key tile of a segment is partial (the empty asm keeps this out of the common path)
          asm volatile("" ::: "memory");
#pragma unroll
          for (int n = 0; n < 16; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              if (8 * n + 2 * qd + e >= nvalid) { s[4 * n + e] = 0xff800000u; s[4 * n + 2 + e] = 0xff800000u; }
        }

        // row max over this thread's 32 columns of each row, then over the quad
        auto tile_max = [&](float (&mx)[2]) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float a = fmaxf(__uint_as_float(s[2 * h]), __uint_as_float(s[2 * h + 1]));
            float c = fmaxf(__uint_as_float(s[4 + 2 * h]), __uint_as_float(s[4 + 2 * h + 1]));
#pragma unroll
            for (int n = 2; n < 16; n += 2) {
              a = fmax3(a, __uint_as_float(s[4 * n + 2 * h]), __uint_as_float(s[4 * n + 2 * h + 1]));
              c = fmax3(c, __uint_as_float(s[4 * n + 4 + 2 * h]), __uint_as_float(s[4 * n + 4 + 2 * h + 1]));
            }
            a = fmaxf(a, c);
            a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1));
            a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 2));
            mx[h] = a;
          }
        };
        uint32_t pk[32];   // pk[2n + h] = bf16x2(P[row h][8n + 2 qd], P[..][+1]) = packed column 4n + qd
        float ts[2];       // this tile's partial row sums
        auto exps = [&]() {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float neg_m = -m[h] * scale_log2;
            const uint64_t negm2 = pack_f2(neg_m, neg_m);
            uint64_t sum_a = pack_f2(0.f, 0.f), sum_b = sum_a;
#pragma unroll
            for (int n = 0; n < 16; ++n) {
              const uint64_t x2 =
                  ffma2(pack_f2(__uint_as_float(s[4 * n + 2 * h]), __uint_as_float(s[4 * n + 2 * h + 1])), scale2, negm2);
              float x0, x1, p0, p1;
              unpack_f2(x2, x0, x1);
              if (kEmuMod > 0 && (n % (kEmuMod > 0 ? kEmuMod : 1)) == (kEmuMod > 0 ? kEmuMod : 1) - 1) {
                ex2_poly2(x0, x1, p0, p1);
              } else {
                p0 = ex2_approx(x0);
                p1 = ex2_approx(x1);
              }
              if constexpr (!kLsum) {
                const uint64_t p2 = pack_f2(p0, p1);
                if (n & 1) sum_b = fadd2(sum_b, p2); else sum_a = fadd2(sum_a, p2);
              }
              pk[2 * n + h] = pack_bf16x2(p0, p1);
            }
            float s0, s1;
            unpack_f2(fadd2(sum_a, sum_b), s0, s1);
            ts[h] = s0 + s1;
          }
        };

        float alpha[2] = {1.f, 1.f};
        if (j == 0) {
          tile_max(m);
          exps();
        } else {
          // optimistic: P against the reference max m while the max chains run beside it on the ALU pipe; only when
          // some row's max grew by more than 2^8 (rare) is O brought back into range and the tile redone.
          exps();
#ifdef MAGI_TRACE
          if (tr) PTRACE(hh, g, 3);
#endif
          float mx[2];
          tile_max(mx);
          const bool need0 = (mx[0] - m[0]) * scale_log2 > kRescaleThreshold;
          const bool need1 = (mx[1] - m[1]) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need0 || need1)) {
            if (need0) { alpha[0] = ex2_approx((m[0] - mx[0]) * scale_log2); m[0] = mx[0]; }
            if (need1) { alpha[1] = ex2_approx((m[1] - mx[1]) * scale_log2); m[1] = mx[1]; }
            // every P V of this item issued so far (up to tile j - 1) must have landed in O: P V number g - 1 overall
            mbar_wait(&bar.pv_done, (g - 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 128; c += 32) {
              uint32_t o[16];   // o[4n + 2h + e]: rows h, columns c + 8n + 2 qd + e
              MAGI_TMEM_LD_16x256b_X4(tmem_base + frag_lane + kColO + c, o, 0);
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha[(k >> 1) & 1]);
              MAGI_TMEM_ST_16x256b_X4(tmem_base + frag_lane + kColO + c, o, 0);
            }
            if constexpr (kLsum) {
              uint32_t o[8];   // the 16 row-sum columns
              MAGI_TMEM_LD_16x256b_X2(tmem_base + frag_lane + kColL, o, 0);
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha[(k >> 1) & 1]);
              MAGI_TMEM_ST_16x256b_X2(tmem_base + frag_lane + kColL, o, 0);
            }
            tmem_wait_st();
            exps();
          }
        }
        if constexpr (kLsum) {
          // single P buffer: P V (and the row-sum MMA) of the previous tile must have drained. It has, long ago, in
          // any steady state; the test is one try_wait.
          if (g > 0) {
            mbar_wait(&bar.pv_done, (g - 1) & 1);
            tc_fence_after();
          }
        }
#ifdef MAGI_TRACE
        if (tr) PTRACE(hh, g, 4);
#endif
        MAGI_TMEM_ST_16x128b_X16(tmem_base + frag_lane + kColP + pb * 64, pk, 0);
        tmem_wait_st();
#ifdef MAGI_TRACE
        if (tr) PTRACE(hh, g, 5);
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(p_full_remote + 8 * pb);
#ifdef MAGI_TRACE
        if (tr) PTRACE(hh, g, 6);
#endif
        l[0] = l[0] * alpha[0] + ts[0];
        l[1] = l[1] * alpha[1] + ts[1];
      }

      // ------------------------------------------------------------------ epilogue of the item
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        l[h] += __shfl_xor_sync(0xffffffffu, l[h], 1);
        l[h] += __shfl_xor_sync(0xffffffffu, l[h], 2);
      }
      if (qd == 0) {
        bar.xl[rl0] = l[0]; bar.xl[rl0 + 8] = l[1];
        bar.xm[rl0] = m[0]; bar.xm[rl0 + 8] = m[1];
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");   // warps w and w + 4 share a lane quarter
      float row_l = bar.xl[erow_l];
      const float row_m = bar.xm[erow_l];
      mbar_wait(&bar.o_full, items & 1);
      ++items;
      tc_fence_after();
      if constexpr (kLsum) {
        row_l = __uint_as_float(tmem_ld_32x32b_x1(tmem_base + quart_lane + kColL));
        tmem_wait_ld();
      }
      const float inv_l = 1.f / row_l;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t orr[32];
        MAGI_TMEM_LD32(tmem_base + quart_lane + kColO + hh * 64 + c, orr, 0);
        tmem_wait_ld();
        if (store_ok) {
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(orr[k + 0]) * inv_l, __uint_as_float(orr[k + 1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(orr[k + 2]) * inv_l, __uint_as_float(orr[k + 3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(orr[k + 4]) * inv_l, __uint_as_float(orr[k + 5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(orr[k + 6]) * inv_l, __uint_as_float(orr[k + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + k) = v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(o_free_remote);
      if (erow_valid && hh == 0 && p.lse != nullptr)
        p.lse[static_cast<int64_t>(erow) * p.Hq + it.h] = (row_m * scale_log2 + log2f(row_l)) * 0.6931471805599453f;
      // xl / xm are rewritten only after the next item's tiles, which need both warps of the quarter (p_full)
    }
    // fused output all-to-all: the rows stored into other GPUs' buffers must be visible system-wide before this kernel
    // (and the device-side barrier that follows it in stream order) completes - see the note in prologue.cu
    if (p.num_peers > 0) __threadfence_system();
  } else if (warp == kMmaWarp) {
    // ====================================================================== MMA issuer (leader CTA)
    // Flattened over this pair's tiles with the S cursor one tile ahead of the P V cursor:
    //   S(0) S(1) PV(0) S(2) PV(1) ... - across item boundaries too, so the next item's first Q K^T is issued before
    //   the current item's last P V. The whole warp runs the loop; only the issue sits behind elect.sync.
    if (rank == 0) {
      const uint32_t tb = uniform(tmem_base);
      const uint32_t q_addr = smem_u32(smem_q), kv_addr = smem_u32(smem_kv);
      uint32_t stage = 0, phase = 0;       // K/V ring (consumer side)
      uint32_t gs = 0, gp = 0;             // S tiles / P V tiles issued so far
      uint32_t items_s = 0, items_p = 0;   // items entered by the S / P V cursor
      auto advance = [&]() { if (++stage == kStages) { stage = 0; phase ^= 1; } };
      Walker cs{0, 0, {}}, cp{0, 0, {}};
      int js = 0, jp = 0;
      bool s_valid = cs.next(ws, total_items, pairs, pair, group, p.len_a);
      bool p_valid = cp.next(ws, total_items, pairs, pair, group, p.len_a);
      auto issue_s = [&]() {
        const int nt = uniform(cs.it.nt);
        const uint32_t qs = items_s % kQStages;
#ifdef MAGI_TRACE
        const bool tr = blockIdx.x == 0 && lane == 0;
        if (tr) PTRACE(2, gs, 0);
#endif
        if (js == 0) mbar_wait_cluster(&bar.q_full[qs], (items_s / kQStages) & 1);
        mbar_wait_cluster(&bar.kv_full[stage], phase);
#ifdef MAGI_TRACE
        if (tr) PTRACE(2, gs, 1);
#endif
        const uint32_t b = gs & 1;
        mbar_wait_cluster(&bar.s_free[b], ((gs >> 1) & 1) ^ 1);
        tc_fence_after();
#ifdef MAGI_TRACE
        if (tr) PTRACE(2, gs, 2);
#endif
        if (elect_one()) {
          issue_qk(tb + kColS + b * 128, desc_kmajor(q_addr + qs * kQBytes), desc_kmajor(kv_addr + stage * kHalfBytes));
          tc_commit_pair(&bar.s_full[b]);
          tc_commit_pair(&bar.kv_empty[stage]);
          if (js == nt - 1) tc_commit_pair(&bar.q_empty[qs]);
        }
        __syncwarp();
#ifdef MAGI_TRACE
        if (tr) PTRACE(2, gs, 3);
#endif
        advance();
        ++gs;
        if (++js == nt) {
          js = 0;
          ++items_s;
          s_valid = cs.next(ws, total_items, pairs, pair, group, p.len_a);
        }
      };
      auto issue_p = [&]() {
        const int nt = uniform(cp.it.nt);
#ifdef MAGI_TRACE
        const bool tr = blockIdx.x == 0 && lane == 0;
        if (tr) PTRACE(3, gp, 0);
#endif
        mbar_wait_cluster(&bar.kv_full[stage], phase);
        if (jp == 0) mbar_wait_cluster(&bar.o_free, (items_p & 1) ^ 1);
#ifdef MAGI_TRACE
        if (tr) PTRACE(3, gp, 1);
#endif
        const uint32_t b = kLsum ? 0u : (gp & 1);
        mbar_wait_cluster(&bar.p_full[b], kLsum ? (gp & 1) : ((gp >> 1) & 1));
        tc_fence_after();
#ifdef MAGI_TRACE
        if (tr) PTRACE(3, gp, 2);
#endif
        if (elect_one()) {
          issue_pv(tb + kColO, tb + kColP + b * 64, desc_v(kv_addr + stage * kHalfBytes), jp > 0 ? 1u : 0u);
          if constexpr (kLsum)
            issue_l(tb + kColL, tb + kColP + b * 64, make_smem_desc_noswizzle(smem_u32(smem_ones), 128, 256),
                    jp > 0 ? 1u : 0u);
          tc_commit_pair(&bar.kv_empty[stage]);
          tc_commit_pair(&bar.pv_done);
          if (jp == nt - 1) tc_commit_pair(&bar.o_full);
        }
        __syncwarp();
#ifdef MAGI_TRACE
        if (tr) PTRACE(3, gp, 3);
#endif
        advance();
        ++gp;
        if (++jp == nt) {
          jp = 0;
          ++items_p;
          p_valid = cp.next(ws, total_items, pairs, pair, group, p.len_a);
        }
      };
      if (s_valid) issue_s();
      while (p_valid) {
        if (s_valid) issue_s();
        issue_p();
      }
    }
  } else if (warp == kTmaWarp) {
    // ====================================================================== TMA producer (both CTAs)
    // Same flattened order as the issuer consumes: K(0) K(1) V(0) K(2) V(1) ... Each CTA loads its own 128 Q rows,
    // its 64 keys of every K tile and its 64 head dims of every V tile; every transaction completes on the LEADER's
    // full barrier, for which the leader's producer posts the expected bytes of both halves.
    uint32_t stage = 0, phase = 0;
    uint32_t items_k = 0;
    Walker ck{0, 0, {}}, cv{0, 0, {}};
    int jk = 0, jv = 0;
    bool k_valid = ck.next(ws, total_items, pairs, pair, group, p.len_a);
    bool v_valid = cv.next(ws, total_items, pairs, pair, group, p.len_a);
    auto advance = [&]() { if (++stage == kStages) { stage = 0; phase ^= 1; } };
    auto load_k = [&]() {
      const int nt = uniform(ck.it.nt);
      const int hkv = uniform(ck.it.hkv);
      if (jk == 0) {
        const uint32_t qs = items_k % kQStages;
        mbar_wait_warp(&bar.q_empty[qs], ((items_k / kQStages) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* qdst = smem_q + qs * kQBytes;
          const int h = ck.it.h, qrow = ck.it.qrow0 + rank * kTileM;
          if (rank == 0) mbar_expect_tx(&bar.q_full[qs], 2 * kQBytes);
          tma_load_3d_pair(qdst, &tm_q, &bar.q_full[qs], 0, h, qrow);
          tma_load_3d_pair(qdst + kQBlk, &tm_q, &bar.q_full[qs], 64, h, qrow);
        }
        __syncwarp();
      }
      int seg, krow0, nvalid;
      tile_info(ck.it, jk, seg, krow0, nvalid);
      seg = uniform(seg); krow0 = uniform(krow0);
      const CUtensorMap* mk = seg == 0 ? &tm_ka : &tm_kb;
      mbar_wait_warp(&bar.kv_empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* dst = smem_kv + stage * kHalfBytes;
        if (rank == 0) mbar_expect_tx(&bar.kv_full[stage], 2 * kHalfBytes);
        // this CTA's 64 keys: rows krow0 + 64 rank .. + 64 (the K maps have a 64-row box)
        tma_load_3d_pair(dst, mk, &bar.kv_full[stage], 0, hkv, krow0 + rank * 64);
        tma_load_3d_pair(dst + kKBlk, mk, &bar.kv_full[stage], 64, hkv, krow0 + rank * 64);
      }
      __syncwarp();
      advance();
      if (++jk == nt) {
        jk = 0;
        ++items_k;
        k_valid = ck.next(ws, total_items, pairs, pair, group, p.len_a);
      }
    };
    auto load_v = [&]() {
      const int nt = uniform(cv.it.nt);
      const int hkv = uniform(cv.it.hkv);
      int seg, krow0, nvalid;
      tile_info(cv.it, jv, seg, krow0, nvalid);
      seg = uniform(seg); krow0 = uniform(krow0);
      const CUtensorMap* mv = seg == 0 ? &tm_va : &tm_vb;
      mbar_wait_warp(&bar.kv_empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* dst = smem_kv + stage * kHalfBytes;
        if (rank == 0) mbar_expect_tx(&bar.kv_full[stage], 2 * kHalfBytes);
        // all 128 keys, this CTA's 64 head dims (the V maps have a 128-row box)
        tma_load_3d_pair(dst, mv, &bar.kv_full[stage], rank * 64, hkv, krow0);
      }
      __syncwarp();
      advance();
      if (++jv == nt) {
        jv = 0;
        v_valid = cv.next(ws, total_items, pairs, pair, group, p.len_a);
      }
    };
    if (k_valid) load_k();
    while (v_valid) {
      if (k_valid) load_k();
      load_v();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA may exit (or free TMEM) while the other can still signal it
  if (warp == kMmaWarp) tmem_dealloc_pair(tmem_base, 512);
}

#ifdef MAGI_TRACE
extern "C" __attribute__((visibility("default"))) int magi_debug_read_trace_pair(long long* host, size_t bytes) {
  return (int)cudaMemcpyFromSymbol(host, g_ptrace, bytes < sizeof(g_ptrace) ? bytes : sizeof(g_ptrace));
}
#endif

cudaError_t attn_pair_launch(const CUtensorMap* maps, const AttnKernelParams& p, int num_ctas, cudaStream_t stream) {
  static bool attr_set[64] = {false};   // per device; the attribute is sticky, so set it once instead of per call
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(attn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    if (dev < 64) attr_set[dev] = true;
  }
  num_ctas &= ~1;
  if (num_ctas < 2) num_ctas = 2;
  attn_pair_kernel<<<num_ctas, kThreads, kSmemBytes, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);
  return cudaGetLastError();
}

}  // namespace magi
