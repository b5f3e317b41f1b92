// extern "C" boundary of libmagi_attn.so (declared in include/magi_attn.h): argument checking, TMA tensor-map
// encoding, launches. No torch types, no device synchronisation, no persistent device allocations.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/magi_attn.h"
#include "attn_fwd.cuh"
#include "prologue.cuh"
#include "gating.cuh"
#include "rope.cuh"
#include "activation.cuh"
#include "qkv_gemm.cuh"
#include "fp8_quant.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define MAGI_CHECK(cond, ...) \
  do {                        \
    if (!(cond)) return fail(__VA_ARGS__); \
  } while (0)

#define MAGI_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(e__));    \
  } while (0)

#ifndef MAGI_DEFAULT_PAIR
#define MAGI_DEFAULT_PAIR 0   // which attention kernel runs when MAGI_ATTN_KERNEL is unset
#endif

#if MAGI_DEFAULT_PAIR
#define MAGI_DEFAULT_KERNEL_NAME "pair"
#else
#define MAGI_DEFAULT_KERNEL_NAME "single"
#endif

#ifndef MAGI_SOURCE_HASH
#define MAGI_SOURCE_HASH "unstamped"
#endif

#define MAGI_STR_(x) #x
#define MAGI_STR(x) MAGI_STR_(x)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time through the runtime API so the library loads on machines without a driver
// (the CPU-only symbol-export test) and carries no link-time dependency on libcuda.so.1.
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// (tokens, heads, 128) bf16 view -> 3-D tensor map {128, heads, tokens}, box {64, 1, 128}, 128B swizzle.
// Rows past `tokens` are zero-filled by the TMA unit, which is what makes ragged range edges safe.
int make_view_map(CUtensorMap* m, const magi_view& v, int heads, const char* what, int box_rows = 128) {
  EncodeTiledFn enc = get_encode_fn();
  MAGI_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  const bool empty = v.tokens <= 0 || v.ptr == nullptr;
  void* base = empty ? nullptr : const_cast<void*>(v.ptr);
  cuuint64_t dims[3] = {128, (cuuint64_t)heads, (cuuint64_t)(empty ? 1 : v.tokens)};
  cuuint64_t strides[2] = {(cuuint64_t)(empty ? 256 : v.head_stride * 2), (cuuint64_t)(empty ? 256 * heads : v.token_stride * 2)};
  if (empty) return 2;  // never read by the kernel (no tiles are scheduled on it); caller substitutes another map
  MAGI_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "%s: pointer must be 16-byte aligned", what);
  MAGI_CHECK(v.head_stride % 8 == 0 && v.token_stride % 8 == 0, "%s: strides must be multiples of 8 elements", what);
  MAGI_CHECK(v.head_stride >= 128, "%s: head stride < 128", what);
  cuuint32_t box[3] = {64, 1, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MAGI_CHECK(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (CUresult %d)", what, (int)r);
  return 0;
}

int device_sm_count(int* out) {
  static int cached[64] = {0};
  int dev = 0;
  MAGI_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && cached[dev] > 0) { *out = cached[dev]; return 0; }
  int n = 0;
  MAGI_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (dev < 64) cached[dev] = n;
  *out = n;
  return 0;
}

int grid_for(int64_t work_threads, int threads, int* grid) {
  int sms = 0;
  if (int rc = device_sm_count(&sms)) return rc;
  int64_t blocks = (work_threads + threads - 1) / threads;
  const int64_t cap = (int64_t)sms * 8;  // 8 resident 256-thread CTAs per SM; grid-stride beyond that
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  *grid = (int)blocks;
  return 0;
}

}  // namespace

extern "C" {

const char* magi_last_error(void) { return g_err; }

const char* magi_build_info(void) {
  return "magi_attn sm_100a tcgen05/TMEM/TMA; default attention kernel: " MAGI_DEFAULT_KERNEL_NAME
         "; source " MAGI_SOURCE_HASH "; built " __DATE__ " " __TIME__
#ifdef __CUDACC_VER_MAJOR__
         "; nvcc " MAGI_STR(__CUDACC_VER_MAJOR__) "." MAGI_STR(__CUDACC_VER_MINOR__) "." MAGI_STR(__CUDACC_VER_BUILD__)
#endif
         ;
}

uint64_t magi_launch_count(void) { return g_launches.load(); }

size_t magi_attn_workspace_bytes(int32_t n_ranges) {
  if (n_ranges < 0) n_ranges = 0;
  return magi::attn_workspace_ints(n_ranges) * sizeof(int);
}

int magi_attn_fwd(const magi_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(a != nullptr, "args is NULL");
  MAGI_CHECK(a->head_dim == MAGI_HEAD_DIM, "head_dim must be 128 (got %d)", a->head_dim);
  MAGI_CHECK(a->num_heads_q > 0 && a->num_heads_kv > 0 && a->num_heads_q % a->num_heads_kv == 0,
             "num_heads_q (%d) must be a positive multiple of num_heads_kv (%d)", a->num_heads_q, a->num_heads_kv);
  MAGI_CHECK(a->n_ranges >= 0, "n_ranges < 0");
  const bool by_ranges = a->q_ranges != nullptr && a->k_ranges != nullptr;
  const bool by_cu = a->cu_seqlens_q != nullptr && a->cu_seqlens_k != nullptr;
  MAGI_CHECK(by_ranges != by_cu, "give exactly one of (q_ranges,k_ranges) or (cu_seqlens_q,cu_seqlens_k)");
  MAGI_CHECK(a->out != nullptr || a->num_peers > 0, "out is NULL");
  MAGI_CHECK(a->num_peers >= 0 && a->num_peers <= MAGI_MAX_PEERS, "num_peers out of range");
  MAGI_CHECK((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->out_token_stride % 8 == 0 &&
                 a->out_head_stride % 8 == 0,
             "out must be 16-byte aligned with strides that are multiples of 8 elements");
  MAGI_CHECK(a->q.tokens >= 0 && a->k_a.tokens >= 0 && a->k_b.tokens >= 0, "negative token count");
  MAGI_CHECK(a->k_a.tokens == a->v_a.tokens && a->k_b.tokens == a->v_b.tokens, "k/v token counts differ");
  MAGI_CHECK(a->q.tokens < (1ll << 31) && a->k_a.tokens + a->k_b.tokens < (1ll << 31), "token counts exceed int32");
  if (a->n_ranges == 0 || a->q.tokens == 0) return 0;
  MAGI_CHECK(a->q.ptr != nullptr, "q is NULL");
  MAGI_CHECK(a->workspace != nullptr && a->workspace_bytes >= magi_attn_workspace_bytes(a->n_ranges),
             "workspace too small: need %zu bytes", magi_attn_workspace_bytes(a->n_ranges));
  MAGI_CHECK((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "workspace must be 16-byte aligned");

  // kernel choice: the CTA-pair kernel (attn_pair.cu) unless MAGI_ATTN_KERNEL=single (attn_fwd.cu) is set or fewer
  // than two SMs are asked for
  static const int kernel_env = [] {
    const char* e = getenv("MAGI_ATTN_KERNEL");
    return (e != nullptr && strcmp(e, "single") == 0) ? 1 : (e != nullptr && strcmp(e, "pair") == 0) ? 2 : 0;
  }();
  int sms = a->num_sms;
  if (sms <= 0) { if (int r2 = device_sm_count(&sms)) return r2; }
  MAGI_CHECK(a->kernel >= 0 && a->kernel <= 2, "kernel must be 0 (default), 1 (single) or 2 (pair)");
  const int kernel_sel = a->kernel != 0 ? a->kernel : kernel_env;
  // an explicit kernel = 2 with fill_uncovered_rows is rejected below; a default / environment choice falls back
  const bool use_pair = (kernel_sel == 2 || (kernel_sel == 0 && MAGI_DEFAULT_PAIR)) && sms >= 2 &&
                        !(a->fill_uncovered_rows != 0 && a->kernel != 2);
  const int k_box = use_pair ? 64 : 128;   // a CTA of a pair stages its 64 keys of every K tile

  CUtensorMap tm_q, tm_ka, tm_va, tm_kb, tm_vb;
  int rc = make_view_map(&tm_q, a->q, a->num_heads_q, "q");
  if (rc) return rc == 2 ? fail("q is empty") : rc;
  int rka = make_view_map(&tm_ka, a->k_a, a->num_heads_kv, "k_a", k_box);
  if (rka == 1) return 1;
  int rva = make_view_map(&tm_va, a->v_a, a->num_heads_kv, "v_a");
  if (rva == 1) return 1;
  int rkb = make_view_map(&tm_kb, a->k_b, a->num_heads_kv, "k_b", k_box);
  if (rkb == 1) return 1;
  int rvb = make_view_map(&tm_vb, a->v_b, a->num_heads_kv, "v_b");
  if (rvb == 1) return 1;
  MAGI_CHECK(rka == rva && rkb == rvb, "k and v of a segment must both be present or both be empty");
  // an empty segment is never read by the kernel; give it any valid descriptor
  if (rka == 2) { tm_ka = rkb == 0 ? tm_kb : tm_q; tm_va = rkb == 0 ? tm_vb : tm_q; }
  if (rkb == 2) { tm_kb = rka == 0 ? tm_ka : tm_q; tm_vb = rka == 0 ? tm_va : tm_q; }

  const int len_a = rka == 2 ? 0 : (int)a->k_a.tokens;
  const int len_b = rkb == 2 ? 0 : (int)a->k_b.tokens;
  int* ws = reinterpret_cast<int*>(a->workspace);

  if (!a->plan_ready) {
    MAGI_CUDA(magi::attn_setup_launch(a->q_ranges, a->k_ranges, a->cu_seqlens_q, a->cu_seqlens_k, a->n_ranges,
                                      a->num_heads_q, (int)a->q.tokens, len_a + len_b, ws, stream));
    g_launches++;
  }

  magi::AttnKernelParams p;
  p.out = a->out;
  p.out_token_stride = a->out_token_stride;
  p.out_head_stride = a->out_head_stride;
  p.lse = a->lse;
  p.work = ws;
  p.Hq = a->num_heads_q;
  p.Hkv = a->num_heads_kv;
  p.len_a = len_a;
  const float scale = a->softmax_scale > 0.f ? a->softmax_scale : 0.08838834764831845f;  // 1/sqrt(128)
  p.scale_log2 = scale * 1.4426950408889634f;
  // with peers the groups are groups of the DESTINATION's heads (peer_head_offset + h), which this call cannot count
  MAGI_CHECK(a->out_group_heads >= 0 &&
                 (a->out_group_heads == 0 || ((a->num_peers > 0 || a->num_heads_q % a->out_group_heads == 0) &&
                                              a->out_group_stride % 8 == 0)),
             "out_group_heads must divide num_heads_q and out_group_stride must be a multiple of 8 elements");
  p.out_group_heads = a->out_group_heads;
  p.out_group_stride = a->out_group_stride;
  p.num_peers = a->num_peers;
  p.peer_head_offset = a->num_peers > 0 ? a->peer_head_offset : 0;
  for (int i = 0; i < 8; ++i) { p.peer_out[i] = nullptr; p.peer_bounds[i] = 0; }
  p.peer_bounds[8] = 0;
  if (a->num_peers > 0) {
    MAGI_CHECK(a->lse == nullptr, "peer output routing excludes lse");
    MAGI_CHECK(a->out_group_heads == 0 || !use_pair, "grouped peer output needs the single-CTA kernel");
    MAGI_CHECK(a->peer_head_offset >= 0, "negative peer_head_offset");
    for (int i = 0; i < a->num_peers; ++i) {
      MAGI_CHECK(a->peer_out[i] && (reinterpret_cast<uintptr_t>(a->peer_out[i]) & 15) == 0, "bad peer_out[%d]", i);
      MAGI_CHECK(a->peer_row_bounds[i] <= a->peer_row_bounds[i + 1] && a->peer_row_bounds[i] >= 0,
                 "peer_row_bounds must be non-decreasing");
      p.peer_out[i] = a->peer_out[i];
      p.peer_bounds[i] = a->peer_row_bounds[i];
    }
    p.peer_bounds[a->num_peers] = a->peer_row_bounds[a->num_peers];
  }

  MAGI_CHECK(a->launch_dependency == 0 || (a->launch_dependency == 1 && a->plan_ready),
             "launch_dependency must be 0, or 1 together with plan_ready");
  p.pdl_wait_at_end = a->launch_dependency == 1;
  p.counters = ws + magi::attn_workspace_ints(a->n_ranges) - 2;
  MAGI_CHECK(a->fill_uncovered_rows == 0 || (a->num_peers == 0 && a->out_group_heads == 0 && !use_pair),
             "fill_uncovered_rows needs plain output addressing and the single-CTA kernel");
  {
    uintptr_t bits = reinterpret_cast<uintptr_t>(a->out);
    for (int i = 0; i < a->num_peers; ++i) bits |= reinterpret_cast<uintptr_t>(a->peer_out[i]);
    p.out_align32 = (bits & 31) == 0 && a->out_token_stride % 16 == 0 && a->out_head_stride % 16 == 0 &&
                    (a->out_group_heads == 0 || a->out_group_stride % 16 == 0);
#ifdef MAGI_NO_STG256   // (A/B builds only)
    p.out_align32 = 0;
#endif
  }
  p.fill_uncovered = a->fill_uncovered_rows != 0;
  p.total_q = (int)a->q.tokens;
  const CUtensorMap maps[5] = {tm_q, tm_ka, tm_va, tm_kb, tm_vb};
  if (use_pair) MAGI_CUDA(magi::attn_pair_launch(maps, p, sms, stream));
  else MAGI_CUDA(magi::attn_fwd_launch(maps, p, sms, len_a + len_b, stream));
  g_launches++;
  return 0;
}

int magi_attn_plan(const magi_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(a != nullptr, "args is NULL");
  MAGI_CHECK(a->num_heads_q > 0 && a->n_ranges >= 0, "num_heads_q must be positive and n_ranges non-negative");
  const bool by_ranges = a->q_ranges != nullptr && a->k_ranges != nullptr;
  const bool by_cu = a->cu_seqlens_q != nullptr && a->cu_seqlens_k != nullptr;
  MAGI_CHECK(by_ranges != by_cu, "give exactly one of (q_ranges,k_ranges) or (cu_seqlens_q,cu_seqlens_k)");
  MAGI_CHECK(a->q.tokens >= 0 && a->k_a.tokens >= 0 && a->k_b.tokens >= 0 && a->q.tokens < (1ll << 31) &&
                 a->k_a.tokens + a->k_b.tokens < (1ll << 31),
             "token counts must be non-negative and fit int32");
  if (a->n_ranges == 0) return 0;
  MAGI_CHECK(a->workspace != nullptr && a->workspace_bytes >= magi_attn_workspace_bytes(a->n_ranges),
             "workspace too small: need %zu bytes", magi_attn_workspace_bytes(a->n_ranges));
  MAGI_CHECK((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "workspace must be 16-byte aligned");
  MAGI_CUDA(magi::attn_setup_launch(a->q_ranges, a->k_ranges, a->cu_seqlens_q, a->cu_seqlens_k, a->n_ranges,
                                    a->num_heads_q, (int)a->q.tokens, (int)(a->k_a.tokens + a->k_b.tokens),
                                    reinterpret_cast<int*>(a->workspace), stream));
  g_launches++;
  return 0;
}

int magi_qk_norm_rope_fwd(const magi_prologue_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(a != nullptr, "args is NULL");
  MAGI_CHECK(a->rows >= 0, "rows < 0");
  MAGI_CHECK(a->x_q != nullptr || a->x_k != nullptr || a->x_v != nullptr || a->x_qx != nullptr ||
                 a->x_tk != nullptr || a->x_tv != nullptr,
             "no input tensor");
  MAGI_CHECK(a->ln_dtype == 0 || a->ln_dtype == 1, "ln_dtype must be 0 (fp32) or 1 (bf16)");
  MAGI_CHECK(a->batch >= 1, "batch must be >= 1");
  if (a->x_q) {
    MAGI_CHECK((a->q_out || a->num_peers > 0) && a->ln_q_w && a->ln_q_b, "q path needs q_out, ln_q_w, ln_q_b");
    MAGI_CHECK(a->num_heads_q > 0 && a->xq_row_stride % 8 == 0 && a->q_out_token_stride % 8 == 0 &&
                   a->q_out_head_stride % 8 == 0,
               "q strides must be multiples of 8 elements");
  }
  MAGI_CHECK(a->xq_head_stride % 8 == 0 && a->xk_head_stride % 8 == 0 && a->xv_head_stride % 8 == 0 &&
                 a->xq_head_stride >= 0 && a->xk_head_stride >= 0 && a->xv_head_stride >= 0,
             "input head strides must be non-negative multiples of 8 elements");
  if (a->x_k) MAGI_CHECK((a->kv_out || a->num_peers > 0 || a->gather_peers > 0) && a->ln_k_w && a->ln_k_b, "k path needs kv_out, ln_k_w, ln_k_b");
  if (a->x_v) MAGI_CHECK(a->kv_out || a->num_peers > 0 || a->gather_peers > 0, "v path needs kv_out");
  if (a->x_k || a->x_v)
    MAGI_CHECK(a->num_heads_kv > 0 && a->xk_row_stride % 8 == 0 && a->xv_row_stride % 8 == 0 &&
                   a->kv_out_token_stride % 8 == 0 && a->kv_out_head_stride % 8 == 0,
               "kv strides must be multiples of 8 elements");
  if (a->cache)
    MAGI_CHECK(a->cache_token_stride % 8 == 0 && a->cache_head_stride % 8 == 0 && a->cache_rows >= 0 &&
                   a->cache_rows <= a->rows && a->cache_row_offset >= 0,
               "bad cache arguments");
  const int cp = a->head_perm_cp > 1 ? a->head_perm_cp : 1;
  if (cp > 1) {
    MAGI_CHECK(a->cache == nullptr, "cache append is not available together with head_perm_cp (cache holds post-all-to-all heads)");
    if (a->x_q) MAGI_CHECK(a->num_heads_q % cp == 0, "num_heads_q %% cp != 0");
    if (a->x_k || a->x_v)
      MAGI_CHECK(a->num_heads_kv % cp == 0 || cp % a->num_heads_kv == 0, "num_heads_kv and cp are not compatible");
  }
  MAGI_CHECK((a->sin == nullptr) == (a->cos == nullptr), "sin and cos must both be given or both be NULL");
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(a->x_q) | reinterpret_cast<uintptr_t>(a->x_k) |
               reinterpret_cast<uintptr_t>(a->x_v) | reinterpret_cast<uintptr_t>(a->q_out) |
               reinterpret_cast<uintptr_t>(a->kv_out) | reinterpret_cast<uintptr_t>(a->cache) |
               reinterpret_cast<uintptr_t>(a->sin) | reinterpret_cast<uintptr_t>(a->cos) |
               reinterpret_cast<uintptr_t>(a->ln_q_w) | reinterpret_cast<uintptr_t>(a->ln_q_b) |
               reinterpret_cast<uintptr_t>(a->ln_k_w) | reinterpret_cast<uintptr_t>(a->ln_k_b)) & 15) == 0,
             "all pointers must be 16-byte aligned");
  if (a->x_qx) {
    MAGI_CHECK(a->qx_out && a->ln_qx_w && a->ln_qx_b && a->num_heads_q > 0, "qx path needs qx_out, ln_qx_w, ln_qx_b");
    MAGI_CHECK(a->xqx_row_stride % 8 == 0 && a->xqx_head_stride % 8 == 0 && a->xqx_head_stride >= 0 &&
                   a->qx_out_token_stride % 8 == 0 && a->qx_out_head_stride % 8 == 0,
               "qx strides must be multiples of 8 elements");
  }
  if (a->x_tk || a->x_tv) {
    MAGI_CHECK(a->tkv_out && a->rows_t >= 0 && a->num_heads_kv > 0, "text k / v need tkv_out and rows_t >= 0");
    if (a->x_tk) MAGI_CHECK(a->ln_tk_w && a->ln_tk_b, "text k needs ln_tk_w, ln_tk_b");
    MAGI_CHECK(a->xt_row_stride % 8 == 0 && a->xt_head_stride % 8 == 0 && a->xt_head_stride >= 0 &&
                   a->tkv_out_token_stride % 8 == 0 && a->tkv_out_head_stride % 8 == 0,
               "text k / v strides must be multiples of 8 elements");
  }
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(a->x_qx) | reinterpret_cast<uintptr_t>(a->x_tk) |
               reinterpret_cast<uintptr_t>(a->x_tv) | reinterpret_cast<uintptr_t>(a->qx_out) |
               reinterpret_cast<uintptr_t>(a->tkv_out) | reinterpret_cast<uintptr_t>(a->ln_qx_w) |
               reinterpret_cast<uintptr_t>(a->ln_qx_b) | reinterpret_cast<uintptr_t>(a->ln_tk_w) |
               reinterpret_cast<uintptr_t>(a->ln_tk_b)) & 15) == 0,
             "all pointers must be 16-byte aligned");
  const bool main_rows = a->x_q || a->x_k || a->x_v || a->x_qx;
  const bool text_rows = (a->x_tk || a->x_tv) && a->rows_t > 0;
  if ((!main_rows || a->rows == 0) && !text_rows) return 0;

  magi::PrologueParams p;
  p.x_q = a->x_q; p.x_k = a->x_k; p.x_v = a->x_v;
  p.rows = a->rows;
  p.x_qx = a->x_qx; p.xqx_row_stride = a->xqx_row_stride;
  p.xqx_head_stride = a->xqx_head_stride > 0 ? a->xqx_head_stride : 128;
  p.ln_qx_w = a->ln_qx_w; p.ln_qx_b = a->ln_qx_b;
  p.qx_out = a->qx_out; p.qx_out_token_stride = a->qx_out_token_stride; p.qx_out_head_stride = a->qx_out_head_stride;
  p.x_tk = a->x_tk; p.x_tv = a->x_tv; p.rows_t = (a->x_tk || a->x_tv) ? a->rows_t : 0;
  p.xt_row_stride = a->xt_row_stride; p.xt_head_stride = a->xt_head_stride > 0 ? a->xt_head_stride : 128;
  p.ln_tk_w = a->ln_tk_w; p.ln_tk_b = a->ln_tk_b;
  p.tkv_out = a->tkv_out; p.tkv_out_token_stride = a->tkv_out_token_stride;
  p.tkv_out_head_stride = a->tkv_out_head_stride;
  if (a->rows == 0) { p.x_q = p.x_k = p.x_v = p.x_qx = nullptr; }
  if (p.rows_t == 0) { p.x_tk = p.x_tv = nullptr; }
  p.xq_row_stride = a->xq_row_stride; p.xk_row_stride = a->xk_row_stride; p.xv_row_stride = a->xv_row_stride;
  p.xq_head_stride = a->xq_head_stride > 0 ? a->xq_head_stride : 128;
  p.xk_head_stride = a->xk_head_stride > 0 ? a->xk_head_stride : 128;
  p.xv_head_stride = a->xv_head_stride > 0 ? a->xv_head_stride : 128;
  p.Hq = a->num_heads_q; p.Hkv = a->num_heads_kv; p.batch = a->batch;
  p.ln_q_w = a->ln_q_w; p.ln_q_b = a->ln_q_b; p.ln_k_w = a->ln_k_w; p.ln_k_b = a->ln_k_b;
  p.eps = a->eps; p.zero_centered_gamma = a->zero_centered_gamma;
  p.sin = a->sin; p.cos = a->cos;
  p.q_out = a->q_out; p.q_out_token_stride = a->q_out_token_stride; p.q_out_head_stride = a->q_out_head_stride;
  p.kv_out = a->kv_out; p.kv_out_token_stride = a->kv_out_token_stride; p.kv_out_head_stride = a->kv_out_head_stride;
  p.cache = a->cache; p.cache_token_stride = a->cache_token_stride; p.cache_head_stride = a->cache_head_stride;
  p.cache_row_offset = a->cache_row_offset; p.cache_rows = a->cache_rows;
  p.head_perm_cp = cp;
  p.use_peers = 0;
  p.peer_row_offset = 0;
  for (int i = 0; i < 8; ++i) { p.peer_q[i] = nullptr; p.peer_kv[i] = nullptr; }
  if (a->num_peers > 0) {
    MAGI_CHECK(a->num_peers == cp && cp <= MAGI_MAX_PEERS, "num_peers must equal head_perm_cp (at most %d)", MAGI_MAX_PEERS);
    MAGI_CHECK(a->peer_row_offset >= 0, "negative peer_row_offset");
    for (int i = 0; i < cp; ++i) {
      if (a->x_q) MAGI_CHECK(a->peer_q_out[i] && (reinterpret_cast<uintptr_t>(a->peer_q_out[i]) & 15) == 0, "bad peer_q_out[%d]", i);
      if (a->x_k || a->x_v)
        MAGI_CHECK(a->peer_kv_out[i] && (reinterpret_cast<uintptr_t>(a->peer_kv_out[i]) & 15) == 0, "bad peer_kv_out[%d]", i);
      p.peer_q[i] = a->peer_q_out[i];
      p.peer_kv[i] = a->peer_kv_out[i];
    }
    p.use_peers = 1;
    p.peer_row_offset = a->peer_row_offset;
  }
  p.gather_peers = 0;
  p.gather_multicast = 0;
  p.gather_rows_per_chunk = 1; p.gather_chunk_stride = 0; p.gather_row_offset = 0;
  if (a->gather_peers > 0) {
    MAGI_CHECK(a->gather_peers <= MAGI_MAX_PEERS, "gather_peers must be at most %d", MAGI_MAX_PEERS);
    MAGI_CHECK(cp == 1 && a->num_peers == 0 && a->cache == nullptr,
               "the K/V all-gather is not available together with head_perm_cp, num_peers or the cache append");
    MAGI_CHECK(a->gather_rows_per_chunk > 0 && a->rows % a->gather_rows_per_chunk == 0 && a->gather_row_offset >= 0 &&
                   a->gather_row_offset + a->gather_rows_per_chunk <= a->gather_chunk_stride,
               "bad gather geometry: rows must be whole stripes and a stripe must fit inside a chunk");
    if (a->x_k || a->x_v)
      for (int i = 0; i < a->gather_peers; ++i) {
        MAGI_CHECK(a->gather_kv_out[i] && (reinterpret_cast<uintptr_t>(a->gather_kv_out[i]) & 15) == 0,
                   "bad gather_kv_out[%d]", i);
        p.peer_kv[i] = a->gather_kv_out[i];
      }
    MAGI_CHECK(a->gather_multicast == 0 || a->gather_peers == 1, "gather_multicast takes exactly one (multicast) pointer");
    p.gather_peers = a->gather_peers;
    p.gather_multicast = a->gather_multicast != 0;
    p.gather_rows_per_chunk = a->gather_rows_per_chunk;
    p.gather_chunk_stride = a->gather_chunk_stride;
    p.gather_row_offset = a->gather_row_offset;
  }

  int64_t units = 0;
  magi::prologue_units(p, &units);
  MAGI_CHECK(units < (int64_t)0x7fffffff, "too many rows for one prologue launch");
  int sms = 0;
  if (int rc = device_sm_count(&sms)) return rc;
  MAGI_CUDA(magi::prologue_launch(p, a->ln_dtype == 1, sms, stream));
  g_launches++;
  return 0;
}

int magi_kv_cache_append(void* cache, int64_t cache_token_stride, const void* kv, int64_t kv_token_stride,
                         int64_t row_offset, int64_t rows, int64_t row_elems, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(rows >= 0 && row_offset >= 0, "negative rows/offset");
  if (rows == 0) return 0;
  MAGI_CHECK(cache && kv, "NULL pointer");
  MAGI_CHECK(row_elems > 0 && row_elems % 8 == 0 && cache_token_stride % 8 == 0 && kv_token_stride % 8 == 0,
             "row_elems and strides must be multiples of 8 elements");
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(cache) | reinterpret_cast<uintptr_t>(kv)) & 15) == 0,
             "pointers must be 16-byte aligned");
  int grid = 1;
  if (int rc = grid_for(rows * (row_elems / 8), 256, &grid)) return rc;
  MAGI_CUDA(magi::kv_append_launch(cache, cache_token_stride, kv, kv_token_stride, row_offset, rows,
                                   (int)(row_elems / 8), grid, stream));
  g_launches++;
  return 0;
}

int magi_lse_merge(void* out_a, float* lse_a, const void* out_b, const float* lse_b, int64_t rows, int32_t heads,
                   int64_t a_ts, int64_t a_hs, int64_t b_ts, int64_t b_hs, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(rows >= 0 && heads > 0, "bad shape");
  if (rows == 0) return 0;
  MAGI_CHECK(out_a && lse_a && out_b && lse_b, "NULL pointer");
  MAGI_CHECK(a_ts % 8 == 0 && a_hs % 8 == 0 && b_ts % 8 == 0 && b_hs % 8 == 0,
             "strides must be multiples of 8 elements");
  int grid = 1;
  if (int rc = grid_for(rows * heads * 16, 256, &grid)) return rc;
  MAGI_CUDA(magi::lse_merge_launch(out_a, lse_a, out_b, lse_b, rows, heads, a_ts, a_hs, b_ts, b_hs, grid, stream));
  g_launches++;
  return 0;
}

int magi_ulysses_relayout(void* dst, const void* src, int64_t seq, int32_t heads, int32_t head_elems, int32_t cp,
                          int32_t kv_repeat, int32_t mode, int32_t dst_heads_total, int32_t dst_head_offset,
                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(seq >= 0 && heads > 0 && cp >= 1 && kv_repeat >= 1, "bad shape");
  MAGI_CHECK(mode == 0 || mode == 1, "mode must be 0 or 1");
  MAGI_CHECK(head_elems > 0 && head_elems % 8 == 0, "head_elems must be a multiple of 8");
  if (seq == 0) return 0;
  MAGI_CHECK(dst && src, "NULL pointer");
  int hn;
  if (mode == 0) {
    MAGI_CHECK((heads * kv_repeat) % cp == 0, "heads * kv_repeat must be divisible by cp");
    hn = heads * kv_repeat / cp;
  } else {
    MAGI_CHECK(kv_repeat == 1, "kv_repeat only applies to mode 0");
    hn = heads;  // per-rank heads
  }
  if (dst_heads_total <= 0) dst_heads_total = hn;
  MAGI_CHECK(mode == 1 || (dst_heads_total == hn && dst_head_offset == 0), "dst head window only applies to mode 1");
  MAGI_CHECK(dst_head_offset >= 0 && dst_head_offset + hn <= dst_heads_total, "dst head window out of range");
  int grid = 1;
  if (int rc = grid_for(seq * cp * hn * (head_elems / 8), 256, &grid)) return rc;
  MAGI_CUDA(magi::ulysses_relayout_launch(dst, src, seq, hn, head_elems / 8, cp, kv_repeat, mode,
                                          dst_heads_total, dst_head_offset, grid, stream));
  g_launches++;
  return 0;
}

int magi_cso_unshuffle(void* dst, int64_t dst_token_stride, const void* src, int64_t src_token_stride, int32_t cp,
                       int32_t dn, int64_t spb, int64_t clip, int64_t row_elems, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(cp >= 1 && dn >= 1 && spb >= 0 && clip >= 0, "bad shape");
  MAGI_CHECK(clip <= spb * cp, "clip must not exceed cp * spb (the padded chunk length)");
  if (clip == 0) return 0;
  MAGI_CHECK(dst && src, "NULL pointer");
  MAGI_CHECK(row_elems > 0 && row_elems % 8 == 0 && dst_token_stride % 8 == 0 && src_token_stride % 8 == 0,
             "row_elems and strides must be multiples of 8 elements");
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0,
             "pointers must be 16-byte aligned");
  int grid = 1;
  if (int rc = grid_for(dn * clip * (row_elems / 8), 256, &grid)) return rc;
  MAGI_CUDA(magi::cso_unshuffle_launch(dst, dst_token_stride, src, src_token_stride, dn, spb, clip,
                                       (int)(row_elems / 8), grid, stream));
  g_launches++;
  return 0;
}

int magi_bias_modulate_add(const MagiGatingArgs* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(a != nullptr, "NULL args");
  MAGI_CHECK(a->rows >= 0 && a->hidden > 0, "bad shape");
  MAGI_CHECK(magi::gating_warps_per_row(a->hidden) > 0, "hidden must be a multiple of 256 (at most 12288)");
  MAGI_CHECK((a->x_dtype == 0 || a->x_dtype == 1) && (a->gate_dtype == 0 || a->gate_dtype == 1),
             "x_dtype / gate_dtype must be 0 (fp32) or 1 (bf16)");
  if (a->rows == 0) return 0;
  MAGI_CHECK(a->x && a->residual && a->cond_map && a->gate && a->ln_w && a->ln_b && a->out, "NULL pointer");
  MAGI_CHECK(a->x_row_stride % 8 == 0 && a->residual_row_stride % 8 == 0 && a->out_row_stride % 8 == 0 &&
                 a->gate_row_stride % 8 == 0,
             "row strides must be multiples of 8 elements");
  MAGI_CHECK(a->x_row_stride >= a->hidden && a->residual_row_stride >= a->hidden && a->out_row_stride >= a->hidden,
             "row strides smaller than hidden");
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->residual) |
               reinterpret_cast<uintptr_t>(a->gate) | reinterpret_cast<uintptr_t>(a->ln_w) |
               reinterpret_cast<uintptr_t>(a->ln_b) | reinterpret_cast<uintptr_t>(a->out)) & 15) == 0,
             "pointers must be 16-byte aligned");
  MAGI_CHECK(a->num_gates > 0 && a->cond_map_stride >= 1, "bad gate / cond_map description");
  magi::GatingParams p;
  p.x = a->x; p.residual = a->residual; p.cond_map = a->cond_map; p.gate = a->gate;
  p.ln_w = a->ln_w; p.ln_b = a->ln_b; p.out = a->out;
  p.rows = a->rows; p.hidden = a->hidden;
  p.x_row_stride = a->x_row_stride; p.res_row_stride = a->residual_row_stride; p.out_row_stride = a->out_row_stride;
  p.gate_row_stride = a->gate_row_stride; p.map_stride = a->cond_map_stride;
  p.map_is_i64 = a->cond_map_is_int64; p.eps = a->eps; p.zero_centered_gamma = a->zero_centered_gamma;
  int sms = 0;
  if (int rc = device_sm_count(&sms)) return rc;
  MAGI_CUDA(magi::gating_launch(p, a->x_dtype == 1, a->gate_dtype == 1, sms, stream));
  g_launches++;
  return 0;
}

namespace {
// (rows, cols) bf16 row-major matrix -> 2-D tensor map {cols, rows}, box {64, 128}, 128B swizzle (rows past the end are
// zero-filled by the TMA unit)
// 2-D map of a row-major GEMM operand, box = 128 rows x one 128-byte swizzle row (64 bf16 or 128 e4m3 bytes)
int make_matrix_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t row_stride, const char* what,
                    bool fp8 = false) {
  EncodeTiledFn enc = get_encode_fn();
  MAGI_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  const int esize = fp8 ? 1 : 2;
  MAGI_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (row_stride * esize) % 16 == 0 && row_stride >= cols,
             "%s: pointer must be 16-byte aligned and the row stride a multiple of 16 bytes", what);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(row_stride * esize)};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MAGI_CHECK(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (CUresult %d)", what, (int)r);
  return 0;
}
}  // namespace

int magi_qkv_proj_norm_rope_fwd(const magi_qkv_proj_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(a != nullptr, "args is NULL");
  MAGI_CHECK(a->rows >= 0 && a->rows < (1ll << 31), "rows out of range");
  const bool fp8 = a->fp8 != 0;
  MAGI_CHECK(a->hidden > 0 && a->hidden % (fp8 ? 128 : 64) == 0,
             "hidden must be a positive multiple of 64 (bf16) / 128 (fp8)");
  MAGI_CHECK(!fp8 || (a->out_scale != nullptr && (reinterpret_cast<uintptr_t>(a->out_scale) & 3) == 0),
             "fp8 needs out_scale (device pointer to 4 floats)");
  MAGI_CHECK(a->num_heads_q > 0 && a->num_heads_kv > 0, "num_heads_q / num_heads_kv must be positive");
  MAGI_CHECK(a->num_heads_qx == 0 || a->num_heads_qx == a->num_heads_q, "num_heads_qx must be 0 or num_heads_q");
  const int heads_total = a->num_heads_q + 2 * a->num_heads_kv + a->num_heads_qx;
  MAGI_CHECK(heads_total % 2 == 0, "the stacked weight must hold an even number of heads (256-column tiles)");
  MAGI_CHECK(a->ln_dtype == 0 || a->ln_dtype == 1, "ln_dtype must be 0 (fp32) or 1 (bf16)");
  MAGI_CHECK(a->batch >= 1, "batch must be >= 1");
  if (a->rows == 0) return 0;
  const int cp = a->num_peers > 1 ? a->num_peers : 0;
  MAGI_CHECK(a->num_peers >= 0 && a->num_peers <= MAGI_MAX_PEERS && a->num_peers != 1, "num_peers must be 0 or 2..8");
  MAGI_CHECK(a->x && a->w && a->ln_q_w && a->ln_q_b && a->ln_k_w && a->ln_k_b, "NULL pointer");
  MAGI_CHECK(cp > 0 || (a->q_out && a->kv_out), "q_out / kv_out are NULL");
  if (cp > 0) {
    MAGI_CHECK(a->cache == nullptr, "the cache append is not available together with peer routing");
    MAGI_CHECK(a->num_heads_q % cp == 0 && (a->num_heads_kv % cp == 0 || cp % a->num_heads_kv == 0),
               "head counts and num_peers are not compatible");
    MAGI_CHECK(a->peer_row_offset >= 0, "negative peer_row_offset");
    for (int i = 0; i < cp; ++i)
      MAGI_CHECK(a->peer_q_out[i] && a->peer_kv_out[i] &&
                     ((reinterpret_cast<uintptr_t>(a->peer_q_out[i]) | reinterpret_cast<uintptr_t>(a->peer_kv_out[i])) & 15) == 0,
                 "bad peer_q_out / peer_kv_out[%d]", i);
  }
  MAGI_CHECK(a->num_heads_qx == 0 || (a->qx_out && a->ln_qx_w && a->ln_qx_b), "qx path needs qx_out, ln_qx_w, ln_qx_b");
  MAGI_CHECK((a->sin == nullptr) == (a->cos == nullptr), "sin and cos must both be given or both be NULL");
  MAGI_CHECK(a->q_out_token_stride % 8 == 0 && a->q_out_head_stride % 8 == 0 && a->kv_out_token_stride % 8 == 0 &&
                 a->kv_out_head_stride % 8 == 0 && a->qx_out_token_stride % 8 == 0 && a->qx_out_head_stride % 8 == 0 &&
                 a->cache_token_stride % 8 == 0 && a->cache_head_stride % 8 == 0,
             "output strides must be multiples of 8 elements");
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(a->q_out) | reinterpret_cast<uintptr_t>(a->kv_out) |
               reinterpret_cast<uintptr_t>(a->qx_out) | reinterpret_cast<uintptr_t>(a->cache) |
               reinterpret_cast<uintptr_t>(a->sin) | reinterpret_cast<uintptr_t>(a->cos)) & 15) == 0,
             "output / table pointers must be 16-byte aligned");
  if (a->cache)
    MAGI_CHECK(a->cache_rows >= 0 && a->cache_rows <= a->rows && a->cache_row_offset >= 0, "bad cache arguments");
  CUtensorMap tm_x, tm_w;
  if (int rc = make_matrix_map(&tm_x, a->x, a->rows, a->hidden, a->x_row_stride, "x", fp8)) return rc;
  if (int rc = make_matrix_map(&tm_w, a->w, (int64_t)heads_total * 128, a->hidden, a->w_row_stride, "w", fp8)) return rc;
  magi::QkvGemmParams p;
  p.rows = a->rows; p.hidden = a->hidden; p.Hq = a->num_heads_q; p.Hkv = a->num_heads_kv; p.heads_total = heads_total;
  p.batch = a->batch;
  p.ln_q_w = a->ln_q_w; p.ln_q_b = a->ln_q_b; p.ln_k_w = a->ln_k_w; p.ln_k_b = a->ln_k_b;
  p.ln_qx_w = a->num_heads_qx ? a->ln_qx_w : nullptr; p.ln_qx_b = a->num_heads_qx ? a->ln_qx_b : nullptr;
  p.ln_bf16 = a->ln_dtype; p.eps = a->eps; p.zero_centered_gamma = a->zero_centered_gamma;
  p.sin = a->sin; p.cos = a->cos;
  p.q_out = a->q_out; p.q_out_token_stride = a->q_out_token_stride; p.q_out_head_stride = a->q_out_head_stride;
  p.kv_out = a->kv_out; p.kv_out_token_stride = a->kv_out_token_stride; p.kv_out_head_stride = a->kv_out_head_stride;
  p.qx_out = a->qx_out; p.qx_out_token_stride = a->qx_out_token_stride; p.qx_out_head_stride = a->qx_out_head_stride;
  p.cache = a->cache; p.cache_token_stride = a->cache_token_stride; p.cache_head_stride = a->cache_head_stride;
  p.cache_row_offset = a->cache_row_offset; p.cache_rows = a->cache ? a->cache_rows : 0;
  p.cp = cp;
  p.hq_loc = cp > 0 ? a->num_heads_q / cp : a->num_heads_q;
  p.hkv_loc = cp > 0 ? (a->num_heads_kv >= cp ? a->num_heads_kv / cp : 1) : a->num_heads_kv;
  p.peer_row_offset = a->peer_row_offset;
  p.fp8 = fp8 ? 1 : 0;
  p.out_scale = fp8 ? a->out_scale : nullptr;
  uintptr_t bits = reinterpret_cast<uintptr_t>(a->qx_out) | reinterpret_cast<uintptr_t>(a->cache);
  for (int i = 0; i < 8; ++i) {
    p.peer_q[i] = i < cp ? a->peer_q_out[i] : nullptr;
    p.peer_kv[i] = i < cp ? a->peer_kv_out[i] : nullptr;
    bits |= reinterpret_cast<uintptr_t>(p.peer_q[i]) | reinterpret_cast<uintptr_t>(p.peer_kv[i]);
  }
  if (cp == 0) bits |= reinterpret_cast<uintptr_t>(a->q_out) | reinterpret_cast<uintptr_t>(a->kv_out);
  p.align32 = (bits & 31) == 0 && a->qx_out_token_stride % 16 == 0 && a->qx_out_head_stride % 16 == 0 &&
              a->cache_token_stride % 16 == 0 && a->cache_head_stride % 16 == 0 &&
              (cp > 0 || (a->q_out_token_stride % 16 == 0 && a->q_out_head_stride % 16 == 0 &&
                          a->kv_out_token_stride % 16 == 0 && a->kv_out_head_stride % 16 == 0));
  int sms = a->num_sms;
  if (sms <= 0) { if (int rc = device_sm_count(&sms)) return rc; }
  MAGI_CUDA(magi::qkv_gemm_launch(tm_x, tm_w, p, sms, stream));
  g_launches++;
  return 0;
}

int magi_fp8_quantize(const void* x, int64_t x_row_stride, const float* scale, int32_t scale_len, void* out,
                      int64_t out_row_stride, int64_t rows, int32_t cols, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(rows >= 0 && cols > 0 && cols % 16 == 0, "rows must be >= 0 and cols a positive multiple of 16");
  MAGI_CHECK(scale_len == 1 || scale_len == cols, "scale_len must be 1 or cols");
  if (rows == 0) return 0;
  MAGI_CHECK(x && out && scale, "NULL pointer");
  MAGI_CHECK(x_row_stride % 8 == 0 && out_row_stride % 16 == 0 && x_row_stride >= cols && out_row_stride >= cols,
             "row strides must cover a row and keep 16-byte alignment (x: multiple of 8, out: multiple of 16)");
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(scale) & (scale_len == 1 ? 3 : 15)) == 0,
             "x / out (and a per-channel scale) must be 16-byte aligned");
  int sms = 0;
  if (int rc = device_sm_count(&sms)) return rc;
  MAGI_CUDA(magi::fp8_quant_launch(x, x_row_stride, scale, scale_len, out, out_row_stride, rows, cols, sms, stream));
  g_launches++;
  return 0;
}

int magi_mlp_activation(const void* x, int64_t x_row_stride, void* out, int64_t out_row_stride, int64_t rows, int32_t f,
                        int32_t gated, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(rows >= 0 && f > 0 && f % 8 == 0, "rows must be >= 0 and f a positive multiple of 8");
  if (rows == 0) return 0;
  MAGI_CHECK(x && out, "NULL pointer");
  MAGI_CHECK(x_row_stride % 8 == 0 && out_row_stride % 8 == 0 && out_row_stride >= f &&
                 x_row_stride >= (gated ? 2 : 1) * (int64_t)f,
             "row strides must be multiples of 8 elements and cover a row");
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "pointers must be 16-byte aligned");
  MAGI_CHECK(!(gated && x == out), "out may not alias x when gated");
  int sms = 0;
  if (int rc = device_sm_count(&sms)) return rc;
  MAGI_CUDA(magi::mlp_act_launch(x, x_row_stride, out, out_row_stride, rows, f, gated != 0, sms, stream));
  g_launches++;
  return 0;
}

int magi_apply_rotary_emb(const magi_rotary_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MAGI_CHECK(a != nullptr, "NULL args");
  MAGI_CHECK(a->dtype == 0 || a->dtype == 1, "dtype must be 0 (fp32) or 1 (bf16)");
  MAGI_CHECK(a->batch >= 0 && a->seqlen >= 0 && a->heads > 0 && a->head_dim > 0, "bad shape");
  MAGI_CHECK(a->rotary_dim >= 0 && a->rotary_dim <= a->head_dim, "rotary_dim must be <= head_dim");
  MAGI_CHECK(a->rotary_dim % 8 == 0 && (a->head_dim - a->rotary_dim) % 4 == 0,
             "rotary_dim / 2 and head_dim - rotary_dim must be multiples of 4");
  if (a->batch == 0 || a->seqlen == 0) return 0;
  MAGI_CHECK(a->x && a->out && (a->rotary_dim == 0 || (a->cos && a->sin)), "NULL pointer");
  const int64_t al = a->dtype == 1 ? 8 : 4;   // elements per 16 bytes
  MAGI_CHECK(a->x_batch_stride % al == 0 && a->x_seq_stride % al == 0 && a->x_head_stride % al == 0 &&
                 a->out_batch_stride % al == 0 && a->out_seq_stride % al == 0 && a->out_head_stride % al == 0,
             "strides must be multiples of 16 bytes");
  MAGI_CHECK(a->x_head_stride >= a->head_dim && a->out_head_stride >= a->head_dim, "head stride < head_dim");
  MAGI_CHECK(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->out) |
               reinterpret_cast<uintptr_t>(a->cos) | reinterpret_cast<uintptr_t>(a->sin)) & 15) == 0,
             "pointers must be 16-byte aligned");
  magi::RotaryParams p;
  p.x = a->x; p.out = a->out; p.batch = a->batch; p.seqlen = a->seqlen;
  p.heads = a->heads; p.head_dim = a->head_dim; p.rotary_dim = a->rotary_dim;
  p.x_bs = a->x_batch_stride; p.x_ss = a->x_seq_stride; p.x_hs = a->x_head_stride;
  p.o_bs = a->out_batch_stride; p.o_ss = a->out_seq_stride; p.o_hs = a->out_head_stride;
  p.cos = a->cos; p.sin = a->sin;
  const int64_t items = a->rotary_dim / 8 + (a->head_dim - a->rotary_dim) / 4;
  int grid = 1;
  if (int rc = grid_for(a->batch * a->seqlen * a->heads * items, 256, &grid)) return rc;
  MAGI_CUDA(magi::rotary_launch(p, a->dtype == 1, grid, stream));
  g_launches++;
  return 0;
}

}  // extern "C"
