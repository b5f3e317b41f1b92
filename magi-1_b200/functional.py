"""Drop-in attention functions with the call signatures the reference uses inside
inference/model/dit/dit_module.py, backed by libmagi_attn.so (hand-written sm_100a kernels) through ctypes.

    flex_flash_attn_func    <- magi_attention.functional.flex_flash_attn_func   (dit_module.py:1032, 1074, 1111)
    flash_attn_func         <- flash_attn.flash_attn_func                       (dit_module.py:1062, 1090)
    flash_attn_varlen_func  <- flash_attn.flash_attn_varlen_func                (dit_module.py:1124)
    attn_ranges             two-segment (KV cache ++ new tokens) form of the same kernel: what
                            _full_adjust_key_and_value + core_attention compute without the torch.cat / chunk copies
                            (dit_module.py:914-946, 1023-1066)
    qk_norm_rope            get_q / get_k / get_v / get_kv / get_xqkv fused      (dit_module.py:953-1021)
    kv_cache_append, lse_merge, ulysses_relayout, cso_unshuffle : helpers around the context-parallel exchange
    bias_modulate_add       <- bias_modulate_add / range_mod_triton / post_norm    (dit_module.py:201-307)
    apply_rotary_emb        <- flash_attn.layers.rotary.apply_rotary_emb (`flash_apply_rotary_emb`, dit_module.py:961,
                            978), for callers that keep the unfused get_q / get_k; qk_norm_rope fuses it away

No CPU fallback: every function requires CUDA tensors and the built extension.
"""
import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import MagiAttnArgs, MagiGatingArgs, MagiPrologueArgs, MagiQkvProjArgs, MagiRotaryArgs, MagiView

HEAD_DIM = 128


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_thd(x: torch.Tensor, name: str) -> None:
    if not x.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (this package has no CPU path)")
    if x.dtype != torch.bfloat16:
        raise ValueError(f"{name} must be bfloat16, got {x.dtype}")
    if x.dim() != 3 or x.shape[-1] != HEAD_DIM:
        raise ValueError(f"{name} must have shape (tokens, heads, 128), got {tuple(x.shape)}")
    if x.shape[0] > 0 and x.stride(-1) != 1:
        raise ValueError(f"{name} must be contiguous in the last dimension")


def _view(x: Optional[torch.Tensor]) -> MagiView:
    if x is None or x.shape[0] == 0:
        return MagiView(None, 0, 0, 0)
    return MagiView(x.data_ptr(), x.shape[0], x.stride(0), x.stride(1))


def _ranges_arg(r: torch.Tensor, name: str) -> torch.Tensor:
    if not r.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor")
    if r.dtype != torch.int32:
        r = r.to(torch.int32)
    return r.contiguous()


class AttnPlan:
    """The device-side work list of one (ranges, num_heads_q, token counts) combination, built by `attn_plan` and
    reusable by every `attn_ranges` call with the same combination - all layers of a denoise forward share their
    ranges (dit_model.py:215-259), so the list is built once per forward instead of once per layer."""
    __slots__ = ("ws", "n", "heads_q", "total_q", "total_k", "by_ranges", "_keep")

    def matches(self, n, heads_q, total_q, total_k, by_ranges) -> bool:
        return (self.n, self.heads_q, self.total_q, self.total_k, self.by_ranges) == (n, heads_q, total_q, total_k,
                                                                                      by_ranges)


def attn_plan(q_ranges: Optional[torch.Tensor] = None, k_ranges: Optional[torch.Tensor] = None, *,
              cu_seqlens_q: Optional[torch.Tensor] = None, cu_seqlens_k: Optional[torch.Tensor] = None,
              num_heads_q: int, total_q: int, total_k: int) -> AttnPlan:
    """magi_attn_plan: sort / clamp the ranges and write the work list on the device (no host reads). total_q = rows
    of q, total_k = rows of the key sequence (cache prefix + new tokens)."""
    by_ranges = q_ranges is not None and k_ranges is not None
    by_cu = cu_seqlens_q is not None and cu_seqlens_k is not None
    if by_ranges == by_cu:
        raise ValueError("give exactly one of (q_ranges, k_ranges) or (cu_seqlens_q, cu_seqlens_k)")
    if by_ranges:
        q_ranges, k_ranges = _ranges_arg(q_ranges, "q_ranges"), _ranges_arg(k_ranges, "k_ranges")
        n, dev, keep = q_ranges.shape[0], q_ranges.device, (q_ranges, k_ranges)
    else:
        cu_seqlens_q, cu_seqlens_k = _ranges_arg(cu_seqlens_q, "cu_seqlens_q"), _ranges_arg(cu_seqlens_k, "cu_seqlens_k")
        n, dev, keep = cu_seqlens_q.shape[0] - 1, cu_seqlens_q.device, (cu_seqlens_q, cu_seqlens_k)
    L = _lib.lib()
    plan = AttnPlan()
    plan.ws = torch.empty((L.magi_attn_workspace_bytes(max(n, 0)) + 3) // 4, dtype=torch.int32, device=dev)
    plan.n, plan.heads_q, plan.total_q, plan.total_k, plan.by_ranges = n, num_heads_q, total_q, total_k, by_ranges
    plan._keep = keep
    a = MagiAttnArgs()
    a.q = MagiView(None, total_q, 0, 0)
    a.k_a = MagiView(None, 0, 0, 0)
    a.k_b = MagiView(None, total_k, 0, 0)
    a.q_ranges = q_ranges.data_ptr() if by_ranges else None
    a.k_ranges = k_ranges.data_ptr() if by_ranges else None
    a.cu_seqlens_q = cu_seqlens_q.data_ptr() if by_cu else None
    a.cu_seqlens_k = cu_seqlens_k.data_ptr() if by_cu else None
    a.n_ranges, a.num_heads_q = n, num_heads_q
    a.workspace, a.workspace_bytes = plan.ws.data_ptr(), plan.ws.numel() * 4
    with torch.cuda.device(dev):
        _lib.check(L.magi_attn_plan(C.byref(a), _stream()), "magi_attn_plan")
    return plan


def attn_ranges(
    q: torch.Tensor,
    k_new: Optional[torch.Tensor],
    v_new: Optional[torch.Tensor],
    q_ranges: Optional[torch.Tensor] = None,
    k_ranges: Optional[torch.Tensor] = None,
    k_cache: Optional[torch.Tensor] = None,
    v_cache: Optional[torch.Tensor] = None,
    cu_seqlens_q: Optional[torch.Tensor] = None,
    cu_seqlens_k: Optional[torch.Tensor] = None,
    softmax_scale: Optional[float] = None,
    out: Optional[torch.Tensor] = None,
    return_lse: bool = True,
    zero_fill: bool = False,
    num_sms: int = 0,
    proj_input: Optional[Tuple[torch.Tensor, int, int]] = None,
    peer_out: Optional[Tuple[Sequence[int], Sequence[int], int, int]] = None,
    kernel: int = 0,
    plan: Optional[AttnPlan] = None,
    independent_of_prev: bool = False,
    peer_proj: Optional[Tuple[int, int]] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Range-masked attention over keys = concat(k_cache, k_new) without materialising the concatenation.

    independent_of_prev: the caller asserts that no tensor of this call (inputs, `out`, the ranges, the plan) is read
    or written by the kernel launched just before it on the stream (magi_attn_args.launch_dependency = 1): the launch
    then overlaps that kernel's tail instead of waiting for it, and still completes in stream order. Needs a `plan`
    and a caller-provided destination (`out` / `proj_input`) allocated BEFORE that earlier kernel was launched - a
    tensor allocated in between may reuse memory the earlier kernel is still reading.

    plan: an `attn_plan` result for the same ranges / head count / token counts (checked): skips the per-call work-list
    kernel and workspace allocation.

    kernel: 0 = library default, 1 = attn_fwd_kernel, 2 = attn_pair_kernel (experimental CTA-pair kernel).

    peer_out = (pointers, row_bounds, head_offset, heads_total[, token_stride]): fused Ulysses output all-to-all. Query row g of
    local head h is stored into the (rows_rho, heads_total, 128) bf16 buffer at pointers[rho] of the rank rho with
    row_bounds[rho] <= g < row_bounds[rho + 1], at row g - row_bounds[rho], head head_offset + h (NVLink peer
    memory; include/magi_attn.h). Returns (None, None); the caller synchronises the ranks.

    peer_proj = (n, hn), with peer_out: the peers' buffers are out-projection inputs (rows_rho, 2 * heads_total * 128)
    laid out "(hn n hd)" (see proj_input); the rows go straight into slot n of the owner's buffer.

    proj_input = (buffer, n, hn): instead of a (Tq, Hq, 128) tensor, write the result straight into the reference's
    out-projection input `buffer` (Tq, 2 * Hq * 128), laid out "(hn n hd)" with n = 0 for self-attention and n = 1
    for cross-attention and hn = 8 head groups (attn_linear_proj, dit_module.py:1332-1338). Returns (buffer, lse).

    q (Tq, Hq, 128), k_*/v_* (T*, Hkv, 128) bf16 views (any token/head strides that are multiples of 8: the packed
    `(T, Hkv, 256)` KV tensor of the reference is passed as `kv[..., :128]`, `kv[..., 128:]`).
    Ranges are int32 CUDA tensors of shape (n, 2) in the coordinate system of the concatenation
    (dit_module.py:914-931), or cu_seqlens of n+1 entries. Returns (out (Tq, Hq, 128) bf16, lse (Tq, Hq) fp32).
    """
    _check_thd(q, "q")
    Hq = q.shape[1]
    segs = []
    for k_, v_, nm in ((k_cache, v_cache, "cache"), (k_new, v_new, "new")):
        if (k_ is None) != (v_ is None):
            raise ValueError(f"k_{nm} and v_{nm} must both be given or both be None")
        if k_ is not None:
            _check_thd(k_, f"k_{nm}")
            _check_thd(v_, f"v_{nm}")
            if k_.shape != v_.shape:
                raise ValueError(f"k_{nm} and v_{nm} shapes differ")
        segs.append((k_, v_))
    ref = segs[1][0] if segs[1][0] is not None else segs[0][0]
    if ref is None:
        raise ValueError("no keys given")
    Hkv = ref.shape[1]
    for k_, _ in segs:
        if k_ is not None and k_.shape[1] != Hkv:
            raise ValueError("cache and new keys have different head counts")
    if Hq % Hkv != 0:
        raise ValueError(f"num_heads_q ({Hq}) must be a multiple of num_heads_kv ({Hkv})")
    by_ranges = q_ranges is not None and k_ranges is not None
    by_cu = cu_seqlens_q is not None and cu_seqlens_k is not None
    if by_ranges == by_cu:
        raise ValueError("give exactly one of (q_ranges, k_ranges) or (cu_seqlens_q, cu_seqlens_k)")
    if by_ranges:
        q_ranges = _ranges_arg(q_ranges, "q_ranges")
        k_ranges = _ranges_arg(k_ranges, "k_ranges")
        if q_ranges.shape != k_ranges.shape or q_ranges.dim() != 2 or q_ranges.shape[1] != 2:
            raise ValueError("q_ranges and k_ranges must both have shape (n, 2)")
        n = q_ranges.shape[0]
    else:
        cu_seqlens_q = _ranges_arg(cu_seqlens_q, "cu_seqlens_q")
        cu_seqlens_k = _ranges_arg(cu_seqlens_k, "cu_seqlens_k")
        if cu_seqlens_q.shape != cu_seqlens_k.shape or cu_seqlens_q.dim() != 1:
            raise ValueError("cu_seqlens_q and cu_seqlens_k must be 1-D and of equal length")
        n = cu_seqlens_q.shape[0] - 1

    Tq = q.shape[0]
    group_heads = group_stride = 0
    if proj_input is not None:
        buf, n_src, hn = proj_input
        if out is not None:
            raise ValueError("give either out or proj_input")
        if (not buf.is_cuda or buf.dtype != torch.bfloat16 or buf.dim() != 2 or buf.shape != (Tq, 2 * Hq * HEAD_DIM)
                or buf.stride(1) != 1 or Hq % hn != 0 or n_src not in (0, 1)):
            raise ValueError(f"proj_input buffer must be a CUDA bf16 tensor ({Tq}, {2 * Hq * HEAD_DIM}); hn must divide Hq")
        group_heads = Hq // hn
        group_stride = 2 * group_heads * HEAD_DIM
        out_ptr = buf.data_ptr() + n_src * group_heads * HEAD_DIM * 2
        out_ts, out_hs = buf.stride(0), HEAD_DIM
    elif peer_out is not None:
        if out is not None or return_lse:
            raise ValueError("peer_out excludes out and return_lse")
        ptrs, bounds, head_off, heads_total = peer_out[:4]
        if len(ptrs) > _lib.MAX_PEERS or len(bounds) != len(ptrs) + 1:
            raise ValueError("peer_out needs at most 8 pointers and one more row bound than pointers")
        # optional 5th entry: token stride of the peers' buffers in elements (batched rows "(sq b)": b * heads_total * 128,
        # with the pointers already advanced to the batch row)
        out_ptr, out_ts, out_hs = None, (peer_out[4] if len(peer_out) > 4 else heads_total * HEAD_DIM), HEAD_DIM
        if peer_proj is not None:
            n_src, hn = peer_proj
            if heads_total % hn != 0 or n_src not in (0, 1):
                raise ValueError("peer_proj: hn must divide heads_total and n must be 0 or 1")
            group_heads = heads_total // hn
            group_stride = 2 * group_heads * HEAD_DIM
            out_ts = 2 * heads_total * HEAD_DIM
            ptrs = [int(p_) + n_src * group_heads * HEAD_DIM * 2 for p_ in ptrs]
    elif out is None:
        # zero_fill: the kernel writes the uncovered rows itself (fill_uncovered_rows); only a call that launches
        # nothing (no ranges) needs the tensors cleared here
        out = (torch.zeros if zero_fill and (n <= 0 or kernel == 2) else torch.empty)(
            (Tq, Hq, HEAD_DIM), dtype=torch.bfloat16, device=q.device)
    else:
        _check_thd(out, "out")
        if out.shape != q.shape:
            raise ValueError("out must have the shape of q")
    if proj_input is not None:
        out = proj_input[0]
    elif peer_out is None:
        out_ptr, out_ts, out_hs = out.data_ptr(), out.stride(0), out.stride(1)
    lse = None
    if return_lse:
        lse = (torch.full((Tq, Hq), float("-inf"), dtype=torch.float32, device=q.device)
               if zero_fill and (n <= 0 or kernel == 2) else torch.empty((Tq, Hq), dtype=torch.float32, device=q.device))
    if n <= 0 or Tq == 0:
        return out, lse

    L = _lib.lib()
    if plan is not None:
        total_k = sum(s_[0].shape[0] for s_ in segs if s_[0] is not None)
        given = (q_ranges, k_ranges) if by_ranges else (cu_seqlens_q, cu_seqlens_k)
        if not plan.matches(n, Hq, Tq, total_k, by_ranges) or any(
                a_.data_ptr() != b_.data_ptr() for a_, b_ in zip(given, plan._keep)):
            raise ValueError("plan was built for other range tensors / heads / token counts")
        ws = plan.ws
    else:
        ws_bytes = L.magi_attn_workspace_bytes(n)
        ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.int32, device=q.device)
    a = MagiAttnArgs()
    a.plan_ready = 1 if plan is not None else 0
    a.q = _view(q)
    a.k_a, a.v_a = _view(segs[0][0]), _view(segs[0][1])
    a.k_b, a.v_b = _view(segs[1][0]), _view(segs[1][1])
    a.out = out_ptr
    a.out_token_stride, a.out_head_stride = out_ts, out_hs
    a.out_group_heads, a.out_group_stride = group_heads, group_stride
    a.lse = lse.data_ptr() if lse is not None else None
    a.q_ranges = q_ranges.data_ptr() if by_ranges else None
    a.k_ranges = k_ranges.data_ptr() if by_ranges else None
    a.cu_seqlens_q = cu_seqlens_q.data_ptr() if by_cu else None
    a.cu_seqlens_k = cu_seqlens_k.data_ptr() if by_cu else None
    a.n_ranges = n
    a.num_heads_q, a.num_heads_kv, a.head_dim = Hq, Hkv, HEAD_DIM
    a.softmax_scale = float(softmax_scale) if softmax_scale is not None else -1.0
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    a.num_sms = num_sms
    a.kernel = kernel
    if zero_fill and proj_input is None and peer_out is None and kernel != 2:   # (the pair kernel gets a pre-filled out)
        a.fill_uncovered_rows = 1
    if independent_of_prev:
        if plan is None or return_lse or peer_out is not None or (proj_input is None and zero_fill):
            raise ValueError("independent_of_prev needs a plan and a destination allocated by the caller (no lse)")
        a.launch_dependency = 1
    if peer_out is not None:
        a.num_peers, a.peer_head_offset = len(ptrs), head_off
        for i, ptr in enumerate(ptrs):
            a.peer_out[i] = ptr
        for i, b_ in enumerate(bounds):
            a.peer_row_bounds[i] = int(b_)
    with torch.cuda.device(q.device):
        _lib.check(L.magi_attn_fwd(C.byref(a), _stream()), "magi_attn_fwd")
    return out, lse


def flex_flash_attn_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    q_ranges: torch.Tensor,
    k_ranges: torch.Tensor,
    max_seqlen_q: Optional[int] = None,
    max_seqlen_k: Optional[int] = None,
    softmax_scale: Optional[float] = None,
    deterministic: bool = False,
    disable_fwd_atomic_reduction: bool = True,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same call as MagiAttention's `flex_flash_attn_func` at dit_module.py:1032-1043.

    Query ranges must be disjoint (the only mode the reference uses: disable_fwd_atomic_reduction=True); the kernel is
    deterministic by construction (no atomics, fixed summation order), so `deterministic` is accepted and ignored.
    Rows covered by no range come back as zeros with lse = -inf.
    """
    if not disable_fwd_atomic_reduction:
        raise ValueError("overlapping q_ranges (disable_fwd_atomic_reduction=False) are not supported")
    out, lse = attn_ranges(q, k, v, q_ranges, k_ranges, softmax_scale=softmax_scale, zero_fill=True)
    return out, lse


def flash_attn_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    dropout_p: float = 0.0,
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    deterministic: bool = False,
) -> torch.Tensor:
    """Same call as `flash_attn.flash_attn_func` at dit_module.py:1062/1090: (b, s, h, 128) dense, non-causal, GQA."""
    if dropout_p != 0.0 or causal:
        raise ValueError("only dropout_p=0, causal=False are supported (what the reference uses)")
    if q.dim() != 4:
        raise ValueError("q must be (batch, seqlen, heads, 128)")
    b, sq = q.shape[0], q.shape[1]
    sk = k.shape[1]
    qf = q.reshape(b * sq, q.shape[2], q.shape[3])
    kf = k.reshape(b * sk, k.shape[2], k.shape[3])
    vf = v.reshape(b * sk, v.shape[2], v.shape[3])
    cu_q = torch.arange(0, (b + 1) * sq, sq, dtype=torch.int32, device=q.device) if sq > 0 else torch.zeros(
        b + 1, dtype=torch.int32, device=q.device)
    cu_k = torch.arange(0, (b + 1) * sk, sk, dtype=torch.int32, device=q.device) if sk > 0 else torch.zeros(
        b + 1, dtype=torch.int32, device=q.device)
    out, _ = attn_ranges(qf, kf, vf, cu_seqlens_q=cu_q, cu_seqlens_k=cu_k, softmax_scale=softmax_scale,
                         return_lse=False)
    return out.view(b, sq, q.shape[2], q.shape[3])


def flash_attn_varlen_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    cu_seqlens_q: torch.Tensor,
    cu_seqlens_k: torch.Tensor,
    max_seqlen_q: Optional[int] = None,
    max_seqlen_k: Optional[int] = None,
    dropout_p: float = 0.0,
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    deterministic: bool = False,
) -> torch.Tensor:
    """Same call as `flash_attn.flash_attn_varlen_func` at dit_module.py:1124-1133 (cross-attention)."""
    if dropout_p != 0.0 or causal:
        raise ValueError("only dropout_p=0, causal=False are supported (what the reference uses)")
    out, _ = attn_ranges(q, k, v, cu_seqlens_q=cu_seqlens_q, cu_seqlens_k=cu_seqlens_k, softmax_scale=softmax_scale,
                         return_lse=False, zero_fill=True)
    return out


def _ptr(t: Optional[torch.Tensor]):
    return t.data_ptr() if t is not None else None


def qk_norm_rope(
    x_q: Optional[torch.Tensor],
    x_k: Optional[torch.Tensor],
    x_v: Optional[torch.Tensor],
    num_heads_q: int,
    num_heads_kv: int,
    ln_q: Optional[Tuple[torch.Tensor, torch.Tensor]],
    ln_k: Optional[Tuple[torch.Tensor, torch.Tensor]],
    sin: Optional[torch.Tensor],
    cos: Optional[torch.Tensor],
    batch: int = 1,
    eps: float = 1e-6,
    zero_centered_gamma: bool = True,
    q_out: Optional[torch.Tensor] = None,
    kv_out: Optional[torch.Tensor] = None,
    cache: Optional[torch.Tensor] = None,
    cache_row_offset: int = 0,
    cache_rows: int = 0,
    head_perm_cp: int = 1,
    peer_q_out: Optional[Sequence[int]] = None,
    peer_kv_out: Optional[Sequence[int]] = None,
    peer_row_offset: int = 0,
    peer_heads: Optional[Tuple[int, int]] = None,
    gather_kv_out: Optional[Sequence[int]] = None,
    gather_geometry: Optional[Tuple[int, int, int]] = None,
    gather_multicast: bool = False,
    cross: Optional[dict] = None,
) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Fused per-head LayerNorm(w+1, b) -> NeoX RoPE on 96 of 128 dims -> bf16, writing q in attention layout and
    [K|V] packed (rows, Hkv, 256), optionally also appending to the KV cache (dit_module.py:953-996, 918-929).

    cross = dict(x_qx=(rows, Hq*128) or (rows, Hq, 128) view, ln_qx=(w, b) bf16, text_kv=(rows_t, Hkv, 256) packed
    text projection, ln_tk=(w, b) bf16[, qx_out=, tkv_out=]): the cross-attention's half of the layer (get_xqkv,
    dit_module.py:1005-1021: two bf16-parameter LayerNorms, no RoPE) rides in the SAME launch - one prologue launch per
    layer instead of three. The results are put back into the dict under "qx" (rows, Hq, 128) and "tkv"
    (rows_t, Hkv, 256); they are always written in place, whatever the routing of q / k / v. Either half may be absent.

    x_* are the projection outputs, (rows, H*128) or strided (rows, H, 128) bf16 views (rows = "(sq b)" flattened;
    strides multiples of 8). ln_* = (weight, bias), both fp32 (self-attention, dit_model.py:621-633) or both bf16 (cross-attention).
    sin/cos: (rows // batch, 48) fp32, or None to skip RoPE (dit_module.py:1005-1021).
    With head_perm_cp = cp > 1 the outputs are laid out "(cp seq) hn hd", i.e. ready to be the all-to-all send buffer
    (context_parallel.py:393-396), including KV-head replication when cp > Hkv.
    With peer_q_out / peer_kv_out (head_perm_cp device pointers each) the all-to-all itself is fused: head slice c is
    stored straight into rank c's receive buffer ((all rows, peer_heads[0], 128) resp. (all rows, peer_heads[1], 256))
    at row peer_row_offset + r over NVLink peer memory (include/magi_attn.h); nothing is returned for that tensor.
    With gather_kv_out (device pointers of every rank's gathered in-flight K/V buffer, (chunks * chunk_stride, Hkv, 256)
    packed, this rank's own included - or ONE NVSwitch multicast pointer) and gather_geometry = (rows_per_chunk,
    chunk_stride, row_offset) the K/V all-gather of cp_ring_kv is fused: local row (chunk j, stripe row t) is stored at
    row j * chunk_stride + row_offset + t of every buffer; q is written locally as usual, no kv tensor is returned.
    """
    x_qx = cross.get("x_qx") if cross else None
    text_kv = cross.get("text_kv") if cross else None
    ref = next(t for t in (x_q, x_k, x_v, x_qx, text_kv) if t is not None)
    dev = ref.device
    rows = next((t.shape[0] for t in (x_q, x_k, x_v, x_qx) if t is not None), 0)
    for t, nm, H in ((x_q, "x_q", num_heads_q), (x_k, "x_k", num_heads_kv), (x_v, "x_v", num_heads_kv)):
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.bfloat16:
            raise ValueError(f"{nm} must be a CUDA bfloat16 tensor")
        ok2 = t.dim() == 2 and t.shape == (rows, H * HEAD_DIM) and t.stride(1) == 1
        ok3 = t.dim() == 3 and t.shape == (rows, H, HEAD_DIM) and t.stride(2) == 1
        if not (ok2 or ok3):
            raise ValueError(f"{nm} must be (rows, {H}*128) or (rows, {H}, 128) with a contiguous last dim, "
                             f"got {tuple(t.shape)}")
    cp = max(1, head_perm_cp)
    ln_dtype = None
    for ln, nm in ((ln_q, "ln_q"), (ln_k, "ln_k")):
        if ln is None:
            continue
        w, b = ln
        if w.dtype != b.dtype or w.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError(f"{nm} weight/bias must both be fp32 or both bf16")
        if ln_dtype is not None and ln_dtype != w.dtype:
            raise ValueError("ln_q and ln_k must have the same dtype")
        ln_dtype = w.dtype
        if w.numel() != HEAD_DIM or b.numel() != HEAD_DIM or not w.is_contiguous() or not b.is_contiguous():
            raise ValueError(f"{nm} weight/bias must be contiguous with 128 elements")
    if (sin is None) != (cos is None):
        raise ValueError("sin and cos must both be given or both be None")
    if sin is not None:
        if sin.dtype != torch.float32 or cos.dtype != torch.float32:
            raise ValueError("sin/cos must be fp32")
        if rows % batch != 0 or sin.shape != (rows // batch, 48) or cos.shape != sin.shape:
            raise ValueError(f"sin/cos must have shape ({rows // max(batch, 1)}, 48), got {tuple(sin.shape)}")
        sin, cos = sin.contiguous(), cos.contiguous()

    use_peers = peer_q_out is not None or peer_kv_out is not None
    if use_peers:
        if cp <= 1 or peer_heads is None:
            raise ValueError("peer routing needs head_perm_cp > 1 and peer_heads")
        if (x_q is not None and (peer_q_out is None or len(peer_q_out) != cp)) or (
                (x_k is not None or x_v is not None) and (peer_kv_out is None or len(peer_kv_out) != cp)):
            raise ValueError("peer pointer lists must have head_perm_cp entries")
    gather = gather_kv_out is not None
    if gather:
        if use_peers or cp > 1 or (cache is not None and cache_rows > 0) or gather_geometry is None:
            raise ValueError("gather_kv_out needs gather_geometry and excludes head_perm_cp, peer routing and the cache "
                             "append")
        if not 1 <= len(gather_kv_out) <= _lib.MAX_PEERS:
            raise ValueError(f"gather_kv_out must hold 1..{_lib.MAX_PEERS} pointers")
        rpc, cstride, roff = gather_geometry
        if rpc <= 0 or rows % rpc != 0 or roff < 0 or roff + rpc > cstride:
            raise ValueError("bad gather_geometry: rows must be whole stripes and a stripe must fit inside a chunk")
    if x_q is not None and q_out is None and not use_peers:
        q_out = torch.empty((rows * cp, num_heads_q // cp, HEAD_DIM), dtype=torch.bfloat16, device=dev)
    if (x_k is not None or x_v is not None) and kv_out is None and not use_peers and not gather:
        hn_kv = max(1, num_heads_kv // cp) if cp > 1 else num_heads_kv
        kv_out = torch.empty((rows * cp, hn_kv, 2 * HEAD_DIM), dtype=torch.bfloat16, device=dev)

    a = MagiPrologueArgs()
    if x_qx is not None:
        if not x_qx.is_cuda or x_qx.dtype != torch.bfloat16 or not (
                (x_qx.dim() == 2 and x_qx.shape == (rows, num_heads_q * HEAD_DIM) and x_qx.stride(1) == 1) or
                (x_qx.dim() == 3 and x_qx.shape == (rows, num_heads_q, HEAD_DIM) and x_qx.stride(2) == 1)):
            raise ValueError(f"cross['x_qx'] must be a CUDA bf16 (rows, {num_heads_q}*128) or (rows, {num_heads_q}, 128)")
        w, b = cross["ln_qx"]
        if w.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or w.numel() != HEAD_DIM or b.numel() != HEAD_DIM:
            raise ValueError("cross['ln_qx'] must be bf16 with 128 elements")
        qx_out = cross.get("qx_out")
        if qx_out is None:
            qx_out = torch.empty((rows, num_heads_q, HEAD_DIM), dtype=torch.bfloat16, device=dev)
        a.x_qx, a.xqx_row_stride = x_qx.data_ptr(), x_qx.stride(0)
        a.xqx_head_stride = x_qx.stride(1) if x_qx.dim() == 3 else HEAD_DIM
        a.ln_qx_w, a.ln_qx_b = w.data_ptr(), b.data_ptr()
        a.qx_out, a.qx_out_token_stride, a.qx_out_head_stride = qx_out.data_ptr(), qx_out.stride(0), qx_out.stride(1)
        cross["qx"] = qx_out
    if text_kv is not None:
        if (not text_kv.is_cuda or text_kv.dtype != torch.bfloat16 or text_kv.dim() != 3 or
                text_kv.shape[1:] != (num_heads_kv, 2 * HEAD_DIM) or text_kv.stride(2) != 1):
            raise ValueError(f"cross['text_kv'] must be a CUDA bf16 (rows_t, {num_heads_kv}, 256) tensor")
        w, b = cross["ln_tk"]
        if w.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or w.numel() != HEAD_DIM or b.numel() != HEAD_DIM:
            raise ValueError("cross['ln_tk'] must be bf16 with 128 elements")
        tkv_out = cross.get("tkv_out")
        if tkv_out is None:
            tkv_out = torch.empty((text_kv.shape[0], num_heads_kv, 2 * HEAD_DIM), dtype=torch.bfloat16, device=dev)
        a.x_tk, a.x_tv = text_kv.data_ptr(), text_kv.data_ptr() + HEAD_DIM * 2
        a.rows_t, a.xt_row_stride, a.xt_head_stride = text_kv.shape[0], text_kv.stride(0), text_kv.stride(1)
        a.ln_tk_w, a.ln_tk_b = w.data_ptr(), b.data_ptr()
        a.tkv_out, a.tkv_out_token_stride, a.tkv_out_head_stride = tkv_out.data_ptr(), tkv_out.stride(0), tkv_out.stride(1)
        cross["tkv"] = tkv_out
    a.x_q, a.x_k, a.x_v = _ptr(x_q), _ptr(x_k), _ptr(x_v)
    a.rows = rows
    a.xq_row_stride = x_q.stride(0) if x_q is not None else 0
    a.xk_row_stride = x_k.stride(0) if x_k is not None else 0
    a.xv_row_stride = x_v.stride(0) if x_v is not None else 0
    a.xq_head_stride = x_q.stride(1) if x_q is not None and x_q.dim() == 3 else HEAD_DIM
    a.xk_head_stride = x_k.stride(1) if x_k is not None and x_k.dim() == 3 else HEAD_DIM
    a.xv_head_stride = x_v.stride(1) if x_v is not None and x_v.dim() == 3 else HEAD_DIM
    a.num_heads_q, a.num_heads_kv, a.batch = num_heads_q, num_heads_kv, batch
    a.ln_q_w, a.ln_q_b = (_ptr(ln_q[0]), _ptr(ln_q[1])) if ln_q is not None else (None, None)
    a.ln_k_w, a.ln_k_b = (_ptr(ln_k[0]), _ptr(ln_k[1])) if ln_k is not None else (None, None)
    a.ln_dtype = 1 if ln_dtype == torch.bfloat16 else 0
    a.eps = eps
    a.zero_centered_gamma = 1 if zero_centered_gamma else 0
    a.sin, a.cos = _ptr(sin), _ptr(cos)
    if q_out is not None:
        a.q_out = q_out.data_ptr()
        a.q_out_token_stride, a.q_out_head_stride = q_out.stride(0), q_out.stride(1)
    if kv_out is not None:
        a.kv_out = kv_out.data_ptr()
        a.kv_out_token_stride, a.kv_out_head_stride = kv_out.stride(0), kv_out.stride(1)
    if cache is not None and cache_rows > 0:
        if cache.dtype != torch.bfloat16 or cache.dim() != 3 or cache.shape[2] != 2 * HEAD_DIM or cache.stride(2) != 1:
            raise ValueError("cache must be (max_tokens, Hkv, 256) bf16")
        if cache_row_offset + cache_rows > cache.shape[0]:
            raise ValueError("KV-cache append out of bounds")  # dit_module.py:927
        a.cache = cache.data_ptr()
        a.cache_token_stride, a.cache_head_stride = cache.stride(0), cache.stride(1)
        a.cache_row_offset, a.cache_rows = cache_row_offset, cache_rows
    a.head_perm_cp = cp
    if use_peers:
        a.num_peers, a.peer_row_offset = cp, peer_row_offset
        a.q_out_token_stride, a.q_out_head_stride = peer_heads[0] * HEAD_DIM, HEAD_DIM
        a.kv_out_token_stride, a.kv_out_head_stride = peer_heads[1] * 2 * HEAD_DIM, 2 * HEAD_DIM
        for i in range(cp):
            a.peer_q_out[i] = peer_q_out[i] if peer_q_out is not None else None
            a.peer_kv_out[i] = peer_kv_out[i] if peer_kv_out is not None else None
    if gather:
        a.gather_peers = len(gather_kv_out)
        if gather_multicast and len(gather_kv_out) != 1:
            raise ValueError("gather_multicast takes exactly one pointer: the multicast mapping of the buffer")
        a.gather_multicast = 1 if gather_multicast else 0
        a.gather_rows_per_chunk, a.gather_chunk_stride, a.gather_row_offset = gather_geometry
        a.kv_out_token_stride, a.kv_out_head_stride = num_heads_kv * 2 * HEAD_DIM, 2 * HEAD_DIM
        for i, ptr in enumerate(gather_kv_out):
            a.gather_kv_out[i] = ptr
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().magi_qk_norm_rope_fwd(C.byref(a), _stream()), "magi_qk_norm_rope_fwd")
    return q_out, kv_out


def kv_cache_append(cache: torch.Tensor, kv: torch.Tensor, row_offset: int, rows: Optional[int] = None) -> None:
    """cache[row_offset : row_offset + rows] = kv[:rows] for packed (tokens, Hkv, 256) bf16 tensors
    (dit_module.py:929, used when the new KV comes out of the Ulysses all-to-all)."""
    rows = kv.shape[0] if rows is None else rows
    if rows == 0:
        return
    if cache.dtype != torch.bfloat16 or kv.dtype != torch.bfloat16 or not cache.is_cuda or not kv.is_cuda:
        raise ValueError("cache and kv must be CUDA bfloat16 tensors")
    if cache.shape[1:] != kv.shape[1:] or not kv[0].is_contiguous() or not cache[0].is_contiguous():
        raise ValueError("cache and kv rows must have the same contiguous shape")
    if row_offset + rows > cache.shape[0]:
        raise ValueError("KV-cache append out of bounds")
    row_elems = kv[0].numel()
    with torch.cuda.device(cache.device):
        _lib.check(_lib.lib().magi_kv_cache_append(cache.data_ptr(), cache.stride(0), kv.data_ptr(), kv.stride(0),
                                                   row_offset, rows, row_elems, _stream()), "magi_kv_cache_append")


def cso_unshuffle(recv: torch.Tensor, cp: int, dn: int, clip: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """"(cp dn spb) hn nhd -> dn (cp spb) hn nhd"[:, :clip] flattened (dit_module.py:1215-1224) as one gather kernel;
    `out` (dn * clip rows, any token stride) may be the KV-cache slot."""
    if not recv.is_cuda or recv.dtype != torch.bfloat16 or recv.dim() != 3 or not recv.is_contiguous():
        raise ValueError("recv must be a contiguous CUDA bfloat16 tensor (tokens, heads, head_elems)")
    if recv.shape[0] % (cp * dn) != 0:
        raise ValueError("recv rows must be divisible by cp * dn")
    spb = recv.shape[0] // (cp * dn)
    if clip > cp * spb:
        raise ValueError("clip exceeds the padded chunk length")
    if out is None:
        out = torch.empty((dn * clip,) + tuple(recv.shape[1:]), dtype=recv.dtype, device=recv.device)
    elif out.shape != (dn * clip,) + tuple(recv.shape[1:]) or out.dtype != recv.dtype or (
            out.shape[0] > 0 and not out[0].is_contiguous()):
        raise ValueError("out must be (dn * clip, heads, head_elems) bf16 with contiguous rows")
    if out.shape[0] == 0:
        return out
    with torch.cuda.device(recv.device):
        _lib.check(_lib.lib().magi_cso_unshuffle(out.data_ptr(), out.stride(0), recv.data_ptr(), recv.stride(0), cp, dn,
                                                 spb, clip, recv[0].numel(), _stream()), "magi_cso_unshuffle")
    return out


def lse_merge(out_a: torch.Tensor, lse_a: torch.Tensor, out_b: torch.Tensor, lse_b: torch.Tensor) -> None:
    """In place: (out_a, lse_a) <- softmax-merge of two partial attention results over disjoint key sets."""
    _check_thd(out_a, "out_a")
    _check_thd(out_b, "out_b")
    if out_a.shape != out_b.shape or lse_a.shape != out_a.shape[:2] or lse_b.shape != lse_a.shape:
        raise ValueError("shape mismatch")
    if lse_a.dtype != torch.float32 or lse_b.dtype != torch.float32 or not lse_a.is_contiguous() or not lse_b.is_contiguous():
        raise ValueError("lse tensors must be contiguous fp32")
    with torch.cuda.device(out_a.device):
        _lib.check(_lib.lib().magi_lse_merge(out_a.data_ptr(), lse_a.data_ptr(), out_b.data_ptr(), lse_b.data_ptr(),
                                             out_a.shape[0], out_a.shape[1], out_a.stride(0), out_a.stride(1),
                                             out_b.stride(0), out_b.stride(1), _stream()), "magi_lse_merge")


def ulysses_relayout(src: torch.Tensor, cp: int, mode: int, kv_repeat: int = 1, out: Optional[torch.Tensor] = None,
                     dst_heads_total: int = 0, dst_head_offset: int = 0) -> torch.Tensor:
    """mode 0: "seq (cp hn) hd -> (cp seq) hn hd" (with optional head replication); mode 1: the inverse
    "(cp seq) hn hd -> seq (cp hn) hd" (context_parallel.py:393-396, 596). In mode 1 `out` may be a wider
    (seq, cp * dst_heads_total, hd) tensor of which heads [dst_head_offset, +hn) of every cp block are filled."""
    if not src.is_cuda or src.dtype != torch.bfloat16 or src.dim() != 3 or not src.is_contiguous():
        raise ValueError("src must be a contiguous CUDA bfloat16 tensor (tokens, heads, head_elems)")
    T, H, E = src.shape
    if mode == 0:
        hn = H * kv_repeat // cp
        shape = (T * cp, hn, E)
        seq, heads = T, H
    else:
        if T % cp != 0:
            raise ValueError("token count must be divisible by cp")
        total = dst_heads_total if dst_heads_total > 0 else H
        shape = (T // cp, total * cp, E)
        seq, heads = T // cp, H
    if out is None:
        out = torch.empty(shape, dtype=src.dtype, device=src.device)
    elif tuple(out.shape) != shape or not out.is_contiguous():
        raise ValueError(f"out must be contiguous with shape {shape}")
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().magi_ulysses_relayout(out.data_ptr(), src.data_ptr(), seq, heads, E, cp, kv_repeat, mode,
                                                    dst_heads_total, dst_head_offset, _stream()),
                   "magi_ulysses_relayout")
    return out


def bias_modulate_add(x: torch.Tensor, residual: torch.Tensor, condition_map: torch.Tensor, gate: torch.Tensor,
                      post_norm, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Same call as the reference's `bias_modulate_add(x, residual, condition_map, gate, post_norm)`
    (dit_module.py:292-307), one kernel instead of seven:
        out = (post_norm(x.float() * gate.float()[condition_map]) + residual.float()).to(residual.dtype)

    x (s, b, h) fp32 (the out-projection runs under fp32 autocast, dit_module.py:1343) or bf16 (MLP output);
    residual (s, b, h) bf16; condition_map (s, b) int64/int32 indices into gate.flatten(0, 1); gate (b, dn, h) fp32 or
    bf16; post_norm: a module with fp32 `weight`, `bias`, and `eps`, `zero_centered_gamma` attributes (FusedLayerNorm,
    dit_module.py:323-357) or a tuple (weight, bias, eps, zero_centered_gamma)."""
    if isinstance(post_norm, (tuple, list)):
        w, b, eps, zcg = post_norm
    else:
        w, b, eps, zcg = post_norm.weight, post_norm.bias, post_norm.eps, post_norm.zero_centered_gamma
    if not x.is_cuda:
        raise ValueError("x must be a CUDA tensor (this package has no CPU path)")
    if x.dtype not in (torch.float32, torch.bfloat16) or gate.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("x and gate must be fp32 or bf16")
    if residual.dtype != torch.bfloat16 or residual.shape != x.shape:
        raise ValueError("residual must be bf16 with the shape of x")
    if w.dtype != torch.float32 or b.dtype != torch.float32:
        raise ValueError("post_norm parameters must be fp32 (dit_model.py:621-638)")
    h = x.shape[-1]
    assert gate.shape[-1] == h   # dit_module.py:295
    if w.numel() != h or b.numel() != h:
        raise ValueError("post_norm parameters do not match the hidden size")
    if condition_map.dtype not in (torch.int64, torch.int32) or condition_map.numel() != x.numel() // h:
        raise ValueError("condition_map must be int32/int64 with one entry per row of x")
    x2, r2 = x.reshape(-1, h), residual.reshape(-1, h)          # "(sq b)" rows; views for contiguous inputs
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    if r2.stride(1) != 1:
        r2 = r2.contiguous()
    cm = condition_map.reshape(-1)
    g2 = gate.reshape(-1, h)
    if g2.stride(1) != 1:
        g2 = g2.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    elif out.shape != x.shape or out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise ValueError("out must be a contiguous bf16 tensor with the shape of x")
    a = MagiGatingArgs()
    a.x, a.x_dtype, a.x_row_stride = x2.data_ptr(), 1 if x2.dtype == torch.bfloat16 else 0, x2.stride(0)
    a.residual, a.residual_row_stride = r2.data_ptr(), r2.stride(0)
    a.cond_map, a.cond_map_is_int64, a.cond_map_stride = cm.data_ptr(), 1 if cm.dtype == torch.int64 else 0, cm.stride(0)
    a.gate, a.gate_dtype, a.gate_row_stride, a.num_gates = (g2.data_ptr(), 1 if g2.dtype == torch.bfloat16 else 0,
                                                            g2.stride(0), g2.shape[0])
    a.ln_w, a.ln_b = w.contiguous().data_ptr(), b.contiguous().data_ptr()
    a.eps, a.zero_centered_gamma = float(eps), 1 if zcg else 0
    a.out, a.out_row_stride = out.data_ptr(), h
    a.rows, a.hidden = x2.shape[0], h
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().magi_bias_modulate_add(C.byref(a), _stream()), "magi_bias_modulate_add")
    return out


def apply_rotary_emb(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, interleaved: bool = False,
                     inplace: bool = False) -> torch.Tensor:
    """Same call as `flash_attn.layers.rotary.apply_rotary_emb(x, cos, sin)` as the reference uses it
    (`flash_apply_rotary_emb(q, cos_emb, sin_emb)`, dit_module.py:961, 978): x (batch, seqlen, heads, head_dim) fp32 or
    bf16 - any batch / sequence / head strides, so the reference's transposed (b, sq, hn, hd) view needs no copy -
    cos / sin (seqlen, rotary_dim / 2); non-interleaved (NeoX) rotation of the first rotary_dim dims in fp32, the rest
    copied; returns a new tensor in x's dtype (or x itself with inplace=True)."""
    if interleaved:
        raise NotImplementedError("the reference only uses the non-interleaved (NeoX) form")
    if not x.is_cuda:
        raise ValueError("x must be a CUDA tensor (this package has no CPU path)")
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError(f"x must be fp32 or bf16, got {x.dtype}")
    if x.dim() != 4:
        raise ValueError(f"x must have shape (batch, seqlen, heads, head_dim), got {tuple(x.shape)}")
    b, s, h, d = x.shape
    if cos.shape != sin.shape or cos.dim() != 2 or cos.shape[0] < s:
        raise ValueError("cos / sin must both be (seqlen, rotary_dim / 2)")
    ro = 2 * cos.shape[1]
    if ro > d:
        raise ValueError("rotary_dim must be <= head_dim")     # flash-attn asserts the same
    if x.numel() > 0 and x.stride(-1) != 1:
        x = x.contiguous()
        inplace = False
    cos = cos.to(device=x.device, dtype=torch.float32).contiguous()
    sin = sin.to(device=x.device, dtype=torch.float32).contiguous()
    out = x if inplace else torch.empty((b, s, h, d), dtype=x.dtype, device=x.device)
    a = MagiRotaryArgs()
    a.x, a.out, a.dtype = x.data_ptr(), out.data_ptr(), 1 if x.dtype == torch.bfloat16 else 0
    a.batch, a.seqlen, a.heads, a.head_dim, a.rotary_dim = b, s, h, d, ro
    a.x_batch_stride, a.x_seq_stride, a.x_head_stride = x.stride(0), x.stride(1), x.stride(2)
    a.out_batch_stride, a.out_seq_stride, a.out_head_stride = out.stride(0), out.stride(1), out.stride(2)
    a.cos, a.sin = cos.data_ptr(), sin.data_ptr()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().magi_apply_rotary_emb(C.byref(a), _stream()), "magi_apply_rotary_emb")
    return out


def mlp_activation(x: torch.Tensor, gated: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The activation between fc1 and fc2 of CustomMLP (dit_module.py:541-549) as ONE memory-bound launch:
    gated -> `flashinfer.activation.silu_and_mul(x)` = silu(x[..., :f]) * x[..., f:] with a single rounding to bf16
    (the 24B model's SwiGLU); otherwise `torch.nn.functional.gelu(x)` (erf form, the 4.5B model).
    x: (..., 2 f) or (..., f) bf16 CUDA tensor, contiguous in the last dimension; returns (..., f) bf16."""
    if not x.is_cuda or x.dtype != torch.bfloat16:
        raise ValueError("x must be a bfloat16 CUDA tensor (this package has no CPU path)")
    width = x.shape[-1]
    f = width // 2 if gated else width
    if (gated and width % 2) or f % 8:
        raise ValueError("the activation width must be a multiple of 8")
    x2 = x.reshape(-1, width)
    if x2.numel() > 0 and x2.stride(1) != 1:
        x2 = x2.contiguous()
    rows = x2.shape[0]
    if out is None:
        out = torch.empty(x.shape[:-1] + (f,), dtype=torch.bfloat16, device=x.device)
    o2 = out.view(-1, f)
    if o2.shape[0] != rows or o2.stride(1) != 1:
        raise ValueError("out must be a (..., f) bf16 tensor contiguous in the last dimension")
    if rows > 0:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().magi_mlp_activation(x2.data_ptr(), x2.stride(0), o2.data_ptr(), o2.stride(0), rows, f,
                                                      1 if gated else 0, _stream()), "magi_mlp_activation")
    return out


def qkv_proj_norm_rope(
    x: torch.Tensor,
    w: torch.Tensor,
    num_heads_q: int,
    num_heads_kv: int,
    ln_q: Tuple[torch.Tensor, torch.Tensor],
    ln_k: Tuple[torch.Tensor, torch.Tensor],
    sin: Optional[torch.Tensor],
    cos: Optional[torch.Tensor],
    ln_qx: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
    batch: int = 1,
    eps: float = 1e-6,
    zero_centered_gamma: bool = True,
    cache: Optional[torch.Tensor] = None,
    cache_row_offset: int = 0,
    cache_rows: int = 0,
    num_sms: int = 0,
    peer_q_out: Optional[Sequence[int]] = None,
    peer_kv_out: Optional[Sequence[int]] = None,
    peer_row_offset: int = 0,
    out_scale: Optional[torch.Tensor] = None,
):
    """The q / k / v (and, with ln_qx, the cross-attention qx) projections of CustomLayerNormLinear
    (dit_module.py:405-425) and the QK-LayerNorm + RoPE prologue (get_q / get_kv, :953-996; q half of get_xqkv,
    :1005-1009) as ONE tensor-core GEMM whose epilogue is the prologue (magi_qkv_proj_norm_rope_fwd).

    x: (rows, hidden) bf16, the LayerNorm'ed hidden states in "(sq b)" row order. w: the stacked weight
    [Wq; Wk; Wv (; Wqx)] of shape ((Hq + 2 Hkv (+ Hq)) * 128, hidden), see `stack_qkv_weights`.
    Returns (q (rows, Hq, 128), kv (rows, Hkv, 256) packed [K | V], qx (rows, Hq, 128) or None).

    peer_q_out / peer_kv_out (cp device pointers each) + peer_row_offset: the Ulysses input all-to-all fused into the
    GEMM's epilogue, with the routing of `qk_norm_rope`'s peer arguments (head slice c -> rank c's receive buffer, at
    row peer_row_offset + r, KV-head replication when cp > Hkv); returns (None, None, qx).

    out_scale (4 fp32 on the device: input_scale * weight_scale of q, k, v, qx) selects the fp8 instantiation
    (PerTensorQuantizedFp8Linear, dit_module.py:431-457): x is then the e4m3 output of `fp8_quantize` and w the stacked
    e4m3 weights (dtype float8_e4m3fn or uint8), hidden a multiple of 128."""
    fp8 = out_scale is not None
    ok_dt = (torch.float8_e4m3fn, torch.uint8) if fp8 else (torch.bfloat16,)
    if not x.is_cuda or x.dtype not in ok_dt or w.dtype not in ok_dt or x.dim() != 2 or w.dim() != 2:
        raise ValueError("x and w must be 2-D bfloat16 (or, with out_scale, float8_e4m3fn) CUDA tensors "
                         "(this package has no CPU path)")
    if fp8 and (out_scale.dtype != torch.float32 or out_scale.numel() != 4 or not out_scale.is_cuda
                or not out_scale.is_contiguous()):
        raise ValueError("out_scale must be 4 contiguous fp32 values on the device")
    rows, hidden = x.shape
    with_qx = ln_qx is not None
    heads_total = num_heads_q + 2 * num_heads_kv + (num_heads_q if with_qx else 0)
    if w.shape != (heads_total * HEAD_DIM, hidden) or x.stride(1) != 1 or w.stride(1) != 1:
        raise ValueError(f"w must be ({heads_total * HEAD_DIM}, {hidden}) and both operands contiguous in the last dim")
    if hidden % (128 if fp8 else 64) or heads_total % 2:
        raise ValueError("hidden must be a multiple of 64 (fp8: 128) and the stacked weight must hold an even number "
                         "of heads")
    ln_dtype = 1 if ln_q[0].dtype == torch.bfloat16 else 0
    for t in (*ln_q, *ln_k):
        if t.dtype != ln_q[0].dtype or t.numel() != HEAD_DIM or not t.is_cuda:
            raise ValueError("q / k LayerNorm parameters must be 128-element CUDA tensors of one dtype")
    dev = x.device
    peers = peer_q_out is not None
    if peers and (peer_kv_out is None or len(peer_q_out) != len(peer_kv_out) or not 2 <= len(peer_q_out) <= _lib.MAX_PEERS
                  or (cache is not None and cache_rows > 0)):
        raise ValueError("peer routing needs 2..8 pointers for q and for kv and excludes the cache append")
    q = None if peers else torch.empty((rows, num_heads_q, HEAD_DIM), dtype=torch.bfloat16, device=dev)
    kv = None if peers else torch.empty((rows, num_heads_kv, 2 * HEAD_DIM), dtype=torch.bfloat16, device=dev)
    qx = torch.empty((rows, num_heads_q, HEAD_DIM), dtype=torch.bfloat16, device=dev) if with_qx else None
    a = MagiQkvProjArgs()
    a.x, a.rows, a.x_row_stride, a.hidden = x.data_ptr(), rows, x.stride(0), hidden
    a.w, a.w_row_stride = w.data_ptr(), w.stride(0)
    a.num_heads_q, a.num_heads_kv, a.num_heads_qx, a.batch = num_heads_q, num_heads_kv, num_heads_q if with_qx else 0, batch
    keep = [t.contiguous() for t in (*ln_q, *ln_k)]
    a.ln_q_w, a.ln_q_b, a.ln_k_w, a.ln_k_b = (t.data_ptr() for t in keep)
    a.ln_dtype = ln_dtype
    if with_qx:
        kx = [t.to(torch.bfloat16).contiguous() for t in ln_qx]
        keep += kx
        a.ln_qx_w, a.ln_qx_b = kx[0].data_ptr(), kx[1].data_ptr()
    a.eps, a.zero_centered_gamma = eps, 1 if zero_centered_gamma else 0
    if sin is not None:
        sin, cos = sin.contiguous().float(), cos.contiguous().float()
        if sin.shape != (rows // batch, 48) or cos.shape != sin.shape:
            raise ValueError("sin / cos must be (rows / batch, 48)")
        a.sin, a.cos = sin.data_ptr(), cos.data_ptr()
    if peers:
        a.num_peers, a.peer_row_offset = len(peer_q_out), peer_row_offset
        for i, (pq, pk) in enumerate(zip(peer_q_out, peer_kv_out)):
            a.peer_q_out[i], a.peer_kv_out[i] = pq, pk
    else:
        a.q_out, a.q_out_token_stride, a.q_out_head_stride = q.data_ptr(), q.stride(0), q.stride(1)
        a.kv_out, a.kv_out_token_stride, a.kv_out_head_stride = kv.data_ptr(), kv.stride(0), kv.stride(1)
    if with_qx:
        a.qx_out, a.qx_out_token_stride, a.qx_out_head_stride = qx.data_ptr(), qx.stride(0), qx.stride(1)
    if cache is not None and cache_rows > 0:
        if cache.dtype != torch.bfloat16 or cache.shape[1:] != (num_heads_kv, 2 * HEAD_DIM) or cache.stride(2) != 1:
            raise ValueError("cache must be (rows, Hkv, 256) bf16")
        a.cache, a.cache_token_stride, a.cache_head_stride = cache.data_ptr(), cache.stride(0), cache.stride(1)
        a.cache_row_offset, a.cache_rows = cache_row_offset, cache_rows
    a.num_sms = num_sms
    if fp8:
        a.fp8, a.out_scale = 1, out_scale.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().magi_qkv_proj_norm_rope_fwd(C.byref(a), _stream()), "magi_qkv_proj_norm_rope_fwd")
    return q, kv, qx


def stack_qkv_weights(wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor, wqx: Optional[torch.Tensor] = None):
    """[Wq; Wk; Wv (; Wqx)]: the row order qkv_proj_norm_rope expects (nn.Linear weights, (out, in)). float8 weights
    are stacked as bytes."""
    ws = [t for t in (wq, wk, wv, wqx) if t is not None]
    if ws[0].dtype == torch.float8_e4m3fn:
        ws = [t.view(torch.uint8) for t in ws]
    return torch.cat(ws, dim=0).contiguous()


FP8_MAX = 448.0   # torch.finfo(torch.float8_e4m3fn).max


def fp8_quantize(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """`div_clamp_to` (dit_module.py:364-384) as one launch of magi_fp8_quantize:
    float8_e4m3fn(bfloat16(clamp(x.float() / scale.float(), -448, 448))), same shape as x. `scale` holds one value
    or one per element of the last dimension (PerTensorQuantizedFp8Linear.input_scale: (in_features,),
    PerChannelQuantizedFp8Linear.smooth_scale: (1, in_features)). Byte-identical to the reference's eager chain."""
    if not x.is_cuda or x.dtype != torch.bfloat16:
        raise ValueError("x must be a bfloat16 CUDA tensor (this package has no CPU path)")
    cols = x.shape[-1]
    if scale.numel() not in (1, cols) or cols % 16:
        raise ValueError("scale must hold 1 or x.shape[-1] values and x.shape[-1] must be a multiple of 16")
    x2 = x.reshape(-1, cols)
    if x2.stride(1) != 1 or x2.stride(0) % 8:
        x2 = x2.contiguous()
    sc = scale.detach().reshape(-1)
    if sc.dtype != torch.float32 or not sc.is_contiguous() or sc.device != x.device:
        sc = sc.to(device=x.device, dtype=torch.float32).contiguous()
    out = torch.empty(x.shape, dtype=torch.float8_e4m3fn, device=x.device)
    rows = x2.shape[0]
    if rows > 0:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().magi_fp8_quantize(x2.data_ptr(), x2.stride(0), sc.data_ptr(), sc.numel(),
                                                    out.data_ptr(), cols, rows, cols, _stream()), "magi_fp8_quantize")
    return out


def fp8_scaled_linear(xq: torch.Tensor, weight: torch.Tensor, input_scale: torch.Tensor, weight_scale: torch.Tensor,
                      out_dtype=torch.bfloat16) -> torch.Tensor:
    """The GEMM of the reference's fp8 linears: `flashinfer.bmm_fp8(xq (1, m, k), weight.transpose(-2, -1), input_scale,
    weight_scale, dtype)` (dit_module.py:450-457, 482-488) = the cuBLASLt fp8 matmul with per-tensor A / B scale
    pointers, out = dtype((xq @ weight^T) * input_scale[0] * weight_scale[0]). A library GEMM in the reference and a
    library GEMM here (torch._scaled_mm -> the same cuBLASLt entry); cuBLASLt reads ONE float behind each scale
    pointer, which is why a vector input_scale contributes its first element.
    xq (..., k) float8_e4m3fn, weight (n, k) float8_e4m3fn."""
    k = xq.shape[-1]
    y = torch._scaled_mm(xq.reshape(-1, k), weight.t(), scale_a=input_scale.detach().reshape(-1)[0],
                         scale_b=weight_scale.detach().reshape(-1)[0], out_dtype=out_dtype)
    return y.view(*xq.shape[:-1], weight.shape[0])
