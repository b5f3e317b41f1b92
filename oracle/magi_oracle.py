"""ORACLE - test infrastructure only. CPU (numpy, fp32) restatement of the reference's algorithm for the hot path.

Nothing in the product (`magi-1_b200/`) imports this module. Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may use it, and only as the checker / reported baseline.

Reference: samiazirar/MAGI-1 (file:line relative to /root/reference). Third-party arithmetic that is NOT under
/root/reference and is restated here from its published definition:
  * flash-attn (pinned `flash-attn==2.4.2`, requirements.txt:20): `flash_attn_func`, `flash_attn_varlen_func` =
    softmax(Q K^T / sqrt(d)) V per (batch|segment, head) with GQA head mapping h -> h // (Hq/Hkv);
    `flash_attn.layers.rotary.apply_rotary_emb` = non-interleaved rotation (its own `apply_rotary_emb_torch`).
  * MagiAttention `flex_flash_attn_func` (unpinned, Dockerfile:17-19): for each i, rows q_ranges[i] attend densely
    to k_ranges[i]; with disjoint q ranges this equals the reference's own fallback loop, dit_module.py:1046-1065.

Parity pinning: the reference ships no tests or golden vectors for this path, so this oracle is
pinned against outputs of the reference's own Python code executed in the authoring container
(oracle/gen_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py).
"""
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------- bf16 helpers
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and return as fp32 (what `.to(torch.bfloat16)` does)."""
    x = np.ascontiguousarray(x, dtype=F32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(F32).copy()
    nan = np.isnan(x)
    out[nan] = np.nan
    return out


# ----------------------------------------------------------------------------------------------- prologue
def layer_norm(x: np.ndarray, weight: np.ndarray, bias: np.ndarray, eps: float = 1e-6,
               zero_centered_gamma: bool = True, weight_is_bf16: bool = False) -> np.ndarray:
    """FusedLayerNorm.forward, dit_module.py:355-357: F.layer_norm(x, (128,), weight + 1, bias, eps), biased variance.
    With bf16 parameters (the `_xattn` norms, dit_model.py:626-627) `weight + 1` is itself rounded to bf16."""
    x = x.astype(F32)
    w = weight.astype(F32)
    if zero_centered_gamma:
        w = w + F32(1.0)
        if weight_is_bf16:
            w = bf16_round(w)
    mean = x.mean(axis=-1, keepdims=True, dtype=F32)
    d = x - mean
    var = (d * d).mean(axis=-1, keepdims=True, dtype=F32)
    rstd = F32(1.0) / np.sqrt(var + F32(eps))
    return (d * rstd * w + bias.astype(F32)).astype(F32)


def apply_rotary_neox(x: np.ndarray, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    """flash_attn.layers.rotary.apply_rotary_emb (non-interleaved), called at dit_module.py:961, 978.
    x (..., S, H, D), cos/sin (S, R/2); rotates dims [0, R/2) with [R/2, R); dims >= R pass through."""
    half = cos.shape[-1]
    x1, x2, rest = x[..., :half], x[..., half:2 * half], x[..., 2 * half:]
    c = cos[:, None, :].astype(F32)
    s = sin[:, None, :].astype(F32)
    o1 = x1 * c - x2 * s
    o2 = x2 * c + x1 * s
    return np.concatenate([o1, o2, rest], axis=-1).astype(F32)


def get_q_or_k(x_lin: np.ndarray, num_heads: int, ln_w: np.ndarray, ln_b: np.ndarray, sin: np.ndarray,
               cos: np.ndarray, eps: float = 1e-6, zero_centered_gamma: bool = True) -> np.ndarray:
    """get_q / get_k, dit_module.py:953-980. x_lin: (sq, b, H*128) holding bf16 values (projection output).
    fp32 LayerNorm -> transpose to (b, sq, H, 128) -> fp32 RoPE -> one rounding to bf16 -> "(sq b) hn hd"."""
    sq, b, _ = x_lin.shape
    x = x_lin.reshape(sq, b, num_heads, 128).astype(F32)
    x = layer_norm(x, ln_w, ln_b, eps, zero_centered_gamma)
    x = np.transpose(x, (1, 0, 2, 3))  # b sq hn hd
    x = apply_rotary_neox(x, cos, sin)
    x = bf16_round(x)
    return np.ascontiguousarray(np.transpose(x, (1, 0, 2, 3))).reshape(sq * b, num_heads, 128)


def get_v(x_lin: np.ndarray, num_heads: int) -> np.ndarray:
    """get_v, dit_module.py:987-989: pure re-layout "sq b (hn hd) -> (sq b) hn hd"."""
    sq, b, _ = x_lin.shape
    return x_lin.reshape(sq * b, num_heads, 128).astype(F32)


def get_kv(xk_lin, xv_lin, num_heads_kv, ln_w, ln_b, sin, cos, eps=1e-6, zero_centered_gamma=True) -> np.ndarray:
    """get_kv, dit_module.py:991-996: cat([k, v], -1) -> (sq*b, Hkv, 256)."""
    k = get_q_or_k(xk_lin, num_heads_kv, ln_w, ln_b, sin, cos, eps, zero_centered_gamma)
    v = get_v(xv_lin, num_heads_kv)
    return np.concatenate([k, v], axis=-1)


def xattn_layer_norm(x: np.ndarray, ln_w_bf16: np.ndarray, ln_b_bf16: np.ndarray, eps: float = 1e-6,
                     zero_centered_gamma: bool = True) -> np.ndarray:
    """q_layernorm_xattn / k_layernorm_xattn inside get_xqkv, dit_module.py:1008, 1020: bf16 tensor, bf16 parameters,
    ATen computes in fp32 and rounds the result to bf16. No RoPE on the cross-attention path."""
    return bf16_round(layer_norm(x, ln_w_bf16, ln_b_bf16, eps, zero_centered_gamma, weight_is_bf16=True))


# ----------------------------------------------------------------------------------------------- KV cache
def full_adjust_key_and_value(cache: np.ndarray, key_and_value: np.ndarray, slice_point: int, clip_token_nums: int,
                              max_batch_size: int, update_kv_cache: bool,
                              distill_nearly_clean_chunk: bool) -> np.ndarray:
    """_full_adjust_key_and_value, dit_module.py:891-931. `cache` (max_seq*max_bs, Hkv, 256) is updated IN PLACE.
    Returns cat(cache[:sequence_start], key_and_value)."""
    sequence_start = slice_point * clip_token_nums * max_batch_size
    prefix = cache[:sequence_start].copy()
    if update_kv_cache:
        total = key_and_value.shape[0]
        clip_size = total - clip_token_nums * max_batch_size if distill_nearly_clean_chunk else total
        sequence_end = sequence_start + clip_size
        assert sequence_end <= cache.shape[0]  # dit_module.py:927
        cache[sequence_start:sequence_end] = key_and_value[:clip_size]
    return np.concatenate([prefix, key_and_value], axis=0)


def adjust_key_and_value_for_inference(key_and_value, cache, has_inference_params, slice_point, clip_token_nums,
                                       max_batch_size, update_kv_cache, extract_prefix_video_feature,
                                       fwd_extra_1st_chunk, distill_nearly_clean_chunk):
    """adjust_key_and_value_for_inference, dit_module.py:933-946. Returns (key, value)."""
    if has_inference_params and (extract_prefix_video_feature or fwd_extra_1st_chunk or slice_point > 0):
        key_and_value = full_adjust_key_and_value(cache, key_and_value, slice_point, clip_token_nums, max_batch_size,
                                                  update_kv_cache, distill_nearly_clean_chunk)
    return key_and_value[..., :128].copy(), key_and_value[..., 128:].copy()


# ----------------------------------------------------------------------------------------------- attention
def _attend(q: np.ndarray, k: np.ndarray, v: np.ndarray, scale: float) -> Tuple[np.ndarray, np.ndarray]:
    """softmax(q k^T * scale) v for q (Lq, Hq, D), k/v (Lk, Hkv, D) with GQA; fp32. Returns (out, lse)."""
    Lq, Hq, D = q.shape
    Lk, Hkv, _ = k.shape
    g = Hq // Hkv
    out = np.zeros((Lq, Hq, D), dtype=F32)
    lse = np.full((Lq, Hq), -np.inf, dtype=F32)
    if Lk == 0 or Lq == 0:
        return out, lse
    for h in range(Hq):
        s = (q[:, h, :].astype(F32) @ k[:, h // g, :].astype(F32).T) * F32(scale)
        m = s.max(axis=-1, keepdims=True)
        p = np.exp(s - m)
        l = p.sum(axis=-1, keepdims=True, dtype=F32)
        out[:, h, :] = (p / l) @ v[:, h // g, :].astype(F32)
        lse[:, h] = (m + np.log(l))[:, 0]
    return out, lse


def attn_ranges(q, k, v, q_ranges, k_ranges, softmax_scale: Optional[float] = None,
                round_output: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Range-masked attention = the reference's fallback loop, dit_module.py:1052-1065 (bs == 1), which is also the
    contract of flex_flash_attn_func with disjoint q ranges (dit_module.py:1032-1043).
    q (Tq, Hq, 128), k/v (Tk, Hkv, 128); ranges (n, 2). Uncovered rows: out 0, lse -inf."""
    D = q.shape[-1]
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    out = np.zeros(q.shape, dtype=F32)
    lse = np.full(q.shape[:2], -np.inf, dtype=F32)
    for (q0, q1), (k0, k1) in zip(np.asarray(q_ranges).reshape(-1, 2), np.asarray(k_ranges).reshape(-1, 2)):
        q0, q1, k0, k1 = int(q0), int(q1), int(k0), int(k1)
        if q1 <= q0:
            continue
        o, l = _attend(q[q0:q1], k[k0:k1], v[k0:k1], scale)
        out[q0:q1] = o
        lse[q0:q1] = l
    return (bf16_round(out) if round_output else out), lse


def core_attention(query, key, value, bs: int, np_q_range, np_k_range, denoising_range_num: int) -> np.ndarray:
    """core_attention, dit_module.py:1023-1066 (fallback branch). query (sq*b, Hq, 128) in "(sq b)" order.
    Returns (sq*b, Hq, 128) in "(sq b)" order, bf16-rounded."""
    assert not (bs > 1 and denoising_range_num > 1)  # dit_module.py:1048
    Hq = query.shape[1]
    qb = np.transpose(query.reshape(-1, bs, Hq, 128), (1, 0, 2, 3))               # b sq hn hd
    kb = np.transpose(key.reshape(-1, bs, key.shape[1], 128), (1, 0, 2, 3))
    vb = np.transpose(value.reshape(-1, bs, value.shape[1], 128), (1, 0, 2, 3))
    outs = []
    for i in range(denoising_range_num):
        j = i if bs == 1 else 0
        q0, q1 = int(np_q_range[j, 0]), int(np_q_range[j, 1])
        k0, k1 = int(np_k_range[j, 0]), int(np_k_range[j, 1])
        per_b = []
        for b in range(bs):
            o, _ = _attend(qb[b, q0:q1], kb[b, k0:k1], vb[b, k0:k1], 1.0 / math.sqrt(128))
            per_b.append(bf16_round(o))
        o = np.stack(per_b, axis=0)                                               # b sq h d
        outs.append(np.transpose(o, (1, 0, 2, 3)).reshape(-1, Hq, 128))           # (sq b) h d
    return np.concatenate(outs, axis=0)


def attn_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, softmax_scale: Optional[float] = None) -> np.ndarray:
    """flash_attn_varlen_func as used by cross_attention, dit_module.py:1124-1133: segment i of q attends to segment
    i of k. A zero-length key segment yields zeros."""
    cu_q = np.asarray(cu_seqlens_q).astype(np.int64)
    cu_k = np.asarray(cu_seqlens_k).astype(np.int64)
    qr = np.stack([cu_q[:-1], cu_q[1:]], axis=1)
    kr = np.stack([cu_k[:-1], cu_k[1:]], axis=1)
    out, _ = attn_ranges(q, k, v, qr, kr, softmax_scale)
    return out


def lse_merge(out_a, lse_a, out_b, lse_b) -> Tuple[np.ndarray, np.ndarray]:
    """Merge of two partial softmax-attention results over disjoint key sets (ring-KV step; new capability, the
    reference has no ring attention). Mathematical identity: softmax over the union."""
    mx = np.maximum(lse_a, lse_b)
    safe = np.where(np.isfinite(mx), mx, 0.0)
    ea, eb = np.exp(lse_a - safe), np.exp(lse_b - safe)
    den = ea + eb
    ok = den > 0
    wa = np.where(ok, ea / np.where(ok, den, 1), 0.0).astype(F32)
    wb = np.where(ok, eb / np.where(ok, den, 1), 0.0).astype(F32)
    out = out_a.astype(F32) * wa[..., None] + out_b.astype(F32) * wb[..., None]
    lse = np.where(ok, safe + np.log(np.where(ok, den, 1)), -np.inf).astype(F32)
    return bf16_round(out), lse


# ----------------------------------------------------------------------------------------------- ranges / schedule
def generate_sequences(chunk_num: int, window_size: int, chunk_offset: int):
    """generate_sequences, inference/pipeline/video_generate.py:164-180."""
    start_index = chunk_offset
    end_index = chunk_num + window_size - 1
    clip_start = [max(chunk_offset, i - window_size + 1) for i in range(start_index, end_index)]
    clip_end = [min(chunk_num, i + 1) for i in range(start_index, end_index)]
    t_start = [max(0, i - chunk_num + 1) for i in range(start_index, end_index)]
    t_end = [min(window_size, i - chunk_offset + 1) if i - chunk_offset < window_size else window_size
             for i in range(start_index, end_index)]
    return clip_start, clip_end, t_start, t_end


def generate_default_kvrange(slice_point: int, denoising_range_num: int, batch_size: int,
                             chunk_token_nums: int) -> np.ndarray:
    """generate_default_kvrange, video_generate.py:454-465."""
    range_num = slice_point + denoising_range_num
    rows = []
    for i in range(batch_size):
        for j in range(denoising_range_num):
            rows.append([i * range_num, slice_point + 1 + j + i * range_num])
    return (np.asarray(rows, dtype=np.int32) * chunk_token_nums).astype(np.int32)


def generate_noise2clean_kvrange(slice_point: int, denoising_range_num: int, noise2clean_kvrange: Sequence[int],
                                 clean_chunk_kvrange: int, denoise_step_of_each_chunk: Sequence[int], num_steps: int,
                                 batch_size: int, chunk_token_nums: int) -> np.ndarray:
    """generate_noise2clean_kvrange, video_generate.py:467-508."""
    assert len(denoise_step_of_each_chunk) == denoising_range_num and len(noise2clean_kvrange) > 0
    if clean_chunk_kvrange == -1:
        clean_chunk_kvrange = noise2clean_kvrange[-1]
    assert num_steps % len(noise2clean_kvrange) == 0
    per_stage = num_steps // len(noise2clean_kvrange)
    widths = [clean_chunk_kvrange if s == num_steps else noise2clean_kvrange[s // per_stage]
              for s in denoise_step_of_each_chunk]
    range_num = slice_point + denoising_range_num
    rows = []
    for i in range(batch_size):
        base = i * range_num
        for j in range(denoising_range_num):
            end = slice_point + j + 1
            start = max(0, end - widths[j])
            rows.append([(base + start) * chunk_token_nums, (base + end) * chunk_token_nums])
    return np.asarray(rows, dtype=np.int32)


def core_q_ranges(clip_token_nums: int, denoising_range_num: int, batch_size: int) -> np.ndarray:
    """q_range construction, dit_model.py:220-246: consecutive clips of clip_token_nums."""
    cu = np.arange(0, denoising_range_num * batch_size + 1, dtype=np.int64) * clip_token_nums
    return np.stack([cu[:-1], cu[1:]], axis=1).astype(np.int32)


def cross_attn_params(clip_token_nums: int, denoising_range_num: int, batch_size: int, y_index: Sequence[int]):
    """cu_seqlens / ranges of the cross-attention, dit_model.py:215-241. y_index = valid caption tokens per
    (batch x chunk). Returns dict(q_ranges, kv_ranges, cu_seqlens_q, cu_seqlens_kv)."""
    cu_q = np.concatenate([[0], np.full(denoising_range_num * batch_size, clip_token_nums)]).cumsum().astype(np.int32)
    cu_k = np.concatenate([[0], np.asarray(y_index)]).cumsum().astype(np.int32)
    return dict(q_ranges=np.stack([cu_q[:-1], cu_q[1:]], 1), kv_ranges=np.stack([cu_k[:-1], cu_k[1:]], 1),
                cu_seqlens_q=cu_q, cu_seqlens_kv=cu_k)


def max_seqlen_k(kv_range: np.ndarray) -> int:
    """dit_model.py:247-248: unique(kv_range)[-1] - unique(kv_range)[0]."""
    u = np.unique(kv_range)
    return int(u[-1] - u[0])


# ----------------------------------------------------------------------------------------------- context parallel
def cp_split_sizes(seq_len: int, cp_size: int) -> List[int]:
    """cp_ulysses_process Part1, context_parallel.py:240-242."""
    sizes = [seq_len // cp_size] * cp_size
    for i in range(seq_len % cp_size):
        sizes[i] += 1
    return sizes


def scatter_to_cp_region(x: np.ndarray, split_sizes: Sequence[int], rank: int) -> np.ndarray:
    """scatter_to_context_parallel_region (cp_shuffle_num == 1), context_parallel.py:50-52."""
    off = sum(split_sizes[:rank])
    return x[off:off + split_sizes[rank]].copy()


def cp_update_cross_attn_qkv_range(cu_seqlens_q, cu_seqlens_k, batch_size: int, split_sizes: Sequence[int],
                                   cp_rank: int):
    """cp_update_cross_attn_qkv_range (cp_shuffle_num == 1, cp_pad_size == 0), context_parallel.py:134-224.
    Returns dict(q_ranges, kv_ranges, cu_seqlens_q, cu_seqlens_kv, max_seqlen_q)."""
    sizes = np.asarray(split_sizes, dtype=np.int64)
    bounds = np.concatenate([[0], sizes.cumsum()])
    total = int(bounds[-1])
    cu_q = np.asarray(cu_seqlens_q, dtype=np.int64)
    cu_k = np.asarray(cu_seqlens_k, dtype=np.int64)
    qs, qe = cu_q[:-1], cu_q[1:]
    xq, xk = [], []
    for i in range(batch_size):
        b = bounds + i * total
        s, e = b[cp_rank], b[cp_rank + 1]
        st, en = np.maximum(s, qs), np.minimum(e, qe)
        m = st < en
        part_q = np.stack([st[m], en[m]], axis=1)
        part_q = part_q - part_q[:, 0].min()
        xq.append(part_q)
        xk.append(np.stack([cu_k[:-1][m], cu_k[1:][m]], axis=1))
    offs = np.concatenate([[0], np.cumsum([r[-1, -1] for r in xq])[:-1]])
    q_ranges = np.concatenate([r + int(o) for r, o in zip(xq, offs)], axis=0).astype(np.int32)
    k_ranges = np.concatenate(xk, axis=0).astype(np.int32)
    return dict(q_ranges=q_ranges, kv_ranges=k_ranges, cu_seqlens_q=np.unique(q_ranges).astype(np.int32),
                cu_seqlens_kv=np.unique(k_ranges).astype(np.int32), max_seqlen_q=int(split_sizes[cp_rank]))


def ulysses_input_split(per_rank: List[np.ndarray], split_sizes: Sequence[int]) -> List[np.ndarray]:
    """all_to_all_input_split as index math, context_parallel.py:381-401. per_rank[r]: (seq_r, H, hd) with
    seq_r == split_sizes[r]. Returns per-rank (sum(split_sizes), H/cp, hd): all tokens, local heads.
    KV heads are replicated when cp % H == 0 and cp != H (context_parallel.py:393-394)."""
    cp = len(per_rank)
    outs = []
    for dst in range(cp):
        blocks = []
        for src in range(cp):
            t = per_rank[src]
            hn = t.shape[1]
            if cp % hn == 0 and cp != hn:
                t = np.repeat(t, cp // hn, axis=1)
            hl = t.shape[1] // cp
            blocks.append(t[:, dst * hl:(dst + 1) * hl])
        outs.append(np.concatenate(blocks, axis=0))
    return outs


def ulysses_output_split(per_rank: List[np.ndarray], split_sizes: Sequence[int]) -> List[np.ndarray]:
    """all_to_all_output_split followed by "(cp sq) hn hd -> sq (cp hn) hd", context_parallel.py:404-423, 596
    (batch folded into sq). per_rank[r]: (sum(split_sizes), hn, hd). Returns per-rank (split_sizes[r], cp*hn, hd)."""
    cp = len(per_rank)
    bounds = np.concatenate([[0], np.cumsum(split_sizes)])
    outs = []
    for dst in range(cp):
        s, e = int(bounds[dst]), int(bounds[dst + 1])
        outs.append(np.concatenate([per_rank[src][s:e] for src in range(cp)], axis=1))
    return outs


# ----------------------------------------------------------------------------------------------- CSO (cp_shuffle_overlap)
def cso_pad_and_split(seq_len: int, cp_shuffle_num: int, cp_size: int) -> Tuple[int, List[int]]:
    """cp_shuffle_overlap_process Part1, context_parallel.py:271-275."""
    assert seq_len % cp_shuffle_num == 0
    chunk = seq_len // cp_shuffle_num
    pad = 0 if chunk % cp_size == 0 else (cp_size - chunk % cp_size) * cp_shuffle_num
    return pad, [(seq_len + pad) // cp_size] * cp_size


def cso_scatter(x: np.ndarray, split_sizes: Sequence[int], cp_shuffle_num: int, cp_pad_size: int, rank: int):
    """scatter_to_context_parallel_region with cp_shuffle_num > 1, context_parallel.py:40-48: every chunk is
    zero-padded and split on its own; a rank keeps the same slice of every chunk."""
    pad = cp_pad_size // cp_shuffle_num
    sizes = [s // cp_shuffle_num for s in split_sizes]
    off = sum(sizes[:rank])
    outs = []
    for c in np.split(x, cp_shuffle_num, axis=0):
        c = np.concatenate([c, np.zeros((pad,) + c.shape[1:], dtype=c.dtype)], axis=0)
        outs.append(c[off:off + sizes[rank]])
    return np.concatenate(outs, axis=0)


def cso_gather(per_rank: List[np.ndarray], cp_shuffle_num: int, cp_pad_size: int) -> np.ndarray:
    """gather_from_context_parallel_region with cp_shuffle_num > 1, context_parallel.py:74-84."""
    pad = cp_pad_size // cp_shuffle_num
    chunks = [np.split(o, cp_shuffle_num, axis=0) for o in per_rank]
    outs = []
    for i in range(cp_shuffle_num):
        c = np.concatenate([ch[i] for ch in chunks], axis=0)
        outs.append(c[:c.shape[0] - pad])
    return np.concatenate(outs, axis=0)


def cso_update_cross_attn_qkv_range(cu_seqlens_q, cu_seqlens_k, batch_size: int, split_sizes: Sequence[int],
                                    cp_rank: int, cp_shuffle_num: int, cp_pad_size: int):
    """cp_update_cross_attn_qkv_range with cp_shuffle_num > 1, context_parallel.py:134-224."""
    sizes = np.asarray([s // cp_shuffle_num for s in split_sizes], dtype=np.int64)
    bounds = np.concatenate([[0], sizes.cumsum()])
    total = int(bounds[-1])
    cu_q = np.asarray(cu_seqlens_q, dtype=np.int64)
    cu_q = cu_q + np.arange(cu_q.shape[0]) * (cp_pad_size // cp_shuffle_num)      # :164-167
    cu_k = np.asarray(cu_seqlens_k, dtype=np.int64)
    qs, qe = cu_q[:-1], cu_q[1:]
    xq, xk = [], []
    for i in range(batch_size):
        iq, ik = [], []
        for j in range(cp_shuffle_num):
            b = bounds + (i * cp_shuffle_num + j) * total
            st, en = np.maximum(b[cp_rank], qs), np.minimum(b[cp_rank + 1], qe)
            m = st < en
            part = np.stack([st[m], en[m]], axis=1)
            iq.append(part - part[:, 0].min())
            ik.append(np.stack([cu_k[:-1][m], cu_k[1:][m]], axis=1))
        ioffs = np.concatenate([[0], np.cumsum([r[-1, -1] for r in iq])[:-1]])
        xq.append(np.concatenate([r + int(o) for r, o in zip(iq, ioffs)], axis=0))
        xk.append(np.concatenate(ik, axis=0))
    offs = np.concatenate([[0], np.cumsum([r[-1, -1] for r in xq])[:-1]])
    q_ranges = np.concatenate([r + int(o) for r, o in zip(xq, offs)], axis=0).astype(np.int32)
    k_ranges = np.concatenate(xk, axis=0).astype(np.int32)
    return dict(q_ranges=q_ranges, kv_ranges=k_ranges, max_seqlen_q=int(split_sizes[cp_rank]))


def cso_core_attention(q_per_rank: List[np.ndarray], kv_per_rank: List[np.ndarray], cp_shuffle_num: int,
                       clip_token_nums: int, np_k_range: np.ndarray, cache_prefix: Optional[np.ndarray] = None):
    """The cp_shuffle_overlap branch of FullyParallelAttention.forward as index math (dit_module.py:1207-1242 with
    cso_communication / CSOHelper, context_parallel.py:603-664), all ranks simulated in one process, heads kept
    together (the head all-to-alls only move data).
    q_per_rank[r] (dn * spb, Hq, 128), kv_per_rank[r] (dn * spb, Hkv, 256): rank r's slice of every padded chunk.
    The key/value sequence is re-assembled chunk by chunk and un-padded (dit_module.py:1215-1224), optionally
    preceded by the KV-cache prefix; the padded query chunk i (all ranks' slices, rank-major) attends k_range[i]
    (full_attention, dit_module.py:1068-1096). Returns per-rank outputs (dn * spb, Hq, 128)."""
    cp, dn = len(q_per_rank), cp_shuffle_num
    spb = q_per_rank[0].shape[0] // dn
    kv = np.concatenate([np.concatenate([kv_per_rank[r][d * spb:(d + 1) * spb] for r in range(cp)], axis=0)
                         [:clip_token_nums] for d in range(dn)], axis=0)
    if cache_prefix is not None:
        kv = np.concatenate([cache_prefix, kv], axis=0)
    outs = [np.zeros_like(q) for q in q_per_rank]
    for d in range(dn):
        q = np.concatenate([q_per_rank[r][d * spb:(d + 1) * spb] for r in range(cp)], axis=0)
        o, _ = attn_ranges(q, kv[..., :128], kv[..., 128:], np.array([[0, q.shape[0]]]), np.asarray(np_k_range)[d:d + 1])
        for r in range(cp):
            outs[r][d * spb:(d + 1) * spb] = o[r * spb:(r + 1) * spb]
    return outs


# ----------------------------------------------------------------------------------------------- adaLN gating chain
def softcap(x: np.ndarray, cap: float = 1.0) -> np.ndarray:
    """dit_module.py:360-361 on a bf16 tensor: (cap * tanh(x.float() / cap)).to(bf16)."""
    return bf16_round(F32(cap) * np.tanh(x.astype(F32) / F32(cap)))


def bias_modulate_add(x: np.ndarray, residual: np.ndarray, condition_map: np.ndarray, gate: np.ndarray,
                      ln_w: np.ndarray, ln_b: np.ndarray, eps: float = 1e-6, zero_centered_gamma: bool = True):
    """bias_modulate_add, dit_module.py:292-307, with range_mod_triton (:244-289; the Triton program :201-241 is
    y[row] = x[row] * gatings[map[row]]) and the fp32 FusedLayerNorm post_norm (:355-357).
    x (s, b, h) holding fp32 (or bf16) values, residual (s, b, h) bf16 values, condition_map (s, b) indices into
    gate.reshape(-1, h), gate (b, dn, h). Returns the bf16-rounded result the layer keeps
    (`.to(params_dtype)`, dit_module.py:1359-1361)."""
    h = x.shape[-1]
    g = gate.reshape(-1, h).astype(F32)[np.asarray(condition_map).astype(np.int64)]     # (s, b, h)
    t = x.astype(F32) * g
    y = layer_norm(t, ln_w, ln_b, eps, zero_centered_gamma)
    return bf16_round(y + residual.astype(F32))


# ----------------------------------------------------------------------------------------------- whole block
def linear_bf16(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray] = None) -> np.ndarray:
    """nn.Linear on bf16 tensors: fp32 accumulate, one rounding of the result to bf16."""
    y = x.astype(F32) @ w.astype(F32).T
    if b is not None:
        y = y + b.astype(F32)
    return bf16_round(y)


# ----------------------------------------------------------------------------------------------- fp8 linears
E4M3_MAX = F32(448.0)   # torch.finfo(torch.float8_e4m3fn).max


def e4m3_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 values with |x| <= 448 to the nearest float8_e4m3fn value (ties to even), returned as fp32: 3
    mantissa bits, exponent bias 7, smallest normal 2^-6, subnormal quantum 2^-9 (what `.to(torch.float8_e4m3fn)`
    does for in-range values)."""
    x = np.asarray(x, dtype=np.float64)
    a = np.abs(x)
    e = np.floor(np.log2(np.where(a > 0, a, 1.0)))
    q = np.exp2(np.maximum(e, -6.0) - 3.0)           # spacing of representable values around x
    return (np.rint(x / q) * q).astype(F32)          # np.rint rounds half to even


def e4m3_encode(x: np.ndarray) -> np.ndarray:
    """The byte pattern (sign | 4-bit exponent | 3-bit mantissa) of values that are exactly representable."""
    x = np.asarray(x, dtype=np.float64)
    a = np.where(np.isfinite(x), np.abs(x), 0.0)      # NaN / infinity have no meaningful pattern here: encoded as 0
    e = np.maximum(np.floor(np.log2(np.where(a > 0, a, 1.0))), -6.0)
    normal = a >= 2.0 ** -6
    mant = np.where(normal, np.rint((a / np.exp2(e) - 1.0) * 8.0), np.rint(a * 512.0)).astype(np.int64)
    expo = np.where(normal, e.astype(np.int64) + 7, 0)
    sign = np.signbit(x).astype(np.int64) << 7
    return (sign | (expo << 3) | mant).astype(np.uint8)


def e4m3_decode(b: np.ndarray) -> np.ndarray:
    b = np.asarray(b, dtype=np.uint8).astype(np.int64)
    expo, mant = (b >> 3) & 15, b & 7
    v = np.where(expo > 0, (1.0 + mant / 8.0) * np.exp2(expo - 7.0), mant / 8.0 * 2.0 ** -6)
    v = np.where((b & 0x7F) == 0x7F, np.nan, v)
    return np.where(b & 0x80, -v, v).astype(F32)


def div_clamp_to(x: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """div_clamp_to, dit_module.py:364-384: clamp(x.float() / scale.float(), -448, 448).bfloat16().to(float8_e4m3fn);
    `scale` broadcasts against the last dimension. Returns the fp8 VALUES as fp32 (e4m3_encode gives the bytes). The
    256 MB partitioning of :370-380 only bounds the memory peak."""
    r = (np.asarray(x, dtype=F32) / np.asarray(scale, dtype=F32).reshape(-1)).astype(F32)
    r = np.clip(r, -E4M3_MAX, E4M3_MAX)
    return e4m3_round(bf16_round(r))


def linear_fp8(x: np.ndarray, weight_bytes: np.ndarray, quant_scale: np.ndarray, input_scale: np.ndarray,
               weight_scale: np.ndarray) -> np.ndarray:
    """PerTensorQuantizedFp8Linear.forward (dit_module.py:446-457; quant_scale = input_scale, a vector of in_features)
    and PerChannelQuantizedFp8Linear.forward (:478-488; quant_scale = smooth_scale (1, in_features), input_scale a
    scalar): div_clamp_to, then `flashinfer.bmm_fp8(xq, weight^T, input_scale, weight_scale, bf16)`. bmm_fp8 is a
    third-party CUDA kernel absent from /root/reference (flashinfer==0.2.0.post2, requirements.txt:21); its published
    definition is the cuBLASLt fp8 matmul with per-tensor scale pointers: out = bf16((A @ B) * a_scale * b_scale) with
    fp32 accumulation, cuBLASLt reading ONE float behind each pointer (the first element of a vector input_scale).
    weight_bytes: (1, out, in) or (out, in) uint8 e4m3 patterns."""
    xq = div_clamp_to(x, quant_scale)
    wq = e4m3_decode(np.asarray(weight_bytes).reshape(-1, x.shape[-1]))
    acc = xq.astype(F32) @ wq.T
    s = F32(np.asarray(input_scale, dtype=F32).reshape(-1)[0]) * F32(np.asarray(weight_scale, dtype=F32).reshape(-1)[0])
    return bf16_round(acc * s)


def linear_named(w, name: str, x: np.ndarray) -> np.ndarray:
    """The linear `name` of a state_dict, by the class its parameters reveal (dit_module.py:405-410, 520-540, 873-880):
    smooth_scale -> PerChannelQuantizedFp8Linear, input_scale -> PerTensorQuantizedFp8Linear, else nn.Linear."""
    if name + ".smooth_scale" in w:
        return linear_fp8(x, w[name + ".weight"], w[name + ".smooth_scale"], w[name + ".input_scale"],
                          w[name + ".weight_scale"])
    if name + ".input_scale" in w:
        return linear_fp8(x, w[name + ".weight"], w[name + ".input_scale"], w[name + ".input_scale"],
                          w[name + ".weight_scale"])
    return linear_bf16(x, w[name + ".weight"])


def gelu(x: np.ndarray) -> np.ndarray:
    """torch.nn.functional.gelu (erf form) on a bf16 tensor: fp32 math, bf16 result."""
    from scipy.special import erf
    x = x.astype(F32)
    return bf16_round(F32(0.5) * x * (F32(1.0) + erf(x / np.sqrt(F32(2.0))).astype(F32)))


def silu_and_mul(x: np.ndarray) -> np.ndarray:
    """flashinfer.activation.silu_and_mul (dit_module.py:546; third-party CUDA kernel absent from /root/reference,
    flashinfer pinned nowhere in requirements.txt - published definition): the last dimension holds [a | b];
    out = silu(a) * b in fp32, ONE rounding to the storage type."""
    x = x.astype(F32)
    f = x.shape[-1] // 2
    a, b = x[..., :f], x[..., f:]
    return bf16_round(a / (F32(1.0) + np.exp(-a)) * b)


def mlp_activation(x: np.ndarray, gated: bool) -> np.ndarray:
    """CustomMLP.forward between fc1 and fc2, dit_module.py:545-548."""
    return silu_and_mul(x) if gated else gelu(x)


def fully_parallel_attention(w, hs, kvs, rope, cache, has_inf, slice_point, dn, clip, np_q_range, np_k_range, y_index,
                             update_kv_cache, fwd_extra_1st_chunk=False, distill_nearly_clean_chunk=False, pre=""):
    """FullyParallelAttention.forward with cp_strategy "none", dit_module.py:1138-1174, composed from the pieces above.
    w: name -> array of the module's state_dict (prefix `pre`); hs (sq, b, hidden), kvs (ytot, hidden) bf16 values.
    Returns (core_attn_out, xattn_out), each (sq, b, Hq * 128)."""
    g = lambda n: w[pre + n]
    sq, b, _ = hs.shape
    Hq = g("linear_qkv.q.weight").shape[-2] // 128      # (out, in), or (1, out, in) for the fp8 linears
    Hkv = g("linear_qkv.k.weight").shape[-2] // 128
    mixed = bf16_round(layer_norm(hs, g("linear_qkv.layer_norm.weight"), g("linear_qkv.layer_norm.bias"), 1e-6,
                                  zero_centered_gamma=False))
    lin = {n: linear_named(w, pre + f"linear_qkv.{n}", mixed) for n in ("q", "qx", "k", "v")}
    sin, cos = rope[:, :48], rope[:, 48:]
    q = get_q_or_k(lin["q"], Hq, g("q_layernorm.weight"), g("q_layernorm.bias"), sin, cos)
    kv = get_kv(lin["k"], lin["v"], Hkv, g("k_layernorm.weight"), g("k_layernorm.bias"), sin, cos)
    key, value = adjust_key_and_value_for_inference(kv, cache, has_inf, slice_point, clip, 1, update_kv_cache, False,
                                                    fwd_extra_1st_chunk, distill_nearly_clean_chunk)
    core = core_attention(q, key, value, b, np_q_range, np_k_range, dn).reshape(sq, b, Hq * 128)
    # get_xqkv + cross_attention, dit_module.py:1005-1021, 1098-1136
    xq = np.transpose(lin["qx"].reshape(sq, b, Hq, 128), (1, 0, 2, 3)).reshape(b * sq, Hq, 128)
    xq = xattn_layer_norm(xq, g("q_layernorm_xattn.weight"), g("q_layernorm_xattn.bias"))
    kvx = linear_bf16(kvs, g("linear_kv_xattn.weight")).reshape(kvs.shape[0], Hkv, 256)
    xk = xattn_layer_norm(kvx[..., :128], g("k_layernorm_xattn.weight"), g("k_layernorm_xattn.bias"))
    p = cross_attn_params(clip, dn, b, y_index)
    xo = attn_varlen(xq, xk, kvx[..., 128:], p["cu_seqlens_q"], p["cu_seqlens_kv"])
    xo = np.transpose(xo.reshape(b, sq, Hq * 128), (1, 0, 2))
    return core, xo


def transformer_layer(w, hs, condition, condition_map, kvs, rope, cache, has_inf, slice_point, dn, clip, np_q_range,
                      np_k_range, y_index, update_kv_cache, gated_linear_unit=False, fwd_extra_1st_chunk=False):
    """TransformerLayer.forward, dit_module.py:1297-1367: attention -> "sq b (n hn hd) -> sq b (hn n hd)" (n = 2,
    hn = 8) -> linear_proj (the float32 autocast of :1343 is not supported by CUDA autocast and disables itself, so
    the projection runs in bf16) -> softcap(adaLN gates) -> bias_modulate_add -> MLP -> bias_modulate_add."""
    core, xo = fully_parallel_attention(w, hs, kvs, rope, cache, has_inf, slice_point, dn, clip, np_q_range, np_k_range,
                                        y_index, update_kv_cache, fwd_extra_1st_chunk=fwd_extra_1st_chunk,
                                        pre="self_attention.")
    sq, b, width = core.shape
    x = np.stack([core, xo], axis=2).reshape(sq, b, 2, 8, width // 8)
    x = np.transpose(x, (0, 1, 3, 2, 4)).reshape(sq, b, 2 * width)
    attn_out = linear_named(w, "self_attention.linear_proj", x)
    c = condition.astype(F32)
    silu = bf16_round(c / (F32(1.0) + np.exp(-c)))
    gate = softcap(linear_bf16(silu, w["ada_modulate_layer.proj.0.weight"], w["ada_modulate_layer.proj.0.bias"]), 1.0)
    h = hs.shape[-1]
    gate_msa, gate_mlp = gate[..., :h], gate[..., h:]
    x1 = bias_modulate_add(attn_out, hs, condition_map, gate_msa, w["self_attn_post_norm.weight"],
                           w["self_attn_post_norm.bias"])
    m = bf16_round(layer_norm(x1, w["mlp.layer_norm.weight"], w["mlp.layer_norm.bias"], 1e-6, zero_centered_gamma=False))
    m = linear_named(w, "mlp.linear_fc1", m)
    m = mlp_activation(m, gated_linear_unit)
    m = linear_named(w, "mlp.linear_fc2", m)
    return bias_modulate_add(m, x1, condition_map, gate_mlp, w["mlp_post_norm.weight"], w["mlp_post_norm.bias"])


def transformer_block(layer_weights, final_ln_w, final_ln_b, hs, condition, condition_map, kvs, rope, caches, has_inf,
                      slice_point, dn, clip, np_q_range, np_k_range, y_index, update_kv_cache,
                      gated_linear_unit=False, fwd_extra_1st_chunk=False, collect=None):
    """TransformerBlock.forward, dit_module.py:1413-1441 (pre_process and post_process both True, pp_size 1): the
    layers in order, each with its own KV cache, then the final LayerNorm (+1p) on the fp32 hidden state. `collect`, if
    a list, receives the hidden state after every layer."""
    x = hs
    for w, cache in zip(layer_weights, caches):
        x = transformer_layer(w, x, condition, condition_map, kvs, rope, cache, has_inf, slice_point, dn, clip,
                              np_q_range, np_k_range, y_index, update_kv_cache, gated_linear_unit=gated_linear_unit,
                              fwd_extra_1st_chunk=fwd_extra_1st_chunk)
        if collect is not None:
            collect.append(x)
    return layer_norm(x, final_ln_w, final_ln_b, 1e-6, zero_centered_gamma=True)


# ----------------------------------------------------------------------------------------------- RoPE table
def rope_table(T_total: int, H: int, W: int, bands: Optional[np.ndarray] = None, keep_last_T: Optional[int] = None):
    """LearnableRotaryEmbeddingCat.get_embed as used at dit_model.py:163-166 (in_pixels=False, dim=128):
    bands = 1/10000^(arange(16)/16) (dit_module.py:591-596); t = arange, h/w centred (dit_module.py:644-647) and
    rescaled x/(f-1)*(r-1) with r = H/sqrt(HW/256) (dit_module.py:648-659, dit_model.py:163-164).
    Returns (sin, cos), each (rows, 48); rows = last keep_last_T*H*W rows (dit_model.py:166)."""
    if bands is None:
        bands = (1.0 / (10000.0 ** (np.arange(16, dtype=F32) / F32(16)))).astype(F32)
    rescale = math.sqrt((H * W) / (16 * 16))
    feat = [T_total, H, W]
    ref = [T_total, H / rescale, W / rescale]
    t = [np.arange(s, dtype=F32) for s in feat]
    t[1] = t[1] - F32((feat[1] - 1) / 2)
    t[2] = t[2] - F32((feat[2] - 1) / 2)
    tr = []
    for x, f, r in zip(t, feat, ref):
        tr.append(x if f == 1 else (x / F32(f - 1) * F32(r - 1)).astype(F32))
    grid = np.stack(np.meshgrid(*tr, indexing="ij"), axis=-1)[..., None].astype(F32)  # T H W 3 1
    pos = grid * bands.astype(F32)                                                    # T H W 3 16
    sin = np.sin(pos).reshape(T_total * H * W, -1).astype(F32)
    cos = np.cos(pos).reshape(T_total * H * W, -1).astype(F32)
    if keep_last_T is not None:
        n = keep_last_T * H * W
        sin, cos = sin[-n:], cos[-n:]
    return sin, cos
