"""Host-side construction of the attention metadata the hot path consumes, without device->host synchronisation.

Reference (samiazirar/MAGI-1):
    k ranges of a forward           inference/pipeline/video_generate.py:454-527 (generate_default_kvrange,
                                    generate_noise2clean_kvrange, generate_kvrange_for_denoising_video)
    uncond / nearly-clean ranges    inference/model/dit/dit_model.py:92-101, 544-552
    q ranges, cross-attention       dit_model.py:198-241 (get_embedding_and_meta, parts 4-5)
    max_seqlen_k, PackedCoreAttn    dit_model.py:246-259, 301-337 (forward_pre_process)

What differs from the reference: the reference builds the k ranges as a device tensor, one tiny tensor per range, and
then reads them back (`.cpu().item()` for max_seqlen_k, two `.cpu().numpy()` for the numpy mirrors, dit_model.py:248,
304-305) - three stream synchronisations per forward. Here the ranges are plain integer
arithmetic on the host (they only depend on the schedule), the numpy mirrors ARE the source, and the device copies
are one pinned, non-blocking transfer per table. The caption lengths stay on the device (sum of the mask -> cumsum);
nothing on this path waits for the GPU, so the step can be captured in a CUDA graph or run ahead of the device.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import context_parallel as cpm
from .common import ModelMetaArgs, PackedCoreAttnParams, PackedCrossAttnParams


# ----------------------------------------------------------------------------------------------- k ranges (host)
def _batched(chunk_ranges: np.ndarray, range_num: int, batch_size: int, chunk_token_nums: int) -> np.ndarray:
    """(dn, 2) ranges in chunk units of ONE batch row -> (batch * dn, 2) token ranges; batch row i sees its own copy of
    the range_num chunks at chunk offset i * range_num (video_generate.py:461-465, 493-497)."""
    shift = (np.arange(batch_size, dtype=np.int64) * range_num)[:, None, None]
    return ((chunk_ranges[None].astype(np.int64) + shift) * chunk_token_nums).reshape(-1, 2).astype(np.int32)


def default_kv_range(slice_point: int, denoising_range_num: int, batch_size: int, chunk_token_nums: int) -> np.ndarray:
    """generate_default_kvrange (video_generate.py:454-465): denoising chunk j attends every chunk up to itself,
    [0, slice_point + j + 1)."""
    ends = slice_point + 1 + np.arange(denoising_range_num, dtype=np.int64)
    return _batched(np.stack([np.zeros_like(ends), ends], 1), slice_point + denoising_range_num, batch_size,
                    chunk_token_nums)


def noise2clean_kv_range(slice_point: int, denoising_range_num: int, noise2clean_kvrange: Sequence[int],
                         clean_chunk_kvrange: int, denoise_step_of_each_chunk: Sequence[int], num_steps: int,
                         batch_size: int, chunk_token_nums: int) -> np.ndarray:
    """generate_noise2clean_kvrange (video_generate.py:467-508): the look-back window of a chunk shrinks stage by stage
    as it gets cleaner; a finished chunk (denoise step == num_steps) uses clean_chunk_kvrange (-1: the last stage's)."""
    if len(denoise_step_of_each_chunk) != denoising_range_num or len(noise2clean_kvrange) == 0:
        raise ValueError("one denoise step per denoising chunk and a non-empty noise2clean_kvrange are required")
    if num_steps % len(noise2clean_kvrange) != 0:
        raise ValueError("num_steps must be a multiple of the number of noise2clean stages")
    clean = noise2clean_kvrange[-1] if clean_chunk_kvrange == -1 else clean_chunk_kvrange
    per_stage = num_steps // len(noise2clean_kvrange)
    steps = np.asarray(denoise_step_of_each_chunk, dtype=np.int64)
    stage = np.minimum(steps // per_stage, len(noise2clean_kvrange) - 1)
    width = np.where(steps == num_steps, clean, np.asarray(noise2clean_kvrange, dtype=np.int64)[stage])
    ends = slice_point + 1 + np.arange(denoising_range_num, dtype=np.int64)
    starts = np.maximum(0, ends - width)
    return _batched(np.stack([starts, ends], 1), slice_point + denoising_range_num, batch_size, chunk_token_nums)


def kv_range_for_denoising_video(slice_point: int, denoising_range_num: int, denoise_step_of_each_chunk: Sequence[int],
                                 noise2clean_kvrange: Sequence[int], clean_chunk_kvrange: int, num_steps: int,
                                 batch_size: int, chunk_token_nums: int) -> np.ndarray:
    """generate_kvrange_for_denoising_video (video_generate.py:510-527): runtime_config.noise2clean_kvrange == []
    selects the default ranges."""
    if len(noise2clean_kvrange) == 0:
        return default_kv_range(slice_point, denoising_range_num, batch_size, chunk_token_nums)
    return noise2clean_kv_range(slice_point, denoising_range_num, noise2clean_kvrange, clean_chunk_kvrange,
                                denoise_step_of_each_chunk, num_steps, batch_size, chunk_token_nums)


def kv_range_for_uncondition(batch_size: int, chunk_token_nums: int) -> np.ndarray:
    """generate_kv_range_for_uncondition (dit_model.py:92-101): every batch row (= chunk of the uncond pass) attends
    only itself."""
    edges = np.arange(batch_size + 1, dtype=np.int64) * chunk_token_nums
    return np.stack([edges[:-1], edges[1:]], 1).astype(np.int32)


def append_nearly_clean_range(kv_range: np.ndarray, chunk_token_nums: int) -> np.ndarray:
    """distill_nearly_clean_chunk (dit_model.py:544-552): the duplicate of the first denoising chunk is appended as a
    self-only range right after the last key of the forward, [kv_max, kv_max + chunk)."""
    kv_range = np.asarray(kv_range, dtype=np.int32).reshape(-1, 2)
    top = int(kv_range.max())
    return np.concatenate([kv_range, np.asarray([[top, top + chunk_token_nums]], dtype=np.int32)], 0)


# ----------------------------------------------------------------------------------------------- tables -> device
def _to_device(table: np.ndarray, device) -> torch.Tensor:
    """One non-blocking transfer from pinned memory (plain copy on CPU devices)."""
    t = torch.from_numpy(np.ascontiguousarray(table, dtype=np.int32))
    if torch.device(device).type == "cuda":
        return t.pin_memory().to(device, non_blocking=True)
    return t.to(device)


def condition_map(batch_size: int, denoising_range_num: int, seqlen_per_chunk: int, device) -> torch.Tensor:
    """dit_model.py:203-205: row (token s of the "(sq b)" order, batch b) -> index of its (batch, chunk) condition."""
    chunk = torch.arange(denoising_range_num, device=device).repeat_interleave(seqlen_per_chunk)     # (sq,)
    return (chunk[:, None] + torch.arange(batch_size, device=device)[None, :] * denoising_range_num).contiguous()


def core_attn_params(kv_range: np.ndarray, clip_token_nums: int, denoising_range_num: int, batch_size: int,
                     device) -> PackedCoreAttnParams:
    """dit_model.py:243-259, 301-308 without the three host reads. kv_range: host (batch * dn, 2) token ranges."""
    kv_range = np.ascontiguousarray(np.asarray(kv_range, dtype=np.int32).reshape(-1, 2))
    n = denoising_range_num * batch_size
    if kv_range.shape[0] != n:
        raise ValueError(f"kv_range must have batch_size * denoising_range_num = {n} rows, got {kv_range.shape[0]}")
    edges = np.arange(n + 1, dtype=np.int64) * clip_token_nums
    q_range = np.stack([edges[:-1], edges[1:]], 1).astype(np.int32)
    max_seqlen_k = int(kv_range.max() - kv_range.min())        # unique(kv_range)[-1] - unique(kv_range)[0]
    return PackedCoreAttnParams(q_range=_to_device(q_range, device), k_range=_to_device(kv_range, device),
                                np_q_range=q_range, np_k_range=kv_range, max_seqlen_q=clip_token_nums,
                                max_seqlen_k=max_seqlen_k)


def cross_attn_params(xattn_mask: torch.Tensor, clip_token_nums: int, denoising_range_num: int, batch_size: int,
                      caption_max_length: int) -> PackedCrossAttnParams:
    """dit_model.py:213-241: query segment i = chunk i of the "(b sq)" order, key segment i = the valid caption tokens
    of (batch, chunk) i. The caption lengths are reduced on the device they live on; no host read."""
    device = xattn_mask.device
    n = denoising_range_num * batch_size
    y_index = xattn_mask.reshape(n, -1).sum(dim=-1)
    cu_k = torch.zeros(n + 1, dtype=torch.int32, device=device)
    cu_k[1:] = y_index.cumsum(0).to(torch.int32)
    cu_q = torch.arange(n + 1, dtype=torch.int32, device=device) * clip_token_nums
    return PackedCrossAttnParams(q_ranges=torch.stack([cu_q[:-1], cu_q[1:]], 1),
                                 kv_ranges=torch.stack([cu_k[:-1], cu_k[1:]], 1), cu_seqlens_q=cu_q, cu_seqlens_kv=cu_k,
                                 max_seqlen_q=clip_token_nums, max_seqlen_kv=caption_max_length)


def pack_caption_tokens(y_xattn: torch.Tensor, xattn_mask: torch.Tensor) -> torch.Tensor:
    """y_xattn_flat (dit_model.py:208): the valid caption tokens of all (batch, chunk) slots, concatenated.
    The row count depends on the mask, so this is the one data-dependent shape of the forward (torch computes it)."""
    y = y_xattn.squeeze(1) if y_xattn.dim() == 4 else y_xattn
    return y[xattn_mask.reshape(y.shape[:-1]).bool()]


def build_meta_args(kv_range: np.ndarray, xattn_mask: torch.Tensor, x: torch.Tensor, rope: torch.Tensor, *,
                    batch_size: int, denoising_range_num: int, slice_point: int, range_num: int, H: int, W: int,
                    caption_max_length: int, cp_size: int = 1, cp_strategy: str = "none",
                    extract_prefix_video_feature: bool = False, fwd_extra_1st_chunk: bool = False,
                    distill_nearly_clean_chunk: bool = False
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, ModelMetaArgs]:
    """forward_pre_process from the embedded tokens on (dit_model.py:198-259, 301-337).

    x: embedded tokens "(T H W) N C"; rope: (T H W, 96) fp32 table; xattn_mask: (N * dn, L) or (N, dn, 1, L) validity
    of the caption tokens; kv_range: HOST token ranges of this forward. Returns (x, condition_map, rope, meta_args)
    with x / condition_map / rope already scattered to this rank's shard by cp_pre_process."""
    seq, n_batch = x.shape[0], x.shape[1]
    if n_batch != batch_size or seq % denoising_range_num != 0:
        raise ValueError("x must be (T H W, batch_size, C) with the token count a multiple of denoising_range_num")
    clip_token_nums = seq // denoising_range_num
    core = core_attn_params(kv_range, clip_token_nums, denoising_range_num, batch_size, x.device)
    cross = cross_attn_params(xattn_mask, clip_token_nums, denoising_range_num, batch_size, caption_max_length)
    cmap = condition_map(batch_size, denoising_range_num, clip_token_nums, x.device)
    ardf_meta = dict(clip_token_nums=clip_token_nums, slice_point=slice_point, range_num=range_num,
                     denoising_range_num=denoising_range_num, q_range=core.q_range, k_range=core.k_range,
                     max_seqlen_q=clip_token_nums, max_seqlen_k=core.max_seqlen_k)
    x, cmap, rope, cp_pad_size, cp_split_sizes, core, cross = cpm.cp_pre_process(
        cp_size, cp_strategy, x, cmap, rope, None, ardf_meta, core, cross)
    meta = ModelMetaArgs(H=H, W=W, cp_pad_size=cp_pad_size, cp_split_sizes=cp_split_sizes, slice_point=slice_point,
                         denoising_range_num=denoising_range_num, range_num=range_num,
                         extract_prefix_video_feature=extract_prefix_video_feature,
                         fwd_extra_1st_chunk=fwd_extra_1st_chunk,
                         distill_nearly_clean_chunk=distill_nearly_clean_chunk, clip_token_nums=clip_token_nums,
                         enable_cuda_graph=False, core_attn_params=core, cross_attn_params=cross)
    return x, cmap, rope, meta
