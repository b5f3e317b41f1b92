"""ORACLE - test infrastructure only. The reference's attention calls restated with torch fp32 CPU ops: the "reference's
CPU torch path" that BASELINE.json's north_star asks to be timed on the box's host cores next to the GPU numbers.

Nothing in the product (`magi-1_b200/`) imports this module; it is used by `bench.py`'s `cpu_baseline` /
`--impl reference` legs (as the reported baseline) and by `tests/test_oracle_golden.py`, which holds it to the numpy
oracle (`magi_oracle.py`) and through it to the fixtures produced by the reference's own Python code.

Why a second restatement: `magi_oracle.py` is written for checking (explicit fp32 softmax, log-sum-exp returned, one
head group at a time) and is several times slower than what the host can do. This file is the same arithmetic through
`torch.nn.functional.scaled_dot_product_attention` (fused CPU kernel, all host threads), i.e. the strongest CPU form
of the reference's own fallback branch - which is a loop of `flash_attn_func` calls, each = softmax(Q K^T / sqrt(d)) V
with the GQA head mapping h -> h // (Hq / Hkv) (flash-attn==2.4.2, requirements.txt:20; CUDA-only, so it cannot itself
run on a CPU).

Reference: samiazirar/MAGI-1 (file:line relative to /root/reference).
"""
import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def flash_attn_func(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, softmax_scale: Optional[float] = None):
    """flash_attn_func as the reference calls it (dit_module.py:1062, 1090): q (b, sq, Hq, d), k/v (b, sk, Hkv, d),
    no mask, no dropout -> (b, sq, Hq, d). fp32 on the CPU."""
    hq, hkv = q.shape[2], k.shape[2]
    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    if hq != hkv:
        kt = kt.repeat_interleave(hq // hkv, dim=1)      # flash-attn's GQA convention
        vt = vt.repeat_interleave(hq // hkv, dim=1)
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(q.shape[-1])
    return F.scaled_dot_product_attention(qt, kt, vt, scale=scale).transpose(1, 2)


def core_attention(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, q_range, k_range,
                   denoising_range_num: Optional[int] = None, first: int = 0) -> torch.Tensor:
    """core_attention, fallback branch, dit_module.py:1046-1065 with bs == 1: for every denoising range i the query
    rows q_range[i] attend densely to the key rows k_range[i]; outputs concatenated along the sequence.
    query (sq, Hq, 128), key/value (sk, Hkv, 128) fp32 -> (sum of range lengths, Hq, 128).
    `first`/`denoising_range_num` select a sub-list of the ranges (bench.py times a bounded sample)."""
    n = len(q_range) if denoising_range_num is None else denoising_range_num
    outs = []
    for i in range(first, first + n):
        q0, q1 = int(q_range[i][0]), int(q_range[i][1])
        k0, k1 = int(k_range[i][0]), int(k_range[i][1])
        if k1 <= k0:                                       # zero-length key segment -> zeros
            outs.append(torch.zeros_like(query[q0:q1]))
            continue
        outs.append(flash_attn_func(query[None, q0:q1], key[None, k0:k1], value[None, k0:k1])[0])
    return torch.cat(outs, dim=0) if outs else query[:0]


def cross_attention(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, cu_seqlens_q: Sequence[int],
                    cu_seqlens_k: Sequence[int]) -> torch.Tensor:
    """cross_attention, fallback branch, dit_module.py:1123-1133 (flash_attn_varlen_func): query segment i attends to
    caption segment i."""
    qr = [(int(cu_seqlens_q[i]), int(cu_seqlens_q[i + 1])) for i in range(len(cu_seqlens_q) - 1)]
    kr = [(int(cu_seqlens_k[i]), int(cu_seqlens_k[i + 1])) for i in range(len(cu_seqlens_k) - 1)]
    return core_attention(query, key, value, qr, kr)


def timed_self_attention_sample(clip: int, Hq: int, Hkv: int, k_chunks: Sequence[Tuple[int, int]], budget_s: float,
                                seed: int = 1234):
    """Times the self-attention of ONE layer of a workload state (query range i = chunk i of the in-flight chunks,
    key range = k_chunks[i] in chunk units) on the host: ranges are taken in order until the time budget is spent
    (always at least one). Returns (flops done, seconds, ranges done, threads)."""
    import time
    g = torch.Generator().manual_seed(seed)
    n = len(k_chunks)
    k_hi = max(e for _, e in k_chunks)
    q = torch.randn(n * clip, Hq, 128, generator=g)
    k = torch.randn(k_hi * clip, Hkv, 128, generator=g)
    v = torch.randn(k_hi * clip, Hkv, 128, generator=g)
    q_range = [(i * clip, (i + 1) * clip) for i in range(n)]
    k_range = [(s * clip, e * clip) for s, e in k_chunks]
    flops, secs, done = 0, 0.0, 0
    for i in range(n):
        t0 = time.perf_counter()
        core_attention(q, k, v, q_range, k_range, denoising_range_num=1, first=i)
        secs += time.perf_counter() - t0
        flops += 4 * 128 * Hq * clip * (k_range[i][1] - k_range[i][0])
        done += 1
        if secs >= budget_s:
            break
    return flops, secs, done, torch.get_num_threads()
