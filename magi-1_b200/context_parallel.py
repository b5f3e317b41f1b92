"""Context-parallel hooks of the hot path, same names / argument meaning as the reference's
inference/infra/parallelism/context_parallel.py, re-designed for one NVSwitch domain:

  * the all-to-all SEND buffers are produced directly by the fused prologue kernel (`head_perm_cp`), so the
    `repeat_interleave` + `rearrange().contiguous()` staging of context_parallel.py:393-396 disappears;
  * the KV all-to-all can land directly in the KV-cache slot (the cache stores post-all-to-all heads,
    dit_module.py:794-819), so the append copy of dit_module.py:929 disappears;
  * the receive-side "(cp sq) hn hd -> sq (cp hn hd)" of context_parallel.py:596 is one re-layout kernel writing each
    head group in place (no torch.cat of groups);
  * a ring-KV exchange (`RingKVScheduler`) is added: K/V blocks rotate over NVLink with point-to-point send/recv,
    double-buffered against the per-step attention, partial results merged by log-sum-exp.

Collectives go through torch.distributed (NCCL on the GPU box). Everything that computes is injected as a callable
(attention, re-layout, merge), exactly as the reference's UlyssesScheduler takes callables
(context_parallel.py:452-597): the product wires in the CUDA kernels, the CPU (gloo) tests wire in test doubles.
The cp_shuffle_overlap (CSO) strategy (context_parallel.py:257-305, 603-664) is built with the same entry points
(`cp_shuffle_overlap_process`, `cso_communication`, `CSOHelper`); on NVSwitch the fused "[q(i+1) | o(i-1)]" message of
the reference (one torch.concat + one chunk copy per step) is sent as two back-to-back all-to-alls instead, so no
staging copy is made.
"""
import math
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist

from . import parallel_state as mpu
from .common import ModelMetaArgs, PackedCoreAttnParams, PackedCrossAttnParams, divide


class FakeHandle:
    """context_parallel.py:89-94."""

    def wait(self):
        pass


#####################################################
# Common primitives
#####################################################
def scatter_to_context_parallel_region(input_, cp_split_sizes, cp_shuffle_num=1, cp_pad_size=0):
    """context_parallel.py:29-53 (contiguous split; no communication, the input is replicated)."""
    world_size = mpu.get_cp_world_size()
    if world_size == 1:
        return input_
    rank = mpu.get_cp_rank()
    if cp_shuffle_num > 1:
        # CSO: every one of the cp_shuffle_num chunks is zero-padded to a multiple of cp and split on its own, so a
        # rank holds the same slice of every chunk (context_parallel.py:40-48)
        pad = divide(cp_pad_size, cp_shuffle_num)
        sizes = [divide(s, cp_shuffle_num) for s in cp_split_sizes]
        lo = sum(sizes[:rank])
        n = sizes[rank]
        chunk = divide(input_.shape[0], cp_shuffle_num)
        out = input_.new_zeros((cp_shuffle_num, n) + tuple(input_.shape[1:]))
        valid = max(0, min(lo + n, chunk) - lo)      # rows of the slice that are not padding
        if valid > 0:
            out[:, :valid] = input_.unflatten(0, (cp_shuffle_num, chunk))[:, lo:lo + valid]
        return out.flatten(0, 1)
    dim_offset = sum(cp_split_sizes[:rank])
    if cp_pad_size:
        input_ = torch.nn.functional.pad(input_, [0, 0] * (input_.dim() - 1) + [0, cp_pad_size])
    return input_[dim_offset:dim_offset + cp_split_sizes[rank]].contiguous()


def gather_from_context_parallel_region(input_, cp_split_sizes, cp_shuffle_num=1, cp_pad_size=0):
    """context_parallel.py:56-86: all-gather of uneven sequence shards along dim 0."""
    world_size = mpu.get_cp_world_size()
    if world_size == 1:
        return input_
    input_ = input_.contiguous()
    dim_size = list(input_.size())
    dim_size[0] = sum(cp_split_sizes)
    output = torch.empty(dim_size, dtype=input_.dtype, device=input_.device)
    outputs = list(torch.split(output, cp_split_sizes, dim=0))
    dist.all_gather(outputs, input_, group=mpu.get_cp_group())
    if cp_shuffle_num > 1:
        # CSO (context_parallel.py:74-84): "(cp dn spb) -> (dn cp spb)", then drop every chunk's zero padding.
        # Equal shards (cp_shuffle_overlap_process pads to make them so): one permuted view, one copy.
        spb = divide(cp_split_sizes[0], cp_shuffle_num)
        assert all(s == cp_split_sizes[0] for s in cp_split_sizes)
        keep = spb * world_size - divide(cp_pad_size, cp_shuffle_num)
        v = output.unflatten(0, (world_size, cp_shuffle_num, spb)).transpose(0, 1).flatten(1, 2)
        return v[:, :keep].flatten(0, 1)
    if cp_pad_size:
        return output[:dim_size[0] - cp_pad_size]
    return output


#####################################################
# Context-parallel pre/post processing
#####################################################
def ulysses_split_sizes(seq_len: int, cp_size: int) -> List[int]:
    """context_parallel.py:240-242: S // cp (+1 on the first S % cp ranks)."""
    sizes = [seq_len // cp_size] * cp_size
    for i in range(seq_len % cp_size):
        sizes[i] += 1
    return sizes


def cp_update_cross_attn_qkv_range(cross_attn_params: PackedCrossAttnParams, batch_size: int,
                                   cp_split_sizes: List[int], device, cp_shuffle_num: int = 1, cp_pad_size: int = 0):
    """context_parallel.py:134-224: intersect every cross-attention q segment with this rank's sequence shard and
    rebase to shard-local rows; text K/V are replicated so kv ranges are only filtered."""
    cp_rank = mpu.get_cp_rank()
    seq_len_cur_rank = cp_split_sizes[cp_rank]
    # host-side interval arithmetic (a handful of segments); one D2H read per forward, like the reference's
    # boolean-mask indexing at context_parallel.py:182-190
    cu_q = cross_attn_params.cu_seqlens_q.detach().cpu().numpy().astype(np.int64)
    cu_k = cross_attn_params.cu_seqlens_kv.detach().cpu().numpy().astype(np.int64)
    # CSO: every q segment (one chunk) grows by the chunk's zero padding (context_parallel.py:164-167)
    cu_q = cu_q + np.arange(cu_q.shape[0], dtype=np.int64) * divide(cp_pad_size, cp_shuffle_num)
    sizes = [divide(x, cp_shuffle_num) for x in cp_split_sizes]
    bounds = np.concatenate([[0], np.cumsum(np.asarray(sizes, dtype=np.int64))])
    total = int(bounds[-1])
    qs, qe = cu_q[:-1], cu_q[1:]
    xq, xk = [], []
    for i in range(batch_size):
        inner_q, inner_k = [], []
        for j in range(cp_shuffle_num):
            off = (i * cp_shuffle_num + j) * total
            s, e = bounds[cp_rank] + off, bounds[cp_rank + 1] + off
            st, en = np.maximum(s, qs), np.minimum(e, qe)
            m = st < en
            part = np.stack([st[m], en[m]], axis=1)
            part = part - part[:, 0].min()
            inner_q.append(part)
            inner_k.append(np.stack([cu_k[:-1][m], cu_k[1:][m]], axis=1))
        ioffs = np.concatenate([[0], np.cumsum([r[-1, -1] for r in inner_q])[:-1]])
        xq.append(np.concatenate([r + int(o) for r, o in zip(inner_q, ioffs)], axis=0))
        xk.append(np.concatenate(inner_k, axis=0))
    offs = np.concatenate([[0], np.cumsum([r[-1, -1] for r in xq])[:-1]])
    q_ranges = np.concatenate([r + int(o) for r, o in zip(xq, offs)], axis=0)
    k_ranges = np.concatenate(xk, axis=0)
    # the reference derives cu_seqlens with torch.unique (context_parallel.py:211-215) and asserts equal lengths, which
    # trips on an empty caption segment; the kernels here consume q_ranges / kv_ranges, so cu_seqlens are rebuilt
    # segment by segment instead and stay well-formed in that case
    cu_q_new = np.concatenate([q_ranges[:1, 0], q_ranges[:, 1]])
    cu_k_new = np.concatenate([k_ranges[:1, 0], k_ranges[:, 1]])

    def t(a):
        return torch.as_tensor(a.astype(np.int32), device=device)

    return PackedCrossAttnParams(q_ranges=t(q_ranges), kv_ranges=t(k_ranges), cu_seqlens_q=t(cu_q_new),
                                 cu_seqlens_kv=t(cu_k_new), max_seqlen_q=seq_len_cur_rank,
                                 max_seqlen_kv=cross_attn_params.max_seqlen_kv)


def cp_ulysses_process(cp_size: int, x: torch.Tensor, condition_map: torch.Tensor, rope: torch.Tensor,
                       xattn_mask_for_cuda_graph: Union[torch.Tensor, None], cross_attn_params: PackedCrossAttnParams):
    """context_parallel.py:227-254."""
    seq_len, N, D = x.shape
    assert seq_len == rope.size(0), f"seq_len: {seq_len} != rope.size(0): {rope.size(0)}"
    assert condition_map.size(0) == seq_len, f"condition_map.size(0): {condition_map.size(0)} != seq_len: {seq_len}"
    if xattn_mask_for_cuda_graph is not None:
        raise NotImplementedError("xattn_mask_for_cuda_graph is hard-wired to None in the reference (dit_model.py:210)")
    cp_split_sizes = ulysses_split_sizes(seq_len, cp_size)
    x = scatter_to_context_parallel_region(x, cp_split_sizes)
    condition_map = scatter_to_context_parallel_region(condition_map, cp_split_sizes)
    rope = scatter_to_context_parallel_region(rope, cp_split_sizes)
    cross_attn_params = cp_update_cross_attn_qkv_range(cross_attn_params, N, cp_split_sizes, x.device)
    return x, condition_map, rope, cp_split_sizes, cross_attn_params


def cp_shuffle_overlap_process(cp_size: int, x: torch.Tensor, condition_map: torch.Tensor, rope: torch.Tensor,
                               xattn_mask_for_cuda_graph: Union[torch.Tensor, None], ardf_meta: dict,
                               core_attn_params: PackedCoreAttnParams, cross_attn_params: PackedCrossAttnParams):
    """context_parallel.py:257-305: every denoising chunk is padded to a multiple of cp and split across the ranks on
    its own (load balance under the chunk-causal mask); q ranges are rescaled to the padded chunk length, k ranges
    stay in unpadded coordinates (the key/value sequence is un-padded again after its all-to-all)."""
    seq_len, N, D = x.shape
    assert seq_len == rope.size(0), f"seq_len: {seq_len} != rope.size(0): {rope.size(0)}"
    assert condition_map.size(0) == seq_len, f"condition_map.size(0): {condition_map.size(0)} != seq_len: {seq_len}"
    if xattn_mask_for_cuda_graph is not None:
        raise NotImplementedError("xattn_mask_for_cuda_graph is hard-wired to None in the reference (dit_model.py:210)")
    cp_shuffle_num = ardf_meta["denoising_range_num"]
    chunk = divide(seq_len, cp_shuffle_num)
    cp_pad_size = 0
    if chunk % cp_size != 0:
        cp_pad_size = (cp_size - chunk % cp_size) * cp_shuffle_num
    cp_split_sizes = [(seq_len + cp_pad_size) // cp_size] * cp_size
    x = scatter_to_context_parallel_region(x, cp_split_sizes, cp_shuffle_num, cp_pad_size)
    condition_map = scatter_to_context_parallel_region(condition_map, cp_split_sizes, cp_shuffle_num, cp_pad_size)
    rope = scatter_to_context_parallel_region(rope, cp_split_sizes, cp_shuffle_num, cp_pad_size)
    g = math.gcd(seq_len, seq_len + cp_pad_size)
    _sq, _psq = seq_len // g, (seq_len + cp_pad_size) // g
    q_range = ardf_meta["q_range"] * _psq // _sq
    np_q_range = (core_attn_params.np_q_range.astype(np.int64) * _psq // _sq).astype(np.int32) \
        if core_attn_params is not None and core_attn_params.np_q_range is not None else q_range.cpu().numpy()
    np_k_range = core_attn_params.np_k_range if core_attn_params is not None and \
        core_attn_params.np_k_range is not None else ardf_meta["k_range"].cpu().numpy()
    core_attn_params = PackedCoreAttnParams(
        q_range=q_range, k_range=ardf_meta["k_range"], np_q_range=np_q_range, np_k_range=np_k_range,
        max_seqlen_q=ardf_meta["max_seqlen_q"] * _psq // _sq, max_seqlen_k=ardf_meta["max_seqlen_k"])
    cross_attn_params = cp_update_cross_attn_qkv_range(cross_attn_params, N, cp_split_sizes, x.device, cp_shuffle_num,
                                                       cp_pad_size)
    return x, condition_map, rope, cp_pad_size, cp_split_sizes, core_attn_params, cross_attn_params


def cp_pre_process(cp_size: int, cp_strategy: str, x: torch.Tensor, condition_map: torch.Tensor, rope: torch.Tensor,
                   xattn_mask_for_cuda_graph, ardf_meta: dict, core_attn_params: PackedCoreAttnParams,
                   cross_attn_params: PackedCrossAttnParams):
    """context_parallel.py:308-361. Returns (x, condition_map, rope, cp_pad_size, cp_split_sizes, core_attn_params,
    cross_attn_params). `cp_ring_kv` (new) stripes every chunk across the ranks (`cp_ring_kv_process`)."""
    if cp_size == 1:
        return x, condition_map, rope, None, None, core_attn_params, cross_attn_params
    if cp_strategy == "cp_ring_kv":
        x, condition_map, rope, cp_split_sizes, cross_attn_params = cp_ring_kv_process(
            cp_size, x, condition_map, rope, ardf_meta, cross_attn_params)
        return x, condition_map, rope, 0, cp_split_sizes, core_attn_params, cross_attn_params
    if cp_strategy == "cp_ulysses":
        x, condition_map, rope, cp_split_sizes, cross_attn_params = cp_ulysses_process(
            cp_size, x, condition_map, rope, xattn_mask_for_cuda_graph, cross_attn_params)
        return x, condition_map, rope, 0, cp_split_sizes, core_attn_params, cross_attn_params
    if cp_strategy == "cp_shuffle_overlap":
        return cp_shuffle_overlap_process(cp_size, x, condition_map, rope, xattn_mask_for_cuda_graph, ardf_meta,
                                          core_attn_params, cross_attn_params)
    raise ValueError(f"Invalid CP strategy: {cp_strategy}, expected cp_ulysses, cp_shuffle_overlap or cp_ring_kv")


def cp_post_process(cp_size: int, cp_strategy: str, x: torch.Tensor, meta_args: ModelMetaArgs) -> torch.Tensor:
    """context_parallel.py:364-375."""
    if cp_size == 1:
        return x
    if cp_strategy == "cp_shuffle_overlap":
        return gather_from_context_parallel_region(x, meta_args.cp_split_sizes, meta_args.denoising_range_num,
                                                   meta_args.cp_pad_size)
    if cp_strategy == "cp_ring_kv":
        return gather_ring_kv(x, meta_args.cp_split_sizes, meta_args.denoising_range_num)
    if cp_strategy == "cp_ulysses":
        return gather_from_context_parallel_region(x, meta_args.cp_split_sizes)
    raise ValueError(f"Invalid CP strategy: {cp_strategy}, expected cp_ulysses, cp_shuffle_overlap or cp_ring_kv")


#####################################################
# Ulysses all-to-all
#####################################################
def _kernel_relayout_in(tensor: torch.Tensor, cp: int, rep: int) -> torch.Tensor:
    """Default send-side re-layout when the caller injects none: the CUDA kernel behind magi_ulysses_relayout (mode 0).
    Imported lazily - the CPU (gloo) tests inject torch doubles and never load the library."""
    from . import functional as F
    return F.ulysses_relayout(tensor.contiguous(), cp, 0, kv_repeat=rep)


def _kernel_relayout_out(recv: torch.Tensor, cp: int, result: torch.Tensor, q_head: int, lo: int) -> None:
    """Default receive-side re-layout (mode 1), writing heads [lo, lo + hn) of every cp block of `result` in place."""
    from . import functional as F
    F.ulysses_relayout(recv, cp, 1, out=result, dst_heads_total=q_head, dst_head_offset=lo)


def all_to_all_input_split(tensor: torch.Tensor, cp_split_sizes: List[int], relayout: Optional[Callable] = None,
                           send_layout: bool = False, out: Optional[torch.Tensor] = None):
    """context_parallel.py:381-401: scatter heads / gather sequence. tensor (seq, cp*hn, hd) -> (sum(sizes), hn, hd).
    send_layout=True means `tensor` already is "(cp seq) hn hd" (written by the prologue kernel): nothing is staged.
    Otherwise `relayout(tensor, cp, kv_repeat)` must produce it. `out` lets the result land in a caller-owned
    contiguous buffer (e.g. the KV-cache slot)."""
    cp = mpu.get_cp_world_size()
    if cp == 1:
        return tensor, FakeHandle()
    assert cp_split_sizes is not None
    if not send_layout:
        _, hn, _ = tensor.shape
        rep = cp // hn if (cp % hn == 0 and cp != hn) else 1
        tensor = (relayout or _kernel_relayout_in)(tensor, cp, rep)
    assert tensor.is_contiguous()
    if out is None:
        out = torch.empty([sum(cp_split_sizes), *tensor.shape[1:]], device=tensor.device, dtype=tensor.dtype)
    else:
        assert out.is_contiguous() and out.shape[0] == sum(cp_split_sizes) and out.shape[1:] == tensor.shape[1:]
    handle = dist.all_to_all_single(out, tensor, output_split_sizes=list(cp_split_sizes), group=mpu.get_cp_group(),
                                    async_op=True)
    return out, handle


def all_to_all_output_split(tensor: torch.Tensor, cp_split_sizes: List[int]):
    """context_parallel.py:404-423: scatter sequence / gather heads. (sum(sizes), hn, hd) -> (cp * local, hn, hd)."""
    cp = mpu.get_cp_world_size()
    if cp == 1:
        return tensor, FakeHandle()
    assert cp_split_sizes is not None and tensor.is_contiguous()
    out = torch.empty([cp_split_sizes[mpu.get_cp_rank()] * cp, *tensor.shape[1:]], device=tensor.device,
                      dtype=tensor.dtype)
    handle = dist.all_to_all_single(out, tensor, input_split_sizes=list(cp_split_sizes), group=mpu.get_cp_group(),
                                    async_op=True)
    return out, handle


def fused_qkv_communication(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cp_split_sizes: List[int],
                            relayout: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """context_parallel.py:426-449: ONE all-to-all for q, k and v (fewer host operations when the sequence is very
    short). q (seq, cp*hn_q, hd), k / v (seq, Hkv, hd) -> (sum(sizes), hn_q, hd), (sum(sizes), hn_kv, hd) x 2, where
    k and v come back as views of one packed (sum(sizes), hn_kv, 2*hd) = [K | V] buffer - the KV-cache / kernel
    layout, so the reference's `torch.cat([k, v], -1)` behind it (context_parallel.py:546) is free.
    `relayout(tensor, cp, kv_repeat)` produces the "(cp seq) hn hd" send layout (incl. KV-head replication)."""
    cp = mpu.get_cp_world_size()
    if cp == 1:
        return q, k, v
    assert cp_split_sizes is not None
    k_head = k.shape[1]
    rep = cp // k_head if (cp % k_head == 0 and cp != k_head) else 1
    relayout = relayout or _kernel_relayout_in
    qs, ks, vs = relayout(q, cp, 1), relayout(k, cp, rep), relayout(v, cp, rep)
    hq, hk, hd = qs.shape[1], ks.shape[1], qs.shape[2]
    # wire format per row: [q heads | (K | V) per kv head], one contiguous message per destination rank
    send = torch.empty((qs.shape[0], (hq + 2 * hk) * hd), device=q.device, dtype=q.dtype)
    send[:, :hq * hd].view(-1, hq, hd).copy_(qs)
    packed = send[:, hq * hd:].view(-1, hk, 2 * hd)
    packed[..., :hd].copy_(ks)
    packed[..., hd:].copy_(vs)
    recv = torch.empty((sum(cp_split_sizes), send.shape[1]), device=q.device, dtype=q.dtype)
    dist.all_to_all_single(recv, send, output_split_sizes=list(cp_split_sizes), group=mpu.get_cp_group())
    kv = recv[:, hq * hd:].view(-1, hk, 2 * hd)
    return recv[:, :hq * hd].view(-1, hq, hd), kv[..., :hd], kv[..., hd:]


def head_groups(q_heads: int, kv_heads: int, overlap_degree: int) -> List[Tuple[int, int, int, int]]:
    """Split the local heads into `overlap_degree` CONTIGUOUS groups (q_lo, q_hi, kv_lo, kv_hi) for the
    attention / all-to-all pipeline of context_parallel.py:565-589. overlap_degree == -1 selects q_heads // kv_heads.
    Unlike the reference's GQA branch (context_parallel.py:574-578), whose reshape/chunk permutes the head order of
    the concatenated output when kv_heads > 1, groups here are contiguous head slices, so the result is identical to
    overlap_degree == 1 for every degree."""
    if overlap_degree == -1:
        overlap_degree = q_heads // kv_heads
    if overlap_degree < 1 or overlap_degree > q_heads:
        raise ValueError(f"overlap_degree {overlap_degree} out of range for {q_heads} heads")
    g = q_heads // kv_heads
    if overlap_degree == 1:
        return [(0, q_heads, 0, kv_heads)]
    if kv_heads % overlap_degree == 0:   # whole kv heads per group
        kpg = kv_heads // overlap_degree
        return [(i * kpg * g, (i + 1) * kpg * g, i * kpg, (i + 1) * kpg) for i in range(overlap_degree)]
    if overlap_degree % kv_heads == 0 and g % (overlap_degree // kv_heads) == 0:  # several groups per kv head
        per = overlap_degree // kv_heads
        qpg = g // per
        return [(i * qpg, (i + 1) * qpg, i // per, i // per + 1) for i in range(overlap_degree)]
    raise ValueError(f"cannot split {q_heads} q heads / {kv_heads} kv heads into {overlap_degree} contiguous groups")


class UlyssesScheduler:
    """context_parallel.py:452-597, callable-based like the reference. Differences: see module docstring."""

    @staticmethod
    def get_attn_and_xattn_with_fused_kv_comm(get_q_func: Callable, get_kv_func: Callable, kv_cache_func: Callable,
                                              core_attn_func: Callable, cross_attn_func: Callable,
                                              overlap_degree: int, batch_size: int, cp_size: int,
                                              cp_split_sizes: List[int] = None, *, relayout_in: Callable = None,
                                              relayout_out: Callable = None, send_layout: bool = False,
                                              kv_recv_buffer: Optional[torch.Tensor] = None):
        """context_parallel.py:503-527. get_kv_func() -> packed kv; get_q_func() -> q; kv_cache_func(kv) -> (key,
        value) views; core_attn_func(q_group, key_group, value_group, out=..., kv_heads=(lo, hi)) -> out;
        cross_attn_func() -> xattn_out."""
        kv = get_kv_func()
        kv, handle_kv = all_to_all_input_split(kv, cp_split_sizes, relayout_in, send_layout, out=kv_recv_buffer)
        query = get_q_func()
        query, handle_q = all_to_all_input_split(query, cp_split_sizes, relayout_in, send_layout)
        handle_kv.wait()
        key, value = kv_cache_func(kv)
        handle_q.wait()
        return UlyssesScheduler.get_attn_and_xattn_base(query, key, value, core_attn_func, cross_attn_func,
                                                        overlap_degree, batch_size, cp_size, cp_split_sizes,
                                                        relayout_out=relayout_out)

    @staticmethod
    def get_attn_and_xattn_with_comm_overlap(get_q_func: Callable, get_k_func: Callable, get_v_func: Callable,
                                             kv_cache_func: Callable, core_attn_func: Callable,
                                             cross_attn_func: Callable, overlap_degree: int, batch_size: int,
                                             cp_size: int, cp_split_sizes: List[int] = None, *,
                                             relayout_in: Callable = None, relayout_out: Callable = None):
        """context_parallel.py:457-501: v, k and q each start their own all-to-all as soon as they are computed
        (v_comm under k_compute, k_comm under q_compute, q_comm under the cache update). kv_cache_func receives the
        packed [K | V] tensor like in the reference (`torch.concat([key, value], -1)`, :493)."""
        value = get_v_func()
        value, handle_v = all_to_all_input_split(value, cp_split_sizes, relayout_in)
        key = get_k_func()
        key, handle_k = all_to_all_input_split(key, cp_split_sizes, relayout_in)
        query = get_q_func()
        query, handle_q = all_to_all_input_split(query, cp_split_sizes, relayout_in)
        handle_v.wait()
        handle_k.wait()
        key, value = kv_cache_func(torch.cat([key, value], dim=-1))
        handle_q.wait()
        return UlyssesScheduler.get_attn_and_xattn_base(query, key, value, core_attn_func, cross_attn_func,
                                                        overlap_degree, batch_size, cp_size, cp_split_sizes,
                                                        relayout_out=relayout_out)

    @staticmethod
    def get_attn_and_xattn_with_fused_qkv_comm(get_qkv_func: Callable, kv_cache_func: Callable,
                                               core_attn_func: Callable, cross_attn_func: Callable,
                                               overlap_degree: int, batch_size: int, cp_size: int,
                                               cp_split_sizes: List[int] = None, *, relayout_in: Callable = None,
                                               relayout_out: Callable = None):
        """context_parallel.py:529-548: q, k and v travel in ONE all-to-all (the reference takes this branch for
        CUDA-graph runs with <= 3 denoising ranges, dit_module.py:1182-1193). get_qkv_func() -> (q, k, v)."""
        q, k, v = get_qkv_func()
        q, k, v = fused_qkv_communication(q, k, v, cp_split_sizes, relayout_in)
        if k.shape[-1] == v.shape[-1] and k.data_ptr() + k.shape[-1] * k.element_size() == v.data_ptr() \
                and k.stride() == v.stride():
            kv = k.as_strided((k.shape[0], k.shape[1], 2 * k.shape[2]), k.stride(), k.storage_offset())  # already packed
        else:
            kv = torch.cat([k, v], dim=-1)
        k, v = kv_cache_func(kv)
        return UlyssesScheduler.get_attn_and_xattn_base(q, k, v, core_attn_func, cross_attn_func, overlap_degree,
                                                        batch_size, cp_size, cp_split_sizes,
                                                        relayout_out=relayout_out)

    @staticmethod
    def get_attn_and_xattn_base(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, core_attn_func: Callable,
                                cross_attn_func: Callable, overlap_degree: int, batch_size: int, cp_size: int,
                                cp_split_sizes: List[int] = None, *, relayout_out: Callable = None):
        """context_parallel.py:550-597: attention of head group g overlaps the output all-to-all of group g-1; the
        cross-attention (local tokens, all heads, no communication) overlaps the last one."""
        q_seq, q_head, q_hidden = query.shape
        kv_head = key.shape[1]
        relayout_out = relayout_out or _kernel_relayout_out
        groups = head_groups(q_head, kv_head, overlap_degree)
        local = cp_split_sizes[mpu.get_cp_rank()] if cp_split_sizes is not None else q_seq
        result = torch.empty((local, cp_size * q_head, q_hidden), dtype=query.dtype, device=query.device)
        pending = None
        for (q_lo, q_hi, k_lo, k_hi) in groups:
            o = torch.empty((q_seq, q_hi - q_lo, q_hidden), dtype=query.dtype, device=query.device)
            core_attn_func(query[:, q_lo:q_hi], key[:, k_lo:k_hi], value[:, k_lo:k_hi], out=o, kv_heads=(k_lo, k_hi))
            if pending is not None:
                recv, handle, lo = pending
                handle.wait()
                relayout_out(recv, cp_size, result, q_head, lo)
            recv, handle = all_to_all_output_split(o, cp_split_sizes)
            pending = (recv, handle, q_lo)
        xattn_out = cross_attn_func()
        recv, handle, lo = pending
        handle.wait()
        relayout_out(recv, cp_size, result, q_head, lo)
        # "(cp sq b) hn hd -> sq b (cp hn hd)"
        core_attn_out = result.view(local // batch_size, batch_size, cp_size * q_head * q_hidden)
        return core_attn_out, xattn_out


#####################################################
# Ulysses with the all-to-alls fused into the kernels (NVLink peer memory)
#####################################################
class UlyssesPeerWorkspace:
    """Receive buffers of the fused Ulysses path, allocated in symmetric memory (torch.distributed._symmetric_memory:
    the same virtual layout on every rank of the group, peer-mapped over NVLink) and shared by all layers:

        q   (total rows, Hq / cp, 128)     written by every rank's prologue kernel   (= all_to_all_input_split)
        kv  (total rows, Hkv / cp, 256)    written by every rank's prologue kernel
        out (max local rows, Hq, 128)      written by every rank's attention epilogue
                                           (= all_to_all_output_split + "(cp sq) hn hd -> sq (cp hn) hd")

    Two device-side barriers per layer order the exchange: A after the prologue (all q / kv rows have landed, and
    every rank has consumed the previous layer's `out`), B after the attention (all `out` rows have landed, and every
    rank is done reading q / kv). torch provides allocation, rendezvous and the barrier (plumbing); the data
    movement is the kernels' own stores."""

    _instances = {}

    def __init__(self, group, total_rows: int, max_local_rows: int, hq_loc: int, hkv_loc: int, heads_q: int, device):
        import torch.distributed._symmetric_memory as symm
        self.group = group

        def alloc(shape):
            t = symm.empty(shape, dtype=torch.bfloat16, device=device)
            return t, symm.rendezvous(t, group.group_name)

        self.q, self.h_q = alloc((total_rows, hq_loc, 128))
        self.kv, self.h_kv = alloc((total_rows, hkv_loc, 256))
        self.out, self.h_out = alloc((max_local_rows, heads_q, 128))
        self.q_ptrs = [int(p) for p in self.h_q.buffer_ptrs]
        self.kv_ptrs = [int(p) for p in self.h_kv.buffer_ptrs]
        self.out_ptrs = [int(p) for p in self.h_out.buffer_ptrs]
        self._alloc, self._proj_shape, self.proj, self.proj_ptrs, self.proj_calls = alloc, (max_local_rows, 2 * heads_q * 128), None, None, 0

    def next_proj(self):
        """(buffer, peer pointers) of this call's out-projection input "sq (hn n hd)" (fused_proj_input under the fused
        Ulysses path): every rank's attention epilogue stores its heads' self-attention rows into slot n = 0 of the
        OWNER's buffer, the local cross-attention fills slot n = 1. Two buffers used alternately, allocated on first
        use (collective: every rank makes the same sequence of calls): the tensor handed to the caller stays valid
        until the call after next - a peer writes buffer b again only after barrier A of that call, which this rank
        enters after everything it issued in between - so no private copy is made."""
        if self.proj is None:
            pairs = [self._alloc(self._proj_shape) for _ in range(2)]
            self.proj = [t for t, _ in pairs]
            self._h_proj = [h for _, h in pairs]
            self.proj_ptrs = [[int(p) for p in h.buffer_ptrs] for h in self._h_proj]
        b = self.proj_calls % 2
        self.proj_calls += 1
        return self.proj[b], self.proj_ptrs[b]

    def barrier(self) -> None:
        """All ranks of the group, on the current stream, no host synchronisation."""
        self.h_q.barrier(channel=0)

    @classmethod
    def get(cls, group, total_rows, max_local_rows, hq_loc, hkv_loc, heads_q, device):
        key = (id(group), total_rows, max_local_rows, hq_loc, hkv_loc, heads_q, str(device))
        ws = cls._instances.get(key)
        if ws is None:     # collective: every rank reaches this with the same key at the same forward
            ws = cls._instances[key] = cls(group, total_rows, max_local_rows, hq_loc, hkv_loc, heads_q, device)
        return ws

    _probe = {}   # id(group) -> bool: result of the one-time symmetric-memory probe

    @staticmethod
    def available(group, device) -> bool:
        """True when the fused (peer-store) paths can run on this group: NCCL backend, at most MAGI_MAX_PEERS ranks
        (the C-ABI's peer tables), and a symmetric-memory rendezvous that actually succeeds - probed ONCE per group
        with a tiny buffer and agreed on by all ranks (a group without P2P / NVLink access, e.g. spanning nodes, fails
        the rendezvous: the callers then fall back to the NCCL paths instead of crashing). Collective on first use."""
        if group is None or not dist.is_initialized() or torch.device(device).type != "cuda":
            return False
        key = id(group)
        ok = UlyssesPeerWorkspace._probe.get(key)
        if ok is not None:
            return ok
        ok = False
        try:
            from ._lib import MAX_PEERS
            if dist.get_backend(group) == "nccl" and dist.get_world_size(group) <= MAX_PEERS:
                import torch.distributed._symmetric_memory as symm
                mine = 1
                try:
                    t = symm.empty((256,), dtype=torch.bfloat16, device=device)
                    h = symm.rendezvous(t, group.group_name)
                    mine = 1 if len(h.buffer_ptrs) == dist.get_world_size(group) else 0
                except Exception:
                    mine = 0
                flag = torch.tensor([mine], device=device, dtype=torch.int32)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # every rank takes the same path
                ok = bool(flag.item())
        except Exception:
            ok = False
        UlyssesPeerWorkspace._probe[key] = ok
        return ok


def fused_ulysses_plan(cp_split_sizes: Sequence[int], rank: int, heads_q: int, cp: int):
    """Routing arguments of the fused path: (row offset of this rank's tokens in the receive buffers, row bounds of
    every rank, head slot of this rank's heads in the owners' output rows)."""
    bounds = [0]
    for s_ in cp_split_sizes:
        bounds.append(bounds[-1] + int(s_))
    return bounds[rank], bounds, rank * divide(heads_q, cp)


#####################################################
# CSO (context shuffle overlap) attention pipeline
#####################################################
def cso_communication(input: torch.Tensor, cp_world_size: int, cp_split_sizes: List[int], comm_type: str = None, *,
                      relayout: Optional[Callable] = None, send_layout: bool = False):
    """context_parallel.py:603-621: equal-split all-to-all along dim 0. comm_type "kv": the packed K/V tensor
    "spb (cp hn) hd" is first brought to send layout "(cp spb) hn hd" (with KV-head replication when cp > hn) by
    `relayout(tensor, cp, rep)`, or already is (send_layout=True: the prologue kernel wrote it that way)."""
    if cp_world_size == 1:
        return input, FakeHandle()
    assert cp_split_sizes is not None
    if comm_type == "kv" and not send_layout:
        _, hn, _ = input.shape
        rep = divide(cp_world_size, hn) if (cp_world_size % hn == 0 and cp_world_size != hn) else 1
        input = relayout(input, cp_world_size, rep)
    assert input.is_contiguous()
    output = torch.empty(input.shape, device=input.device, dtype=input.dtype)
    handle = dist.all_to_all_single(output, input, input_split_sizes=list(cp_split_sizes), group=mpu.get_cp_group(),
                                    async_op=True)
    return output, handle


def cso_unshuffle_rows(cp: int, dn: int, spb: int, clip: int) -> np.ndarray:
    """Row gather of "(cp dn spb) -> dn (cp spb)" followed by [:, :clip] (dit_module.py:1215-1224): source row of every
    row of the un-padded, chunk-ordered key/value sequence inside the all-to-all receive buffer."""
    t = np.arange(clip, dtype=np.int64)
    c, s = t // spb, t % spb
    d = np.arange(dn, dtype=np.int64)[:, None]
    return ((c[None, :] * dn + d) * spb + s[None, :]).reshape(-1)


class CSOHelper:
    """context_parallel.py:624-664. Chunk i's attention runs while chunk i+1's queries arrive and chunk i-1's outputs
    leave. The reference concatenates q(i+1) and o(i-1) into one message and splits it again on arrival (two copies
    per step); here they are two all-to-alls issued back to back, which on NVSwitch costs one more launch and no
    copy. Entry points, argument meaning and the returned (core_attn_outs, handle_attn) are the reference's."""

    def __init__(self, cp_shuffle_num, cp_world_size, cp_split_sizes):
        self.cp_shuffle_num = cp_shuffle_num
        self.cp_world_size = cp_world_size
        self.cp_split_sizes = [divide(x, self.cp_shuffle_num) for x in cp_split_sizes]

    def split_query_for_overlap(self, query, *, relayout: Optional[Callable] = None, send_layout: bool = False):
        """query "(dn spb) (cp hn) hd" -> dn send buffers "(cp spb) hn hd"; chunk 0's all-to-all is started.
        send_layout=True: `query` already is "(dn cp spb) hn hd" (written that way by the prologue kernel)."""
        dn, cp = self.cp_shuffle_num, self.cp_world_size
        if not send_layout and cp > 1:
            t, h, d = query.shape
            if relayout is not None:
                query = torch.cat([relayout(c, cp, 1) for c in torch.chunk(query, dn, dim=0)], dim=0)
            else:
                query = query.view(dn, t // dn, cp, h // cp, d).permute(0, 2, 1, 3, 4).reshape(t * cp, h // cp, d)
        querys = list(torch.chunk(query, dn, dim=0))
        querys[0], handle_q = cso_communication(querys[0], cp, self.cp_split_sizes)
        return querys, handle_q

    def overlap(self, fattn, qs, k, v):
        """fattn(q, k, v, i) -> o. Returns ([o_0 .. o_{dn-1}] each "(cp sq b) hn hd", handle of the last one)."""
        dn, cp = self.cp_shuffle_num, self.cp_world_size
        outs = []
        pending_q = pending_o = None
        o = None
        for i in range(dn):
            q = qs[0] if i == 0 else None
            if i > 0:
                recv_q, h = pending_q
                h.wait()
                q = recv_q
            # issue before this chunk's attention: queries of the next chunk, outputs of the previous one
            if i + 1 < dn:
                pending_q = cso_communication(qs[i + 1], cp, self.cp_split_sizes)
            if pending_o is not None:
                pending_o[1].wait()
            if i > 0:
                pending_o = cso_communication(o, cp, self.cp_split_sizes)
                outs.append(pending_o[0])
            o = fattn(q, k, v, i)
        if pending_o is not None:
            pending_o[1].wait()
        last_o, handle_attn = cso_communication(o, cp, self.cp_split_sizes)
        outs.append(last_o)
        return outs, handle_attn


#####################################################
# Ring-KV exchange (new capability)
#####################################################
def ring_split_sizes(clip: int, cp_size: int) -> List[int]:
    """Rows of EVERY chunk a rank holds under cp_ring_kv: clip // cp (+1 on the first clip % cp ranks). Striping
    every chunk across the ranks (instead of cutting the flattened sequence into cp contiguous pieces) gives every
    rank the same share of every (q-range x k-range) rectangle: under the chunk-causal mask a contiguous split leaves
    the rank holding the last chunk with several times the work of the one holding the first (measured on 8 x B200,
    24B / 720p state: 4724 vs 8870 TFLOP/s for Ulysses; profiles/README.md)."""
    return ulysses_split_sizes(clip, cp_size)


def cp_ring_kv_process(cp_size: int, x: torch.Tensor, condition_map: torch.Tensor, rope: torch.Tensor, ardf_meta: dict,
                       cross_attn_params: PackedCrossAttnParams):
    """Pre-processing of cp_ring_kv: per-chunk striped scatter (no padding: uneven slices), cross-attention ranges of
    the local slices. Local row order is (chunk, slice row). Returns (x, condition_map, rope, cp_split_sizes,
    cross_attn_params) with cp_split_sizes[r] = dn * (slice rows of rank r), what the exit all-gather needs."""
    seq_len, N, _ = x.shape
    dn = ardf_meta["denoising_range_num"]
    if N > 1:
        # the batched uncond pass (dit_model.py:413-494: chunks moved to the batch dim, one range, no KV cache): with
        # one chunk per batch row the stripes ARE the contiguous Ulysses shards, and the attention block runs it as
        # cp_ulysses (each row attends only itself, nothing to gather across chunks)
        assert dn == 1, "cp_ring_kv with batch size > 1 needs denoising_range_num == 1 (dit_module.py:1048)"
        return cp_ulysses_process(cp_size, x, condition_map, rope, None, cross_attn_params)
    clip = divide(seq_len, dn)
    sizes = ring_split_sizes(clip, cp_size)
    rank = mpu.get_cp_rank()
    lo, n = sum(sizes[:rank]), sizes[rank]

    def stripe(t):
        return t.unflatten(0, (dn, clip))[:, lo:lo + n].flatten(0, 1).contiguous()

    x, condition_map, rope = stripe(x), stripe(condition_map), stripe(rope)
    # every local chunk slice is one query segment against that chunk's caption (dit_model.py:215-259)
    kv = cross_attn_params.kv_ranges
    cu_k = cross_attn_params.cu_seqlens_kv
    q_ranges = torch.arange(dn + 1, dtype=torch.int32, device=kv.device) * n
    cross = PackedCrossAttnParams(q_ranges=torch.stack([q_ranges[:-1], q_ranges[1:]], 1), kv_ranges=kv,
                                  cu_seqlens_q=q_ranges, cu_seqlens_kv=cu_k, max_seqlen_q=n,
                                  max_seqlen_kv=cross_attn_params.max_seqlen_kv)
    return x, condition_map, rope, [dn * s_ for s_ in sizes], cross


def gather_ring_kv(input_: torch.Tensor, cp_split_sizes: List[int], dn: int) -> torch.Tensor:
    """Exit all-gather of cp_ring_kv: stripes padded to the widest one (equal-sized messages),
    then "(cp dn n_r) -> dn (cp n_r)" with the padding dropped."""
    world = mpu.get_cp_world_size()
    if world == 1:
        return input_
    sizes = [divide(s_, dn) for s_ in cp_split_sizes]
    n_max = max(sizes)
    tail = tuple(input_.shape[1:])
    mine = input_.unflatten(0, (dn, input_.shape[0] // dn))
    if mine.shape[1] < n_max:
        mine = torch.nn.functional.pad(mine, [0, 0] * len(tail) + [0, n_max - mine.shape[1]])
    output = torch.empty((world, dn, n_max) + tail, dtype=input_.dtype, device=input_.device)
    dist.all_gather(list(output.unbind(0)), mine.contiguous(), group=mpu.get_cp_group())
    return torch.cat([output[r, :, :sizes[r]] for r in range(world)], dim=1).flatten(0, 1)


class RingGatherWorkspace:
    """In-flight K/V buffers of the fused cp_ring_kv path, in symmetric memory (same virtual layout on every rank,
    peer-mapped over NVLink, plus an NVSwitch multicast mapping when the fabric offers one):

        kv[2]  (max in-flight rows, Hkv, 256)   every rank's prologue kernel stores its stripe of every chunk into
                                                EVERY rank's buffer, chunk-major (= the single-GPU row order)

    The cp - 1 ring steps of K/V blocks, their attention launches and log-sum-exp merges collapse into: one prologue
    launch whose K/V stores are the all-gather, ONE device-side barrier, one attention launch per rank over its query
    stripe. Two buffers used alternately make one barrier per layer enough: a rank that writes buffer b for call n has
    passed the barrier of call n - 1, which every rank enters only after its attention of call n - 2 (the previous
    reader of b) in stream order."""

    _instances = {}

    def __init__(self, group, rows: int, heads_kv: int, device):
        import torch.distributed._symmetric_memory as symm
        self.group, self.rows, self.calls = group, rows, 0
        self.kv, self.handles, self.ptrs, self.mc_ptrs = [], [], [], []
        for _ in range(2):
            t = symm.empty((rows, heads_kv, 256), dtype=torch.bfloat16, device=device)
            h = symm.rendezvous(t, group.group_name)
            self.kv.append(t)
            self.handles.append(h)
            self.ptrs.append([int(p) for p in h.buffer_ptrs])
            self.mc_ptrs.append(int(getattr(h, "multicast_ptr", 0) or 0))

    def next(self):
        """(buffer, peer pointers, multicast pointer or 0) for this call; every rank makes the same sequence of calls."""
        b = self.calls % 2
        self.calls += 1
        return self.kv[b], self.ptrs[b], self.mc_ptrs[b]

    def barrier(self) -> None:
        self.handles[0].barrier(channel=0)

    @classmethod
    def get(cls, group, rows, heads_kv, device):
        key = (id(group), heads_kv, str(device))
        ws = cls._instances.get(key)
        if ws is None or ws.rows < rows:   # collective: every rank reaches this with the same key at the same forward
            ws = cls._instances[key] = cls(group, rows, heads_kv, device)
        return ws

    available = staticmethod(UlyssesPeerWorkspace.available)


def ring_multicast_enabled() -> bool:
    """MAGI_RING_MULTICAST=0 keeps the per-peer stores even when the fabric offers an NVSwitch multicast mapping."""
    import os
    return os.environ.get("MAGI_RING_MULTICAST", "1") != "0"


def ring_step_ranges(chunk_ranges: Sequence[Tuple[int, int]], slice_point: int, clip: int, n_q: int, n_blk: int,
                     with_cache: bool):
    """Host-side interval arithmetic for one ring step under the striped sharding.

    chunk_ranges[i] = (s_i, e_i): query chunk i attends absolute chunks [s_i, e_i) (dit_module.py:914-931 coordinates
    divided by clip); chunks below slice_point live in the (replicated) KV cache, the in-flight chunk j is absolute
    chunk slice_point + j. This rank's queries of chunk i are local rows [i * n_q, (i + 1) * n_q); the visiting block
    holds, for every in-flight chunk j, rows [j * n_blk, (j + 1) * n_blk) - chunk order, so the in-flight part of a
    range is ONE contiguous interval of the block. Returns (q_ranges, k_ranges) with k rows in the coordinates of
    concat(cache[: slice_point * clip] if with_cache else [], block). The cached part of a range is attended exactly
    once, in the step that has with_cache (it is contiguous with the block part, which then starts at block row 0).
    Rectangles with an empty key interval are dropped (their rows keep out = 0, lse = -inf for the merge)."""
    qs, ks = [], []
    base = slice_point * clip if with_cache else 0
    for i, (s_i, e_i) in enumerate(chunk_ranges):
        b0, b1 = max(s_i, slice_point) - slice_point, max(e_i, slice_point) - slice_point   # in-flight chunks
        k0, k1 = base + b0 * n_blk, base + b1 * n_blk
        if with_cache and s_i < slice_point:
            k0 = s_i * clip
            if e_i <= slice_point:
                k1 = e_i * clip
        if k0 >= k1 or n_q == 0:
            continue
        qs.append([i * n_q, (i + 1) * n_q])
        ks.append([k0, k1])
    return (np.asarray(qs, dtype=np.int32).reshape(-1, 2), np.asarray(ks, dtype=np.int32).reshape(-1, 2))


class RingKVScheduler:
    """Sequence-sharded attention: every rank keeps a stripe of every query chunk with ALL heads; the packed K/V blocks
    of the new tokens travel around the ring (cp - 1 point-to-point steps, next transfer in flight while the current
    block is attended), partial results are merged with their log-sum-exp. The KV cache of clean chunks is kept
    replicated (full heads, full sequence, chunk-major) on every rank: each visiting block is scattered into it as it
    passes, which at 24B / 720p / 8 chunks is 34 GB of the 180 GB per GPU and means cached chunks are never
    re-communicated. Oracle: the single-GPU result; the reference has no ring attention."""

    @staticmethod
    def run(query: torch.Tensor, kv_local: torch.Tensor, cache: Optional[torch.Tensor],
            chunk_ranges: Sequence[Tuple[int, int]], slice_point: int, clip: int, dn: int, sizes: Sequence[int],
            cache_chunks: int, attn_func: Callable, merge_func: Callable):
        """query (dn * n_me, Hq, 128); kv_local packed (dn * n_me, Hkv, 256); cache packed (max_seq, Hkv, 256) or None.
        sizes[r]: rows of every chunk on rank r. cache_chunks: leading in-flight chunks that go to the cache (0 = no
        update). attn_func(q, kv_block, cache_prefix_or_None, q_ranges, k_ranges) -> (out, lse) with uncovered rows
        0 / -inf; merge_func(out_a, lse_a, out_b, lse_b) merges in place into a."""
        cp, rank, group = mpu.get_cp_world_size(), mpu.get_cp_rank(), mpu.get_cp_group()
        offs = np.concatenate([[0], np.cumsum(np.asarray(sizes, dtype=np.int64))]).astype(np.int64)
        n_me, n_max = int(sizes[rank]), int(max(sizes))
        nxt, prv = mpu.cp_global_rank((rank + 1) % cp), mpu.cp_global_rank((rank - 1) % cp)
        tail = tuple(kv_local.shape[1:])
        bufs = [torch.empty((dn * n_max,) + tail, dtype=kv_local.dtype, device=kv_local.device)
                for _ in range(2)] if cp > 1 else []
        cur = kv_local
        out = lse = None
        seq_start = slice_point * clip
        for step in range(cp):
            src = (rank - step) % cp
            n_blk = int(sizes[src])
            reqs = None
            if step + 1 < cp:
                # forward the block we hold to the next rank while we attend to it (fixed-size messages)
                send = cur if cur.shape[0] == dn * n_max else torch.nn.functional.pad(
                    cur[:dn * n_blk], [0, 0] * (cur.dim() - 1) + [0, dn * n_max - dn * n_blk])
                ops = [dist.P2POp(dist.isend, send.contiguous(), nxt, group=group),
                       dist.P2POp(dist.irecv, bufs[step % 2], prv, group=group)]
                reqs = dist.batch_isend_irecv(ops)
            blk = cur[:dn * n_blk]
            if cache is not None and cache_chunks > 0 and n_blk > 0:
                # replicated cache, chunk-major: chunk j of the block goes to rows [(sp + j) * clip + offs[src], + n_blk)
                dst = cache[seq_start:seq_start + cache_chunks * clip].unflatten(0, (cache_chunks, clip))
                dst[:, int(offs[src]):int(offs[src]) + n_blk].copy_(blk.unflatten(0, (dn, n_blk))[:cache_chunks])
            with_cache = cache is not None and seq_start > 0 and step == 0
            qr, kr = ring_step_ranges(chunk_ranges, slice_point, clip, n_me, n_blk, with_cache)
            if qr.shape[0] > 0:
                o, l = attn_func(query, blk, cache[:seq_start] if with_cache else None, qr, kr)
                if out is None:
                    out, lse = o, l
                else:
                    merge_func(out, lse, o, l)
            if reqs is not None:
                for r in reqs:
                    r.wait()
                cur = bufs[step % 2]
        if out is None:
            out = torch.zeros_like(query)
            lse = torch.full(query.shape[:2], float("-inf"), dtype=torch.float32, device=query.device)
        return out, lse
