"""Drop-in `FullyParallelAttention` for the reference's DiT block (inference/model/dit/dit_module.py:776-1246):
same constructor `(model_config, engine_config, layer_number)`, same sub-module / parameter names (so reference
state_dicts load: `linear_qkv.{layer_norm,q,qx,k,v}`, `linear_kv_xattn`, `linear_proj`, `q_layernorm`,
`q_layernorm_xattn`, `k_layernorm`, `k_layernorm_xattn`), same `forward(hidden_states, key_value_states,
inference_params, rotary_pos_emb, meta_args) -> (core_attn_out, xattn_out)`.

What runs where
  * LayerNorm(hidden) and the kv_xattn / out projections are plain torch / cuBLAS library calls; the q / qx / k / v
    projections run as ONE hand-written tcgen05 GEMM whose epilogue is the prologue (EngineConfig.fused_qkv_proj;
    bf16, or e4m3 under fp8_quant behind the hand-written activation quantiser), with cuBLAS as the opt-out.
  * everything between the projections and the two attention outputs is hand-written CUDA behind the C-ABI:
    fused per-head LayerNorm(+1p) + RoPE + packing + KV-cache append (one kernel instead of ~15 eager ones),
    range-masked tcgen05 attention reading the KV cache in place (no torch.cat / chunk().contiguous()),
    varlen cross-attention on the same kernel.
"""
import math
from functools import partial
from typing import Optional

import torch
import torch.nn as nn

from . import context_parallel as cpm
from . import functional as F
from . import parallel_state as mpu
from .common import EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs, PackedCrossAttnParams, divide
from .fp8_linear import PerChannelQuantizedFp8Linear, PerTensorQuantizedFp8Linear, uses_fp8


class HeadLayerNorm(nn.Module):
    """Parameter holder for the per-head LayerNorm (FusedLayerNorm, dit_module.py:323-357). The arithmetic is fused
    into the prologue kernel; `weight`/`bias` keep the reference's names, shapes and (torch.empty) initialisation."""

    def __init__(self, model_config: ModelConfig, hidden_size: int):
        super().__init__()
        self.zero_centered_gamma = model_config.apply_layernorm_1p
        self.eps = model_config.layernorm_epsilon
        self.weight = nn.Parameter(torch.empty(hidden_size, dtype=model_config.params_dtype))
        self.bias = nn.Parameter(torch.empty(hidden_size, dtype=model_config.params_dtype))


class CustomLayerNormLinear(nn.Module):
    """dit_module.py:390-425: LayerNorm(hidden) and the q / qx / k / v linears - bf16 `nn.Linear`, or
    `PerTensorQuantizedFp8Linear` on the layers `fp8_quant` covers (:407-410)."""

    def __init__(self, input_size, output_size_q, output_size_kv, layer_number, model_config, engine_config):
        super().__init__()
        self.layer_norm = nn.LayerNorm(input_size, eps=model_config.layernorm_epsilon, dtype=model_config.params_dtype)
        self.layer_number = layer_number
        self.is_fp8 = uses_fp8(engine_config, model_config, layer_number)
        for name, size in {"q": output_size_q, "qx": output_size_q, "k": output_size_kv, "v": output_size_kv}.items():
            if self.is_fp8:
                setattr(self, name, PerTensorQuantizedFp8Linear(input_size, size))
            else:
                setattr(self, name, nn.Linear(input_size, size, bias=False, dtype=model_config.params_dtype))

    def forward_ln(self, hidden_states):
        return self.layer_norm(hidden_states)

    def _versions(self, names):
        ps = [p for n in names for p in getattr(self, n).parameters()]
        return tuple((p.data_ptr(), p._version) for p in ps)

    def stacked_weight(self, with_qx: bool) -> torch.Tensor:
        """[Wq; Wk; Wv (; Wqx)] for the fused projection GEMM, built once and rebuilt only if a weight changes."""
        names = ["q", "k", "v"] + (["qx"] if with_qx else [])
        key = self._versions(names)
        cached = getattr(self, "_stacked", None)
        if cached is None or cached[0] != key:
            ws = [getattr(self, n).weight.detach() for n in names]
            if self.is_fp8:
                ws = [w[0] for w in ws]                       # (1, out, in) float8_e4m3fn, dit_module.py:441
            cached = self._stacked = (key, F.stack_qkv_weights(*ws))
        return cached[1]

    def fp8_fusable(self) -> bool:
        """The fused fp8 GEMM quantises x ONCE, so the four linears must divide by the same `input_scale` vector -
        which they do when calibrated on their common input, the LayerNorm output. Checked once per set of weights
        (one host sync); otherwise every linear quantises its own copy and runs as a library GEMM."""
        key = self._versions(["q", "k", "v", "qx"])
        cached = getattr(self, "_fusable", None)
        if cached is None or cached[0] != key:
            ref = self.q.input_scale
            same = all(torch.equal(getattr(self, n).input_scale, ref) for n in ("k", "v", "qx"))
            scale = torch.stack([(getattr(self, n).input_scale[0] * getattr(self, n).weight_scale[0]).detach()
                                 for n in ("q", "k", "v", "qx")]).float().contiguous()
            cached = self._fusable = (key, bool(same), scale)
        return cached[1]

    def fused_operands(self, x2d: torch.Tensor, with_qx: bool):
        """(x, stacked w, out_scale) for `F.qkv_proj_norm_rope`, or None when this layer cannot use the fused GEMM.
        bf16 layers: x itself and the stacked bf16 weights. fp8 layers (PerTensorQuantizedFp8Linear.forward,
        dit_module.py:446-457): x quantised once by `div_clamp_to`, the stacked e4m3 weights, and per projection the
        product of the two scalars `bmm_fp8` applies."""
        if x2d.dtype != torch.bfloat16:
            return None
        if not self.is_fp8:
            return (x2d, self.stacked_weight(with_qx), None) if x2d.shape[-1] % 64 == 0 else None
        if x2d.shape[-1] % 128 != 0 or not self.fp8_fusable():
            return None
        return self.q.quantize_input(x2d), self.stacked_weight(with_qx), self._fusable[2]

    def forward_q(self, x):
        return self.q(x)

    def forward_qx(self, x):
        return self.qx(x)

    def forward_k(self, x):
        return self.k(x)

    def forward_v(self, x):
        return self.v(x)


class KVOffload:
    """kv_offload (config.py:106, dit_module.py:803-811, 914-929): the per-layer KV caches live in pinned host memory.
    The reference copies the cache prefix to the GPU synchronously inside every layer's forward
    (`inference_key_and_value_memory[:sequence_start].cuda()`) and slice-assigns the new rows back.
    Here one manager per generation (kept on the InferenceParams) moves both directions on a copy stream:
      * prefix(layer): the device copy of cache[:start]. While layer L computes, the prefix of layer L + 1 is already
        on its way into the other of two staging buffers (the layers run in order, dit_module.py:1425-1440), so
        after the first layer the H2D transfer is off the critical path;
      * store(layer): the new rows go back device -> host asynchronously after the kernels that produced them;
        any later read of that layer's cache waits for the write-back event."""

    def __init__(self, device):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.stage = [None, None]          # ping-pong device staging buffers (rows, Hkv, 256)
        self.prefetched = {}               # layer -> (buffer index, rows, event)
        self.written = {}                  # layer -> event of the last device -> host write-back
        self.last_store = None             # event of the most recent write-back of any layer
        self.h2d_bytes = self.d2h_bytes = self.prefetch_hits = 0

    @classmethod
    def of(cls, inference_params: InferenceParams, device) -> "KVOffload":
        st = getattr(inference_params, "kv_offload_state", None)
        if st is None:
            st = inference_params.kv_offload_state = cls(device)
        return st

    def _buffer(self, i: int, like: torch.Tensor, rows: int) -> torch.Tensor:
        b = self.stage[i]
        if b is None or b.shape[0] < rows or b.shape[1:] != like.shape[1:]:
            b = self.stage[i] = torch.empty((max(rows, like.shape[0]),) + tuple(like.shape[1:]), dtype=like.dtype,
                                            device=self.device)
        return b

    def _h2d(self, layer: int, cache: torch.Tensor, rows: int, buf_index: int, stream) -> torch.Tensor:
        buf = self._buffer(buf_index, cache, rows)
        with torch.cuda.stream(stream):
            if layer in self.written:
                stream.wait_event(self.written[layer])
            buf[:rows].copy_(cache[:rows], non_blocking=True)
        self.h2d_bytes += rows * cache[0].numel() * cache.element_size()
        return buf

    def prefix(self, layer: int, cache: torch.Tensor, rows: int, next_layer_cache: Optional[torch.Tensor]):
        """Device view of cache[:rows] for `layer`, and kick the prefetch of the next layer's prefix."""
        cur = torch.cuda.current_stream(self.device)
        hit = self.prefetched.pop(layer, None)
        if hit is not None and hit[1] == rows:
            cur.wait_event(hit[2])
            idx = hit[0]
            buf = self.stage[idx]
            self.prefetch_hits += 1
        else:
            idx = layer & 1
            # any copy-stream transfer still in flight into a staging buffer (an orphaned prefetch: its row count no
            # longer matches, or it belongs to another layer after an aborted forward) must land before the compute
            # stream writes the same buffer
            if hit is not None:
                cur.wait_event(hit[2])
            for stale in [l for l, e in self.prefetched.items() if e[0] == idx]:
                cur.wait_event(self.prefetched.pop(stale)[2])
            buf = self._h2d(layer, cache, rows, idx, cur)
        if next_layer_cache is not None and rows > 0:
            # the other staging buffer was last read by the previous layer's attention, which the compute stream has
            # already been given: order the copy stream behind it
            self.copy_stream.wait_stream(cur)
            self._h2d(layer + 1, next_layer_cache, rows, idx ^ 1, self.copy_stream)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self.prefetched[layer + 1] = (idx ^ 1, rows, ev)
        return buf[:rows]

    def store(self, layer: int, cache: torch.Tensor, kv_new: torch.Tensor, row_offset: int, rows: int) -> None:
        """cache[row_offset : row_offset + rows] = kv_new[:rows], device -> pinned host, asynchronously."""
        if rows <= 0:
            return
        cur = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(cur)
        kv_new.record_stream(self.copy_stream)
        with torch.cuda.stream(self.copy_stream):
            cache[row_offset:row_offset + rows].copy_(kv_new[:rows], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        self.written[layer] = ev
        self.last_store = ev
        self.prefetched.pop(layer, None)   # a staged copy of this layer is stale now
        self.d2h_bytes += rows * cache[0].numel() * cache.element_size()

    def synchronize(self) -> None:
        self.copy_stream.synchronize()


def _shared_plan(holder, tag: str, heads_q: int, total_q: int, total_k: int, q_ranges, k_ranges):
    """The attention work list (F.attn_plan) of one range table, built once per forward: `holder` is the per-forward
    parameter object (PackedCoreAttnParams / PackedCrossAttnParams, the same instance for every layer of a denoise
    forward, dit_model.py:301-337), on which the plans are kept. Layers 1.. reuse what layer 0 built."""
    plans = holder.__dict__.get("_magi_plans")
    if plans is None:
        plans = {}
        object.__setattr__(holder, "_magi_plans", plans)   # frozen dataclass: plain attribute, not a field
    key = (tag, heads_q, total_q, total_k, q_ranges.data_ptr(), k_ranges.data_ptr())
    plan = plans.get(key)
    if plan is None:
        plan = plans[key] = F.attn_plan(q_ranges, k_ranges, num_heads_q=heads_q, total_q=total_q, total_k=total_k)
    return plan


class FullyParallelAttention(nn.Module):
    def __init__(self, model_config: ModelConfig, engine_config: EngineConfig, layer_number: int):
        super().__init__()
        self.model_config = model_config
        self.engine_config = engine_config
        self.layer_number = layer_number
        hd = model_config.kv_channels
        if hd != F.HEAD_DIM:
            raise ValueError(f"kv_channels must be {F.HEAD_DIM}")
        self.hidden_size_per_attention_head = hd
        self.query_projection_size = hd * model_config.num_attention_heads
        self.kv_projection_size = hd * model_config.num_query_groups
        # dit_module.py:794-798 takes the tp-cp world size from parallel_state; with cp_strategy "none" the reference
        # asserts cp_size == 1 (dit_module.py:1167), so the engine config is authoritative here
        world_size = 1 if engine_config.cp_strategy == "none" else engine_config.cp_size
        if engine_config.cp_strategy != "none" and mpu.get_cp_group() is not None:
            assert world_size == mpu.get_cp_world_size(), "engine_config.cp_size != context-parallel group size"
        groups = model_config.num_query_groups
        if engine_config.cp_strategy == "cp_ring_kv":
            self.num_query_groups_per_partition = groups   # ring-KV keeps all heads (and a replicated cache)
        elif world_size > groups and world_size % groups == 0:
            self.num_query_groups_per_partition = 1
        else:
            self.num_query_groups_per_partition = divide(groups, world_size)

        self.linear_qkv = CustomLayerNormLinear(model_config.hidden_size, self.query_projection_size,
                                                self.kv_projection_size, layer_number, model_config, engine_config)
        self.linear_kv_xattn = nn.Linear(int(model_config.hidden_size * model_config.xattn_cond_hidden_ratio),
                                         2 * self.kv_projection_size, dtype=model_config.params_dtype, bias=False)
        # dit_module.py:873-880: the out-projection is a PerChannelQuantizedFp8Linear on the layers fp8_quant covers
        self.adapt_linear_quant = uses_fp8(engine_config, model_config, layer_number)
        proj_cls = PerChannelQuantizedFp8Linear if self.adapt_linear_quant else nn.Linear
        self.linear_proj = proj_cls(2 * self.query_projection_size, model_config.hidden_size,
                                    dtype=model_config.params_dtype, bias=False)
        self.q_layernorm = HeadLayerNorm(model_config, hd)
        self.q_layernorm_xattn = HeadLayerNorm(model_config, hd)
        self.k_layernorm = HeadLayerNorm(model_config, hd)
        self.k_layernorm_xattn = HeadLayerNorm(model_config, hd)

    # ------------------------------------------------------------------ KV cache (dit_module.py:800-819, 891-946)
    def _allocate_key_and_value_memory(self, sequence_length, batch_size, dtype, device):
        shape = (sequence_length * batch_size, self.num_query_groups_per_partition,
                 2 * self.hidden_size_per_attention_head)
        if self.engine_config.kv_offload:
            # dit_module.py:803-811: the cache lives in pinned host memory; see KVOffload for how it is staged
            return torch.empty(shape, dtype=dtype, device="cpu", pin_memory=True)
        if self._fused_ulysses(batch_size, device) or self._fused_ring(batch_size, device):
            # fused Ulysses / fused ring-KV: the cache lives in symmetric (NVLink peer-mapped) memory so that the other
            # ranks' prologue kernels store their K/V rows straight into the cache slot. Collective: every rank of
            # the group allocates its layer cache in the same forward.
            import torch.distributed._symmetric_memory as symm
            cache = symm.empty(shape, dtype=dtype, device=device)
            handle = symm.rendezvous(cache, mpu.get_cp_group().group_name)
            self._cache_peer_ptrs = [int(p_) for p_ in handle.buffer_ptrs]
            self._cache_mc_ptr = int(getattr(handle, "multicast_ptr", 0) or 0)
            self._cache_peer_of = cache.data_ptr()
            return cache
        return torch.empty(shape, dtype=dtype, device=device)

    def _fused_ulysses(self, batch_size, device) -> bool:
        return (self.engine_config.cp_strategy == "cp_ulysses" and batch_size == 1 and not self.engine_config.kv_offload
                and getattr(self.engine_config, "cp_fused_all_to_all", False)
                and cpm.UlyssesPeerWorkspace.available(mpu.get_cp_group(), device))

    def _fused_ulysses_forward(self, batch_size, device, inference_params, meta_args) -> bool:
        """Does this forward take the fused all-to-all path: batch 1 always; batch > 1 when it is the cache-less pass
        with one range shared by all batch rows (the uncond pass of the 3-CFG models, dit_model.py:413-494 - the only
        batched form core_attention accepts, dit_module.py:1048)."""
        if batch_size == 1:
            return self._fused_ulysses(1, device)
        return (self._fused_ulysses(1, device) and not self._uses_cache(inference_params, meta_args)
                and meta_args.denoising_range_num == 1)

    def _fused_ring(self, batch_size, device) -> bool:
        return (self.engine_config.cp_strategy == "cp_ring_kv" and batch_size == 1
                and getattr(self.engine_config, "cp_fused_all_to_all", False)
                and cpm.RingGatherWorkspace.available(mpu.get_cp_group(), device))

    def _cache(self, inference_params: InferenceParams, dtype, device):
        d = inference_params.key_value_memory_dict
        if self.layer_number not in d:
            d[self.layer_number] = self._allocate_key_and_value_memory(
                inference_params.max_sequence_length, inference_params.max_batch_size, dtype, device)
        return d[self.layer_number]

    @staticmethod
    def _uses_cache(inference_params, meta_args: ModelMetaArgs) -> bool:
        """dit_module.py:936-943."""
        return inference_params is not None and (
            meta_args.extract_prefix_video_feature or meta_args.fwd_extra_1st_chunk or meta_args.slice_point > 0)

    def _cache_plan(self, inference_params, meta_args, new_rows: int, dtype, device):
        """(cache, sequence_start, rows_to_append) following dit_module.py:914-929."""
        if not self._uses_cache(inference_params, meta_args):
            return None, 0, 0
        cache = self._cache(inference_params, dtype, device)
        bs = inference_params.max_batch_size
        start = meta_args.slice_point * meta_args.clip_token_nums * bs
        rows = 0
        if inference_params.update_kv_cache:
            rows = new_rows - meta_args.clip_token_nums * bs if meta_args.distill_nearly_clean_chunk else new_rows
            assert start + rows <= cache.size(0)   # dit_module.py:927
        return cache, start, rows

    def _offloaded(self, cache) -> bool:
        """kv_offload (dit_module.py:803-811): this layer's cache lives in pinned host memory."""
        return cache is not None and self.engine_config.kv_offload and not cache.is_cuda

    def _offload_views(self, inference_params, cache, start):
        """kv_offload: device views (k, v) of the cache prefix cache[:start], staged through KVOffload."""
        off = KVOffload.of(inference_params, torch.device("cuda", torch.cuda.current_device()))
        nxt = inference_params.key_value_memory_dict.get(self.layer_number + 1)
        pre = off.prefix(self.layer_number, cache, start, nxt)
        return pre[..., :F.HEAD_DIM], pre[..., F.HEAD_DIM:]

    # ------------------------------------------------------------------ pieces
    def _project(self, mixed):
        sq, b, _ = mixed.shape
        x = mixed.reshape(sq * b, -1)
        return (self.linear_qkv.forward_q(x), self.linear_qkv.forward_k(x), self.linear_qkv.forward_v(x))

    def _ln(self, mod: HeadLayerNorm):
        return (mod.weight, mod.bias)

    PROJ_GROUPS = 8   # "hn=8 is hardcoded to align with TP8 training and TP1 inference", dit_module.py:1337

    # ------------------------------------------------------------------ reference-named pieces (dit_module.py:953-1021)
    # The reference hands these to UlyssesScheduler / CSOHelper as callables; they are kept under the same names and
    # argument meaning. Each one is a single launch of the fused prologue kernel.
    def get_q(self, mixed_qqkv: torch.Tensor, cos_emb: torch.Tensor, sin_emb: torch.Tensor, head_perm_cp: int = 1):
        """dit_module.py:953-963: [sq, b, hidden] -> [(sq b), hn, hd], LayerNorm(+1p) + RoPE in fp32, bf16 out."""
        sq, b, _ = mixed_qqkv.shape
        q, _ = F.qk_norm_rope(self.linear_qkv.forward_q(mixed_qqkv.reshape(sq * b, -1)), None, None,
                              self.model_config.num_attention_heads, self.model_config.num_query_groups,
                              self._ln(self.q_layernorm), None, sin_emb, cos_emb, batch=b,
                              eps=self.model_config.layernorm_epsilon,
                              zero_centered_gamma=self.model_config.apply_layernorm_1p, head_perm_cp=head_perm_cp)
        return q

    def get_kv(self, mixed_qqkv: torch.Tensor, cos_emb: torch.Tensor, sin_emb: torch.Tensor, head_perm_cp: int = 1):
        """dit_module.py:991-996 (get_k + get_v + cat): packed [(sq b), hn, 2 * hd] = [K | V]."""
        sq, b, _ = mixed_qqkv.shape
        x = mixed_qqkv.reshape(sq * b, -1)
        _, kv = F.qk_norm_rope(None, self.linear_qkv.forward_k(x), self.linear_qkv.forward_v(x),
                               self.model_config.num_attention_heads, self.model_config.num_query_groups, None,
                               self._ln(self.k_layernorm), sin_emb, cos_emb, batch=b,
                               eps=self.model_config.layernorm_epsilon,
                               zero_centered_gamma=self.model_config.apply_layernorm_1p, head_perm_cp=head_perm_cp)
        return kv

    def get_k(self, mixed_qqkv, cos_emb, sin_emb):
        """dit_module.py:970-980 (a view of the packed result)."""
        return self.get_kv(mixed_qqkv, cos_emb, sin_emb)[..., :F.HEAD_DIM]

    def get_v(self, mixed_qqkv):
        """dit_module.py:987-989: pure re-layout of the projection output."""
        sq, b, _ = mixed_qqkv.shape
        return self.linear_qkv.forward_v(mixed_qqkv.reshape(sq * b, -1)).view(sq * b, -1, F.HEAD_DIM)

    def get_qkv(self, mixed_qqkv, cos_emb, sin_emb):
        """dit_module.py:998-1003: one prologue launch for all three."""
        sq, b, _ = mixed_qqkv.shape
        x = mixed_qqkv.reshape(sq * b, -1)
        q, kv = F.qk_norm_rope(self.linear_qkv.forward_q(x), self.linear_qkv.forward_k(x), self.linear_qkv.forward_v(x),
                               self.model_config.num_attention_heads, self.model_config.num_query_groups,
                               self._ln(self.q_layernorm), self._ln(self.k_layernorm), sin_emb, cos_emb, batch=b,
                               eps=self.model_config.layernorm_epsilon,
                               zero_centered_gamma=self.model_config.apply_layernorm_1p)
        return q, kv[..., :F.HEAD_DIM], kv[..., F.HEAD_DIM:]

    def get_xqkv(self, mixed_qqkv, key_value_states):
        """dit_module.py:1005-1021: (query_xattn [(b sq), hn, hd], key_xattn, value_xattn [ytot, hn, hd])."""
        q_rows, kv_x = self.cross_attention_prepare(mixed_qqkv, key_value_states)
        q = q_rows[0] if len(q_rows) == 1 else torch.cat(q_rows, dim=0)
        return q, kv_x[..., :F.HEAD_DIM], kv_x[..., F.HEAD_DIM:]

    def adjust_key_and_value_for_inference(self, key_and_value, inference_params, meta_args):
        """dit_module.py:933-946, for callers that want the reference's materialised (key, value): performs the cache
        append and returns views of `cat(cache prefix, new)`. The fast paths of this class never call it - they hand
        the cache prefix and the new tokens to the kernel as two segments and skip the concatenation."""
        if inference_params is None:
            return key_and_value[..., :F.HEAD_DIM], key_and_value[..., F.HEAD_DIM:]
        cache, start, app = self._cache_plan(inference_params, meta_args, key_and_value.shape[0], key_and_value.dtype,
                                             key_and_value.device)
        if self._offloaded(cache):
            off = KVOffload.of(inference_params, key_and_value.device)
            pre = off.prefix(self.layer_number, cache, start,
                             inference_params.key_value_memory_dict.get(self.layer_number + 1)) if start > 0 else None
            off.store(self.layer_number, cache, key_and_value, start, app)
            if pre is not None:
                key_and_value = torch.cat([pre, key_and_value], dim=0)
        elif cache is not None:
            if app > 0:
                F.kv_cache_append(cache, key_and_value, start, app)
            if start > 0:
                key_and_value = torch.cat([cache[:start], key_and_value], dim=0)
        return key_and_value[..., :F.HEAD_DIM], key_and_value[..., F.HEAD_DIM:]

    def full_attention(self, bs: int, meta_args: ModelMetaArgs, q, k, v, i: int, out=None, k_cache=None, v_cache=None):
        """dit_module.py:1068-1096 (the cp_shuffle_overlap entry): q is already sliced to denoising range i, k / v are
        the full key / value sequence (or, with k_cache / v_cache, the new tokens that follow the cache prefix); rows
        attend to k_range[i]."""
        assert bs == 1   # dit_module.py:1070
        p = meta_args.core_attn_params
        if out is None:
            out = torch.empty((q.shape[0], q.shape[1], F.HEAD_DIM), dtype=q.dtype, device=q.device)
        q_range = p.q_range[i:i + 1] - p.q_range[i, :1]         # rebased to the slice (dit_module.py:1072)
        F.attn_ranges(q, k, v, q_range, p.k_range[i:i + 1], k_cache=k_cache, v_cache=v_cache, out=out,
                      return_lse=False)
        return out

    def core_attention(self, query, k_new, v_new, bs: int, meta_args: ModelMetaArgs, k_cache=None, v_cache=None,
                       out=None, proj_buf=None):
        """dit_module.py:1023-1066 on "(sq b)"-ordered rows. bs > 1 only occurs with one range shared by all batch
        rows (dit_module.py:1048-1061): each batch row is a strided view, no transposes or copies.
        proj_buf: (sq*b, 2*Hq*128) out-projection input to write into (slot n = 0) instead of `out`."""
        p = meta_args.core_attn_params
        if out is None and proj_buf is None:
            out = torch.empty_like(query)

        def row(x, i):
            return x.unflatten(0, (-1, bs))[:, i]

        def dest(i=None):
            if proj_buf is not None:
                return dict(proj_input=(proj_buf if i is None else row(proj_buf, i), 0, self.PROJ_GROUPS))
            return dict(out=out if i is None else row(out, i))

        if bs == 1:
            total_k = k_new.shape[0] + (0 if k_cache is None else k_cache.shape[0])
            plan = _shared_plan(p, "core", query.shape[1], query.shape[0], total_k, p.q_range, p.k_range)
            F.attn_ranges(query, k_new, v_new, p.q_range, p.k_range, k_cache=k_cache, v_cache=v_cache,
                          return_lse=False, plan=plan, **dest())
            return out
        assert not (bs > 1 and meta_args.denoising_range_num > 1)
        assert k_cache is None
        for i in range(bs):
            F.attn_ranges(row(query, i), row(k_new, i), row(v_new, i), p.q_range[:1], p.k_range[:1],
                          return_lse=False, **dest(i))
        return out

    def cross_prologue_inputs(self, mixed_qqkv, key_value_states, with_qx=True):
        """The `cross=` argument of F.qk_norm_rope: the cross-attention's projections (qx, packed text [K|V]) and their
        bf16 LayerNorm parameters, so that they ride in the SAME prologue launch as q / k / v (one prologue launch per
        layer instead of three). None when that is not possible (batch > 1: the "(sq b) -> (b sq)" row order of the
        query needs one strided view per batch row; `_xattn` norms promoted to fp32)."""
        sq, b, _ = mixed_qqkv.shape
        lq, lk = self._ln(self.q_layernorm_xattn), self._ln(self.k_layernorm_xattn)
        if b != 1 or any(t.dtype != torch.bfloat16 for t in lq + lk) or mixed_qqkv.dtype != torch.bfloat16:
            return None
        Hkv = self.model_config.num_query_groups
        kvx = torch.nn.functional.linear(key_value_states, self.linear_kv_xattn.weight)
        d = dict(text_kv=kvx.view(key_value_states.shape[0], Hkv, 2 * F.HEAD_DIM), ln_tk=lk)
        if with_qx:
            d.update(x_qx=self.linear_qkv.forward_qx(mixed_qqkv.reshape(sq * b, -1)), ln_qx=lq)
        return d

    def cross_attention_prepare(self, mixed_qqkv, key_value_states, q_rows=None, cross_attn_params=None,
                                proj_buf=None, kv_x=None):
        """get_xqkv, dit_module.py:1005-1021: qx / text-kv projections + the two bf16 LayerNorms (no RoPE). Split from
        the attention itself so that, under context parallelism, this half can overlap the input all-to-alls and the
        attention half the output all-to-all.
        With cross_attn_params (batch 1) the result also carries everything else the attention launch needs - its
        destination and its work list - so that `cross_attention_attend(..., independent=True)` right after the
        self-attention launch touches nothing that was allocated or written after that launch was issued."""
        prepared = self._cross_attention_prepare(mixed_qqkv, key_value_states, q_rows, kv_x)
        sq, b, _ = mixed_qqkv.shape
        if cross_attn_params is None or b != 1:
            return prepared
        q_rows_, kv_x = prepared
        qr, kr = cross_attn_params.q_ranges, cross_attn_params.kv_ranges
        if not (qr.dtype == torch.int32 and qr.is_contiguous() and kr.is_contiguous()):
            return prepared
        plan = _shared_plan(cross_attn_params, "cross", q_rows_[0].shape[1], q_rows_[0].shape[0], kv_x.shape[0], qr, kr)
        out = None if proj_buf is not None else torch.empty(
            (sq, b, self.model_config.num_attention_heads, F.HEAD_DIM), dtype=kv_x.dtype, device=kv_x.device)
        return q_rows_, kv_x, out, plan

    def _cross_attention_prepare(self, mixed_qqkv, key_value_states, q_rows=None, kv_x=None):
        sq, b, _ = mixed_qqkv.shape
        Hq, Hkv = self.model_config.num_attention_heads, self.model_config.num_query_groups
        eps, z = self.model_config.layernorm_epsilon, self.model_config.apply_layernorm_1p
        if kv_x is not None and q_rows is not None:   # both came out of the layer's one prologue launch
            return q_rows, kv_x
        kvx = torch.nn.functional.linear(key_value_states, self.linear_kv_xattn.weight)  # (ytot, Hkv*256)
        kvx = kvx.view(key_value_states.shape[0], Hkv, 2 * F.HEAD_DIM)
        # text K: LayerNorm (bf16 params, no RoPE); V: copy -> packed (ytot, Hkv, 256)
        _, kv_x = F.qk_norm_rope(None, kvx[..., :F.HEAD_DIM], kvx[..., F.HEAD_DIM:], Hq, Hkv, None,
                                 self._ln(self.k_layernorm_xattn), None, None, eps=eps, zero_centered_gamma=z)
        if q_rows is not None:   # the normalised cross-attention query came out of the fused projection GEMM
            return q_rows, kv_x
        qx = self.linear_qkv.forward_qx(mixed_qqkv.reshape(sq * b, -1))                 # rows in (sq b) order
        q_rows = []
        for i in range(b):
            # "(sq b) -> (b sq)": batch row i is a strided view of the projection output
            q_i, _ = F.qk_norm_rope(qx.view(sq, b, Hq, F.HEAD_DIM)[:, i], None, None, Hq, Hkv,
                                    self._ln(self.q_layernorm_xattn), None, None, None, eps=eps,
                                    zero_centered_gamma=z)
            q_rows.append(q_i)
        return q_rows, kv_x

    def cross_attention_attend(self, prepared, shape, cross_attn_params: PackedCrossAttnParams, proj_buf=None,
                               independent=False):
        """cross_attention, dit_module.py:1098-1136, on the tensors of cross_attention_prepare.
        independent: the launch directly follows the self-attention launch of the same layer; with a destination and
        plan that `cross_attention_prepare` made before that launch it shares nothing with it and may overlap its
        tail (functional.attn_ranges independent_of_prev)."""
        q_rows, kv_x = prepared[:2]
        pre_out, pre_plan = (prepared[2], prepared[3]) if len(prepared) == 4 else (None, None)
        sq, b = shape
        Hq = self.model_config.num_attention_heads
        out = None if proj_buf is not None else pre_out if pre_out is not None else torch.empty(
            (sq, b, Hq, F.HEAD_DIM), dtype=kv_x.dtype, device=kv_x.device)
        independent = independent and pre_plan is not None and b == 1
        # q_ranges / kv_ranges (the flex_flash_attn_func arguments, dit_module.py:1111-1122) rather than cu_seqlens:
        # they stay valid when a caption segment is empty and after the per-rank rewrite under context parallelism
        n = cross_attn_params.q_ranges.shape[0]
        per_b = n // b
        for i in range(b):
            qr = cross_attn_params.q_ranges[i * per_b:(i + 1) * per_b]
            if i > 0:
                qr = qr - i * sq
            kr = cross_attn_params.kv_ranges[i * per_b:(i + 1) * per_b]
            if proj_buf is not None:
                dst = dict(proj_input=(proj_buf.view(sq, b, -1)[:, i], 1, self.PROJ_GROUPS))
            else:
                dst = dict(out=out[:, i])
            plan = pre_plan
            if plan is None and b == 1 and qr.dtype == torch.int32 and qr.is_contiguous() and kr.is_contiguous():
                plan = _shared_plan(cross_attn_params, "cross", q_rows[i].shape[1], q_rows[i].shape[0], kv_x.shape[0],
                                    qr, kr)
            F.attn_ranges(q_rows[i], kv_x[..., :F.HEAD_DIM], kv_x[..., F.HEAD_DIM:], qr, kr, return_lse=False,
                          plan=plan, independent_of_prev=independent, **dst)
        return None if proj_buf is not None else out.view(sq, b, Hq * F.HEAD_DIM)

    def cross_attention(self, mixed_qqkv, key_value_states, cross_attn_params: PackedCrossAttnParams, proj_buf=None,
                        q_rows=None):
        """get_xqkv + cross_attention, dit_module.py:1005-1021, 1098-1136. Returns (sq, b, Hq*128), or writes slot
        n = 1 of the out-projection input `proj_buf` (sq*b, 2*Hq*128) and returns None."""
        prepared = self.cross_attention_prepare(mixed_qqkv, key_value_states, q_rows)
        return self.cross_attention_attend(prepared, mixed_qqkv.shape[:2], cross_attn_params, proj_buf)

    @staticmethod
    def attn_linear_proj_input(core_attn_out: torch.Tensor, cross_attn_out: torch.Tensor) -> torch.Tensor:
        """What the reference builds before linear_proj (attn_linear_proj, dit_module.py:1332-1338): the torch
        cat + rearrange that `forward(..., fused_proj_input=True)` makes unnecessary. Kept for callers that use the
        two-output interface and as the checker of the fused layout."""
        sq, b, w = core_attn_out.shape
        x = torch.stack([core_attn_out, cross_attn_out], dim=2)                       # sq b n (hn hd)
        x = x.view(sq, b, 2, FullyParallelAttention.PROJ_GROUPS, w // FullyParallelAttention.PROJ_GROUPS)
        return x.transpose(2, 3).reshape(sq, b, 2 * w)                                # sq b (hn n hd)

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, key_value_states: torch.Tensor,
                inference_params: Optional[InferenceParams], rotary_pos_emb: torch.Tensor, meta_args: ModelMetaArgs,
                fused_proj_input: bool = False):
        """Returns (core_attn_out, xattn_out) like the reference. With fused_proj_input=True (cp_strategy "none") both
        attention kernels write straight into the out-projection input "sq b (hn n hd)" and that single tensor is
        returned instead (no torch.cat, no rearrange copy)."""
        assert rotary_pos_emb is not None, "FullyParallelAttention needs rotary_pos_emb"
        sin_emb, cos_emb = rotary_pos_emb.tensor_split(2, -1)   # dit_module.py:1147: table is [sin | cos]
        sq, batch_size, _ = hidden_states.shape
        Hq, Hkv = self.model_config.num_attention_heads, self.model_config.num_query_groups
        eps, z = self.model_config.layernorm_epsilon, self.model_config.apply_layernorm_1p
        mixed_qqkv = self.linear_qkv.forward_ln(hidden_states)
        strategy = self.engine_config.cp_strategy
        rows = sq * batch_size
        dev, dt = hidden_states.device, hidden_states.dtype

        if strategy == "none":
            assert self.engine_config.cp_size == 1
            cache, start, app = self._cache_plan(inference_params, meta_args, rows, dt, dev)
            offload = self._offloaded(cache)
            xq_rows = None
            query = None
            if getattr(self.engine_config, "fused_qkv_proj", False):
                # projections + prologue as one GEMM (its epilogue is the prologue); with batch 1 the cross-attention
                # query rides along as a fourth group of heads
                with_qx = batch_size == 1 and (2 * Hq + 2 * Hkv) % 2 == 0
                ops = (self.linear_qkv.fused_operands(mixed_qqkv.reshape(rows, -1), with_qx)
                       if with_qx or (Hq + 2 * Hkv) % 2 == 0 else None)
                if ops is not None:
                    query, kv_new, qx = F.qkv_proj_norm_rope(
                        ops[0], ops[1], Hq, Hkv,
                        self._ln(self.q_layernorm), self._ln(self.k_layernorm), sin_emb, cos_emb,
                        ln_qx=self._ln(self.q_layernorm_xattn) if with_qx else None, batch=batch_size, eps=eps,
                        zero_centered_gamma=z, cache=cache if app > 0 and not offload else None,
                        cache_row_offset=start, cache_rows=app, out_scale=ops[2])
                    xq_rows = [qx] if with_qx else None
            x_kv = None
            if query is None:
                xq, xk, xv = self._project(mixed_qqkv)
                cross = self.cross_prologue_inputs(mixed_qqkv, key_value_states)
                query, kv_new = F.qk_norm_rope(xq, xk, xv, Hq, Hkv, self._ln(self.q_layernorm),
                                               self._ln(self.k_layernorm), sin_emb, cos_emb, batch=batch_size, eps=eps,
                                               zero_centered_gamma=z,
                                               cache=cache if app > 0 and not offload else None,
                                               cache_row_offset=start, cache_rows=app, cross=cross)
                if cross is not None:
                    xq_rows, x_kv = [cross["qx"]], cross["tkv"]
            kc = vc = None
            if cache is not None and start > 0:
                kc, vc = (self._offload_views(inference_params, cache, start) if offload else
                          (cache[:start, :, :F.HEAD_DIM], cache[:start, :, F.HEAD_DIM:]))
            if offload:
                KVOffload.of(inference_params, dev).store(self.layer_number, cache, kv_new, start, app)
            # the cross-attention's own projections / norms, destination and plan come first: the launch of its
            # attention then follows the self-attention launch with nothing shared between the two
            xp = meta_args.cross_attn_params
            # ... provided a self-attention kernel IS launched in between (batch 1: one launch over all ranges)
            overlap = batch_size == 1 and rows > 0 and meta_args.core_attn_params.q_range.shape[0] > 0
            if fused_proj_input:
                buf = torch.empty((rows, 2 * Hq * F.HEAD_DIM), dtype=dt, device=dev)
                x_prepared = self.cross_attention_prepare(mixed_qqkv, key_value_states, xq_rows, xp, proj_buf=buf,
                                                          kv_x=x_kv)
                self.core_attention(query, kv_new[..., :F.HEAD_DIM], kv_new[..., F.HEAD_DIM:], batch_size, meta_args,
                                    kc, vc, proj_buf=buf)
                self.cross_attention_attend(x_prepared, (sq, batch_size), xp, proj_buf=buf, independent=overlap)
                return buf.view(sq, batch_size, 2 * Hq * F.HEAD_DIM)
            x_prepared = self.cross_attention_prepare(mixed_qqkv, key_value_states, xq_rows, xp, kv_x=x_kv)
            core_out = torch.empty_like(query)
            core = self.core_attention(query, kv_new[..., :F.HEAD_DIM], kv_new[..., F.HEAD_DIM:], batch_size,
                                       meta_args, kc, vc, out=core_out)
            core_attn_out = core.view(sq, batch_size, Hq * F.HEAD_DIM)
            xattn_out = self.cross_attention_attend(x_prepared, (sq, batch_size), xp, independent=overlap)

        elif fused_proj_input and not (strategy == "cp_ulysses" and batch_size == 1
                                       and self._fused_ulysses(batch_size, dev)):
            raise NotImplementedError("fused_proj_input is available with cp_strategy 'none' and on the fused "
                                      "cp_ulysses path (see supports_fused_proj_input)")
        elif strategy == "cp_ulysses":
            if self._fused_ulysses_forward(batch_size, dev, inference_params, meta_args):
                res = self._forward_ulysses_fused(mixed_qqkv, key_value_states, inference_params, sin_emb, cos_emb,
                                                  meta_args, fused_proj_input=fused_proj_input)
                if fused_proj_input:
                    return res
                core_attn_out, xattn_out = res
            else:
                core_attn_out, xattn_out = self._forward_ulysses(mixed_qqkv, key_value_states, inference_params,
                                                                 sin_emb, cos_emb, meta_args)
        elif strategy == "cp_ring_kv":
            fwd = self._forward_ring_fused if self._fused_ring(batch_size, dev) else self._forward_ring
            core_attn_out, xattn_out = fwd(mixed_qqkv, key_value_states, inference_params, sin_emb, cos_emb, meta_args)
        elif strategy == "cp_shuffle_overlap":
            core_attn_out, xattn_out = self._forward_cso(mixed_qqkv, key_value_states, inference_params, sin_emb,
                                                         cos_emb, meta_args)
        else:
            raise ValueError(f"Unsupported cp_strategy: {strategy}")   # dit_module.py:1244
        return core_attn_out, xattn_out

    # ------------------------------------------------------------------ Ulysses (dit_module.py:1175-1205)
    def _forward_ulysses(self, mixed_qqkv, key_value_states, inference_params, sin_emb, cos_emb, meta_args):
        sq, b, _ = mixed_qqkv.shape
        cp = self.engine_config.cp_size
        Hq, Hkv = self.model_config.num_attention_heads, self.model_config.num_query_groups
        eps, z = self.model_config.layernorm_epsilon, self.model_config.apply_layernorm_1p
        rows = sq * b
        dev, dt = mixed_qqkv.device, mixed_qqkv.dtype
        split = [x * b for x in meta_args.cp_split_sizes]   # dit_module.py:1150
        total_rows = sum(split)
        x2d = mixed_qqkv.reshape(rows, -1)
        cache, start, app = self._cache_plan(inference_params, meta_args, total_rows, dt, dev)
        hkv_loc = self.num_query_groups_per_partition
        offload = self._offloaded(cache)
        # the KV all-to-all lands straight in the cache slot when every received row is to be cached
        direct = cache is not None and app == total_rows and not offload
        kv_recv = cache[start:start + total_rows] if direct else None

        def get_kv():
            _, kv = F.qk_norm_rope(None, self.linear_qkv.forward_k(x2d), self.linear_qkv.forward_v(x2d), Hq, Hkv, None,
                                   self._ln(self.k_layernorm), sin_emb, cos_emb, batch=b, eps=eps,
                                   zero_centered_gamma=z, head_perm_cp=cp)
            return kv

        def get_q():
            q, _ = F.qk_norm_rope(self.linear_qkv.forward_q(x2d), None, None, Hq, Hkv, self._ln(self.q_layernorm), None,
                                  sin_emb, cos_emb, batch=b, eps=eps, zero_centered_gamma=z, head_perm_cp=cp)
            return q

        def kv_cache_func(kv):
            if offload:
                KVOffload.of(inference_params, dev).store(self.layer_number, cache, kv, start, app)
            elif cache is not None and app > 0 and not direct:
                F.kv_cache_append(cache, kv, start, app)
            if cache is not None and start > 0:
                self._kc, self._vc = (self._offload_views(inference_params, cache, start) if offload else
                                      (cache[:start, :, :F.HEAD_DIM], cache[:start, :, F.HEAD_DIM:]))
            else:
                self._kc = self._vc = None
            return kv[..., :F.HEAD_DIM], kv[..., F.HEAD_DIM:]

        def core(q, k, v, out, kv_heads):
            kc = vc = None
            if self._kc is not None:   # same kv-head slice of the cache as of the new tokens
                kc, vc = self._kc[:, kv_heads[0]:kv_heads[1]], self._vc[:, kv_heads[0]:kv_heads[1]]
            return self.core_attention(q, k, v, b, meta_args, kc, vc, out=out)

        def relayout_out(recv, cp_, result, q_head, lo):
            F.ulysses_relayout(recv, cp_, 1, out=result, dst_heads_total=q_head, dst_head_offset=lo)

        kv = get_kv()
        kv, h_kv = cpm.all_to_all_input_split(kv, split, send_layout=True, out=kv_recv)
        query = get_q()
        query, h_q = cpm.all_to_all_input_split(query, split, send_layout=True)
        # the cross-attention needs no communication (context_parallel.py:591): its projections + norms run under the
        # two input all-to-alls, its attention under the (last) output all-to-all
        x_prepared = self.cross_attention_prepare(mixed_qqkv, key_value_states)
        h_kv.wait()
        key, value = kv_cache_func(kv)
        h_q.wait()
        out = cpm.UlyssesScheduler.get_attn_and_xattn_base(
            query, key, value, core, partial(self.cross_attention_attend, x_prepared, (sq, b),
                                             meta_args.cross_attn_params),
            self.engine_config.ulysses_overlap_degree, b, cp, split, relayout_out=relayout_out)
        self._kc = self._vc = None
        return out

    # ------------------------------------------------------------------ Ulysses, all-to-alls fused into the kernels
    def supports_fused_proj_input(self, batch_size: int, device) -> bool:
        """Can forward(..., fused_proj_input=True) run here: single GPU, or Ulysses with the fused all-to-alls."""
        if self.model_config.num_attention_heads % self.PROJ_GROUPS != 0:
            return False
        s_ = self.engine_config.cp_strategy
        return s_ == "none" or (s_ == "cp_ulysses" and self._fused_ulysses(batch_size, device))

    def _forward_ulysses_fused(self, mixed_qqkv, key_value_states, inference_params, sin_emb, cos_emb, meta_args,
                               fused_proj_input=False):
        """dit_module.py:1175-1205 without NCCL on the data path: ONE prologue launch stores every head slice of q and
        [K|V] straight into the owning rank's receive buffer, the attention epilogue stores every output row straight
        into the buffer of the rank that owns the row (NVLink peer memory; include/magi_attn.h "peer" fields). Two
        device-side barriers per layer replace three all-to-all calls and the re-layout kernel, and the transfers
        overlap the kernels that issue them. Same values as _forward_ulysses (same kernels, same per-head inputs)."""
        sq, b, _ = mixed_qqkv.shape
        if b != 1:
            assert not fused_proj_input
            return self._forward_ulysses_fused_batched(mixed_qqkv, key_value_states, sin_emb, cos_emb, meta_args)
        cp, rank = self.engine_config.cp_size, mpu.get_cp_rank()
        Hq, Hkv = self.model_config.num_attention_heads, self.model_config.num_query_groups
        eps, z = self.model_config.layernorm_epsilon, self.model_config.apply_layernorm_1p
        dev, dt = mixed_qqkv.device, mixed_qqkv.dtype
        split = list(meta_args.cp_split_sizes)
        total_rows = sum(split)
        hq_loc, hkv_loc = divide(Hq, cp), self.num_query_groups_per_partition
        ws = cpm.UlyssesPeerWorkspace.get(mpu.get_cp_group(), total_rows, max(split), hq_loc, hkv_loc, Hq, dev)
        row_off, bounds, head_off = cpm.fused_ulysses_plan(split, rank, Hq, cp)
        cache, start, app = self._cache_plan(inference_params, meta_args, total_rows, dt, dev)
        # when every row of this forward is cached, the peers store K/V straight into the cache slot
        direct = (cache is not None and app == total_rows
                  and getattr(self, "_cache_peer_of", None) == cache.data_ptr())
        row_bytes = hkv_loc * 2 * F.HEAD_DIM * 2
        kv_ptrs = [p_ + start * row_bytes for p_ in self._cache_peer_ptrs] if direct else ws.kv_ptrs
        lnx = self._ln(self.q_layernorm_xattn) + self._ln(self.k_layernorm_xattn)
        ops = (self.linear_qkv.fused_operands(mixed_qqkv.reshape(sq, -1), True)
               if getattr(self.engine_config, "fused_qkv_proj", False)
               and all(t.dtype == torch.bfloat16 for t in lnx) else None)
        if ops is not None:
            # projections + prologue + input all-to-all as ONE launch: the GEMM's epilogue normalises / rotates every
            # head and stores it into the owning rank's receive buffer while the main loop computes the next tile
            _, _, qx = F.qkv_proj_norm_rope(
                ops[0], ops[1], Hq, Hkv, self._ln(self.q_layernorm),
                self._ln(self.k_layernorm), sin_emb, cos_emb, ln_qx=self._ln(self.q_layernorm_xattn), batch=1, eps=eps,
                zero_centered_gamma=z, peer_q_out=ws.q_ptrs, peer_kv_out=kv_ptrs, peer_row_offset=row_off,
                out_scale=ops[2])
            kvx = torch.nn.functional.linear(key_value_states, self.linear_kv_xattn.weight)
            cross = dict(text_kv=kvx.view(key_value_states.shape[0], Hkv, 2 * F.HEAD_DIM),
                         ln_tk=self._ln(self.k_layernorm_xattn))
            F.qk_norm_rope(None, None, None, Hq, Hkv, None, None, None, None, eps=eps, zero_centered_gamma=z, cross=cross)
            xq_rows, x_kv = [qx], cross["tkv"]
        else:
            xq, xk, xv = self._project(mixed_qqkv)
            # the cross-attention's projections + norms ride in the same launch (they need no communication)
            cross = self.cross_prologue_inputs(mixed_qqkv, key_value_states)
            F.qk_norm_rope(xq, xk, xv, Hq, Hkv, self._ln(self.q_layernorm), self._ln(self.k_layernorm), sin_emb,
                           cos_emb, batch=1, eps=eps, zero_centered_gamma=z, head_perm_cp=cp, peer_q_out=ws.q_ptrs,
                           peer_kv_out=kv_ptrs, peer_row_offset=row_off, peer_heads=(hq_loc, hkv_loc), cross=cross)
            xq_rows, x_kv = ([cross["qx"]], cross["tkv"]) if cross is not None else (None, None)
        # no communication in the cross-attention (context_parallel.py:591): its projections + norms are issued
        # before the barrier, its attention before the second one
        proj = proj_ptrs = None
        if fused_proj_input:
            # the out-projection input of this rank's rows, in symmetric memory: slot 0 of "(hn n hd)" is filled by
            # every rank's attention epilogue (its heads), slot 1 by the local cross-attention - no cat, no rearrange,
            # no copy out of the receive buffer
            proj, proj_ptrs = ws.next_proj()
        x_prepared = self.cross_attention_prepare(mixed_qqkv, key_value_states, xq_rows, meta_args.cross_attn_params,
                                                  proj_buf=proj[:sq] if fused_proj_input else None, kv_x=x_kv)
        ws.barrier()                                                     # A: every rank's q / kv rows have landed
        kv = cache[start:start + total_rows] if direct else ws.kv
        if cache is not None and app > 0 and not direct:
            F.kv_cache_append(cache, kv, start, app)
        kc = vc = None
        if cache is not None and start > 0:
            kc, vc = cache[:start, :, :F.HEAD_DIM], cache[:start, :, F.HEAD_DIM:]
        p = meta_args.core_attn_params
        plan = _shared_plan(p, "core-ulysses", ws.q.shape[1], ws.q.shape[0],
                            total_rows + (0 if kc is None else kc.shape[0]), p.q_range, p.k_range)
        if fused_proj_input:
            F.attn_ranges(ws.q, kv[..., :F.HEAD_DIM], kv[..., F.HEAD_DIM:], p.q_range, p.k_range, k_cache=kc,
                          v_cache=vc, return_lse=False, peer_out=(proj_ptrs, bounds, head_off, Hq),
                          peer_proj=(0, self.PROJ_GROUPS), plan=plan)
            self.cross_attention_attend(x_prepared, (sq, b), meta_args.cross_attn_params, proj_buf=proj[:sq],
                                        independent=True)
            ws.barrier()                                                 # B: every rank's output rows have landed
            return proj[:sq].view(sq, b, 2 * Hq * F.HEAD_DIM)
        F.attn_ranges(ws.q, kv[..., :F.HEAD_DIM], kv[..., F.HEAD_DIM:], p.q_range, p.k_range, k_cache=kc, v_cache=vc,
                      return_lse=False, peer_out=(ws.out_ptrs, bounds, head_off, Hq), plan=plan)
        xattn_out = self.cross_attention_attend(x_prepared, (sq, b), meta_args.cross_attn_params, independent=True)
        ws.barrier()                                                     # B: every rank's output rows have landed
        # the workspace is reused by the next layer: hand out a private copy of this rank's rows
        core_attn_out = ws.out[:split[rank]].reshape(sq, b, Hq * F.HEAD_DIM).clone()
        return core_attn_out, xattn_out

    def _forward_ulysses_fused_batched(self, mixed_qqkv, key_value_states, sin_emb, cos_emb, meta_args):
        """The fused path for batch > 1 without a KV cache (the uncond pass of the 3-CFG models): rows stay in the
        reference's "(sq b)" order everywhere - the receive buffers hold "(S b)" rows (this rank's block starts at row
        bounds[rank] * b), every batch row is a strided view of them (as in core_attention, dit_module.py:1048-1061), and
        the epilogue of batch row i stores into the owners' "(sq b) (cp hn) hd" output at token stride b * Hq * 128."""
        sq, b, _ = mixed_qqkv.shape
        cp, rank = self.engine_config.cp_size, mpu.get_cp_rank()
        Hq, Hkv = self.model_config.num_attention_heads, self.model_config.num_query_groups
        eps, z = self.model_config.layernorm_epsilon, self.model_config.apply_layernorm_1p
        dev = mixed_qqkv.device
        split = list(meta_args.cp_split_sizes)           # sequence rows per rank (dit_module.py:1150 multiplies by b)
        total = sum(split)
        hq_loc, hkv_loc = divide(Hq, cp), (divide(Hkv, cp) if Hkv >= cp else 1)   # (also serves cp_ring_kv: no cache)
        ws = cpm.UlyssesPeerWorkspace.get(mpu.get_cp_group(), total * b, max(split) * b, hq_loc, hkv_loc, Hq, dev)
        _, bounds, head_off = cpm.fused_ulysses_plan(split, rank, Hq, cp)
        ops = (self.linear_qkv.fused_operands(mixed_qqkv.reshape(sq * b, -1), False)
               if getattr(self.engine_config, "fused_qkv_proj", False) and (Hq + 2 * Hkv) % 2 == 0 else None)
        if ops is not None:
            # same choice as the single-GPU forward makes for batch > 1: q / k / v through the fused projection GEMM (no
            # qx: its "(sq b) -> (b sq)" row order needs one strided view per batch row), here with peer routing
            F.qkv_proj_norm_rope(ops[0], ops[1], Hq, Hkv,
                                 self._ln(self.q_layernorm), self._ln(self.k_layernorm), sin_emb, cos_emb, batch=b,
                                 eps=eps, zero_centered_gamma=z, peer_q_out=ws.q_ptrs, peer_kv_out=ws.kv_ptrs,
                                 peer_row_offset=bounds[rank] * b, out_scale=ops[2])
        else:
            xq, xk, xv = self._project(mixed_qqkv)
            F.qk_norm_rope(xq, xk, xv, Hq, Hkv, self._ln(self.q_layernorm), self._ln(self.k_layernorm), sin_emb,
                           cos_emb, batch=b, eps=eps, zero_centered_gamma=z, head_perm_cp=cp, peer_q_out=ws.q_ptrs,
                           peer_kv_out=ws.kv_ptrs, peer_row_offset=bounds[rank] * b, peer_heads=(hq_loc, hkv_loc))
        x_prepared = self.cross_attention_prepare(mixed_qqkv, key_value_states)
        ws.barrier()                                                     # A: every rank's q / kv rows have landed
        p = meta_args.core_attn_params
        assert meta_args.denoising_range_num == 1                        # dit_module.py:1048
        q_all, kv_all = ws.q[:total * b].unflatten(0, (total, b)), ws.kv[:total * b].unflatten(0, (total, b))
        tstride = b * Hq * F.HEAD_DIM
        for i in range(b):
            ptrs = [p_ + i * Hq * F.HEAD_DIM * 2 for p_ in ws.out_ptrs]
            kv_i = kv_all[:, i]
            F.attn_ranges(q_all[:, i], kv_i[..., :F.HEAD_DIM], kv_i[..., F.HEAD_DIM:], p.q_range[:1], p.k_range[:1],
                          return_lse=False, peer_out=(ptrs, bounds, head_off, Hq, tstride))
        xattn_out = self.cross_attention_attend(x_prepared, (sq, b), meta_args.cross_attn_params)
        ws.barrier()                                                     # B: every rank's output rows have landed
        core_attn_out = ws.out[:sq * b].reshape(sq, b, Hq * F.HEAD_DIM).clone()
        return core_attn_out, xattn_out

    # ------------------------------------------------------------------ CSO (dit_module.py:1207-1242)
    def _forward_cso(self, mixed_qqkv, key_value_states, inference_params, sin_emb, cos_emb, meta_args):
        """cp_shuffle_overlap: every rank holds the same slice of every (padded) chunk; K/V are exchanged once
        (head-sharded, un-shuffled and un-padded on arrival, straight into the KV-cache slot when all of it is
        cached), then chunk i's attention overlaps the query all-to-all of chunk i+1 and the output all-to-all of
        chunk i-1."""
        sq, b, _ = mixed_qqkv.shape
        assert b == 1                                       # dit_module.py:1070
        cp, dn = self.engine_config.cp_size, meta_args.denoising_range_num
        Hq, Hkv = self.model_config.num_attention_heads, self.model_config.num_query_groups
        eps, z = self.model_config.layernorm_epsilon, self.model_config.apply_layernorm_1p
        dev, dt = mixed_qqkv.device, mixed_qqkv.dtype
        split = [x * b for x in meta_args.cp_split_sizes]   # dit_module.py:1150
        rows = sq * b
        spb = divide(rows, dn)
        clip = meta_args.clip_token_nums
        hq_loc, hkv_loc = divide(Hq, cp), self.num_query_groups_per_partition
        x2d = mixed_qqkv.reshape(rows, -1)

        # K/V: prologue writes the all-to-all send layout "(cp spb) hn hd" (with KV-head replication when cp > Hkv)
        _, kv = F.qk_norm_rope(None, self.linear_qkv.forward_k(x2d), self.linear_qkv.forward_v(x2d), Hq, Hkv, None,
                               self._ln(self.k_layernorm), sin_emb, cos_emb, batch=b, eps=eps, zero_centered_gamma=z,
                               head_perm_cp=cp)
        kv_recv, handle_kv = cpm.cso_communication(kv, cp, split, "kv", send_layout=True)

        # Q: one send buffer per chunk, "(dn cp spb) hn hd", written in place by the prologue (one launch per chunk)
        xq = self.linear_qkv.forward_q(x2d)
        q_send = torch.empty((dn * cp * spb, hq_loc, F.HEAD_DIM), dtype=dt, device=dev)
        for d in range(dn):
            F.qk_norm_rope(xq[d * spb:(d + 1) * spb], None, None, Hq, Hkv, self._ln(self.q_layernorm), None,
                           sin_emb[d * spb:(d + 1) * spb], cos_emb[d * spb:(d + 1) * spb], batch=b, eps=eps,
                           zero_centered_gamma=z, head_perm_cp=cp, q_out=q_send[d * cp * spb:(d + 1) * cp * spb])
        helper = cpm.CSOHelper(dn, cp, split)
        query, handle_q = helper.split_query_for_overlap(q_send, send_layout=True)
        # the cross-attention needs no communication: its projections + norms run under the input all-to-alls
        x_prepared = self.cross_attention_prepare(mixed_qqkv, key_value_states)

        handle_kv.wait()
        total_rows = dn * clip
        cache, start, app = self._cache_plan(inference_params, meta_args, total_rows, dt, dev)
        offload = self._offloaded(cache)                   # pinned-host cache staged through KVOffload
        if cache is not None and app == total_rows and not offload:   # un-shuffle straight into the cache slot
            kv_new = F.cso_unshuffle(kv_recv, cp, dn, clip, out=cache[start:start + total_rows])
        else:
            kv_new = F.cso_unshuffle(kv_recv, cp, dn, clip)
            if offload:
                KVOffload.of(inference_params, dev).store(self.layer_number, cache, kv_new, start, app)
            elif cache is not None and app > 0:
                F.kv_cache_append(cache, kv_new, start, app)
        kc = vc = None
        if cache is not None and start > 0:
            kc, vc = (self._offload_views(inference_params, cache, start) if offload else
                      (cache[:start, :, :F.HEAD_DIM], cache[:start, :, F.HEAD_DIM:]))

        handle_q.wait()
        outs, handle_attn = helper.overlap(partial(self.full_attention, b, meta_args, k_cache=kc, v_cache=vc), query,
                                           kv_new[..., :F.HEAD_DIM], kv_new[..., F.HEAD_DIM:])
        xattn_out = self.cross_attention_attend(x_prepared, (sq, b), meta_args.cross_attn_params)
        handle_attn.wait()
        # "(dn cp sq b) hn hd -> (dn sq) b (cp hn hd)": one re-layout per chunk, written in place
        result = torch.empty((rows, Hq, F.HEAD_DIM), dtype=dt, device=dev)
        for d, o in enumerate(outs):
            F.ulysses_relayout(o, cp, 1, out=result[d * spb:(d + 1) * spb])
        return result.view(sq, b, Hq * F.HEAD_DIM), xattn_out

    # ------------------------------------------------------------------ ring-KV (new)
    _stripe_ranges = {}

    def _forward_ring_fused(self, mixed_qqkv, key_value_states, inference_params, sin_emb, cos_emb, meta_args):
        """cp_ring_kv without NCCL on the data path and without ring steps: the prologue kernel stores this rank's
        stripe of every in-flight chunk's [K|V] into EVERY rank's gathered buffer (NVLink peer stores, or one store
        to the NVSwitch multicast mapping), chunk-major - straight into the replicated KV-cache slot when every row of
        the forward is cached. One device-side barrier, then ONE attention launch of the local query stripe against
        cache ++ gathered K/V with the unmodified k ranges: the same per-row key order as the single-GPU forward,
        hence the same values, and no log-sum-exp merges (`_forward_ring` is the NCCL send/recv ring it replaces)."""
        sq, b, _ = mixed_qqkv.shape
        assert b == 1
        group, rank = mpu.get_cp_group(), mpu.get_cp_rank()
        Hq, Hkv = self.model_config.num_attention_heads, self.model_config.num_query_groups
        eps, z = self.model_config.layernorm_epsilon, self.model_config.apply_layernorm_1p
        dev, dt = mixed_qqkv.device, mixed_qqkv.dtype
        p = meta_args.core_attn_params
        clip, dn = meta_args.clip_token_nums, len(p.np_k_range)
        sizes = [s // dn for s in meta_args.cp_split_sizes]
        n_me, off_me = sizes[rank], sum(sizes[:rank])
        total_rows = dn * clip
        cache, start, app = self._cache_plan(inference_params, meta_args, total_rows, dt, dev)
        # kv_offload: the (replicated) cache lives in pinned host memory; the gathered K/V land in the workspace, go
        # back to the host on the copy stream and the cache prefix is staged through KVOffload like everywhere else
        offload = self._offloaded(cache)
        ws = cpm.RingGatherWorkspace.get(group, total_rows, Hkv, dev)
        buf, ptrs, mc = ws.next()
        direct = (cache is not None and app == total_rows and not offload
                  and getattr(self, "_cache_peer_of", None) == cache.data_ptr())
        if direct:
            slot = start * Hkv * 2 * F.HEAD_DIM * 2
            ptrs = [p_ + slot for p_ in self._cache_peer_ptrs]
            mc = self._cache_mc_ptr + slot if self._cache_mc_ptr else 0
        use_mc = bool(mc and cpm.ring_multicast_enabled())
        targets = [mc] if use_mc else ptrs
        query = xq_rows = x_kv = None
        if sq > 0:
            xq, xk, xv = self._project(mixed_qqkv)
            cross = self.cross_prologue_inputs(mixed_qqkv, key_value_states)
            query, _ = F.qk_norm_rope(xq, xk, xv, Hq, Hkv, self._ln(self.q_layernorm), self._ln(self.k_layernorm),
                                      sin_emb, cos_emb, batch=1, eps=eps, zero_centered_gamma=z,
                                      gather_kv_out=targets, gather_geometry=(n_me, clip, off_me),
                                      gather_multicast=use_mc, cross=cross)
            if cross is not None:
                xq_rows, x_kv = [cross["qx"]], cross["tkv"]
        x_prepared = self.cross_attention_prepare(mixed_qqkv, key_value_states, xq_rows, meta_args.cross_attn_params,
                                                  kv_x=x_kv)
        if offload:
            # the other workspace buffer is rewritten by the peers once they are past THIS barrier: the write-back that
            # still reads it (issued by the previous layer on the copy stream) has to be done before we enter
            off = KVOffload.of(inference_params, dev)
            if off.last_store is not None:
                torch.cuda.current_stream(dev).wait_event(off.last_store)
        ws.barrier()                                              # every rank's K/V stripes have landed everywhere
        kv = cache[start:start + total_rows] if direct else buf[:total_rows]
        if offload:
            off.store(self.layer_number, cache, kv, start, app)
        elif cache is not None and app > 0 and not direct:
            F.kv_cache_append(cache, kv, start, app)
        kc = vc = None
        if cache is not None and start > 0:
            kc, vc = (self._offload_views(inference_params, cache, start) if offload else
                      (cache[:start, :, :F.HEAD_DIM], cache[:start, :, F.HEAD_DIM:]))
        out = torch.empty((sq, Hq, F.HEAD_DIM), dtype=dt, device=dev)
        if sq > 0:
            key = (dn, n_me, str(dev))
            q_range = self._stripe_ranges.get(key)
            if q_range is None:    # local rows are (chunk, stripe row): chunk i = rows [i * n_me, (i + 1) * n_me)
                e = torch.arange(dn + 1, dtype=torch.int32) * n_me
                q_range = self._stripe_ranges[key] = torch.stack([e[:-1], e[1:]], 1).contiguous().to(dev)
            F.attn_ranges(query, kv[..., :F.HEAD_DIM], kv[..., F.HEAD_DIM:], q_range, p.k_range, k_cache=kc, v_cache=vc,
                          out=out, return_lse=False)
        xattn_out = self.cross_attention_attend(x_prepared, (sq, b), meta_args.cross_attn_params,
                                                independent=sq > 0)
        if direct and self.model_config.num_layers < 2:
            # the peers' next forward writes this layer's cache slot again: with a single layer there is no other
            # layer's barrier between this attention and that prologue
            ws.barrier()
        return out.view(sq, 1, Hq * F.HEAD_DIM), xattn_out

    def _forward_ring(self, mixed_qqkv, key_value_states, inference_params, sin_emb, cos_emb, meta_args):
        sq, b, _ = mixed_qqkv.shape
        if b != 1:
            # the batched uncond pass of the 3-CFG models (dit_model.py:413-494): one range shared by all batch rows, no
            # KV cache; its shards are the contiguous Ulysses shards (cp_ring_kv_process), so it runs as cp_ulysses
            if self._uses_cache(inference_params, meta_args) or meta_args.denoising_range_num != 1:
                raise NotImplementedError("cp_ring_kv with batch size > 1 supports the cache-less uncond pass only")
            if self._fused_ring(1, mixed_qqkv.device):
                return self._forward_ulysses_fused_batched(mixed_qqkv, key_value_states, sin_emb, cos_emb, meta_args)
            return self._forward_ulysses(mixed_qqkv, key_value_states, inference_params, sin_emb, cos_emb, meta_args)
        if self.engine_config.kv_offload:
            raise NotImplementedError("cp_ring_kv with kv_offload runs on the fused gather path only "
                                      "(cp_fused_all_to_all with NVLink peer access); the NCCL ring keeps the cache "
                                      "resident")
        Hq, Hkv = self.model_config.num_attention_heads, self.model_config.num_query_groups
        eps, z = self.model_config.layernorm_epsilon, self.model_config.apply_layernorm_1p
        dev, dt = mixed_qqkv.device, mixed_qqkv.dtype
        total_rows = sum(meta_args.cp_split_sizes)
        cache, start, app = self._cache_plan(inference_params, meta_args, total_rows, dt, dev)
        xq, xk, xv = self._project(mixed_qqkv)
        query, kv_local = F.qk_norm_rope(xq, xk, xv, Hq, Hkv, self._ln(self.q_layernorm), self._ln(self.k_layernorm),
                                         sin_emb, cos_emb, batch=1, eps=eps, zero_centered_gamma=z)
        p = meta_args.core_attn_params

        def attn(q, blk, cache_prefix, qr, kr):
            t = lambda a: torch.as_tensor(a, dtype=torch.int32).to(dev, non_blocking=True)
            kc = vc = None
            if cache_prefix is not None:
                kc, vc = cache_prefix[..., :F.HEAD_DIM], cache_prefix[..., F.HEAD_DIM:]
            return F.attn_ranges(q, blk[..., :F.HEAD_DIM], blk[..., F.HEAD_DIM:], t(qr), t(kr), k_cache=kc, v_cache=vc,
                                 zero_fill=True)

        clip = meta_args.clip_token_nums
        dn = len(p.np_k_range)
        chunk_ranges = [(int(a) // clip, int(e) // clip) for a, e in p.np_k_range]
        sizes = [s // dn for s in meta_args.cp_split_sizes]
        out, _ = cpm.RingKVScheduler.run(query, kv_local, cache, chunk_ranges, start // clip, clip, dn, sizes,
                                         app // clip, attn, F.lse_merge)
        core_attn_out = out.view(sq, 1, Hq * F.HEAD_DIM)
        xattn_out = self.cross_attention(mixed_qqkv, key_value_states, meta_args.cross_attn_params)
        return core_attn_out, xattn_out
