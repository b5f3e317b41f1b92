"""CUDA-graph execution of the DiT block stack (EngineConfig.enable_cuda_graph, inference/common/config.py:107).

The reference carries the flag through ModelMetaArgs.enable_cuda_graph (dit_model.py:334) but never captures anything
in this snapshot (xattn_mask_for_cuda_graph is always None, dit_model.py:210). Here it is honoured: every entry point of
libmagi_attn.so is asynchronous on the stream it is given, never synchronises, never reads ranges on the host and
allocates nothing behind the caller's back (include/magi_attn.h), and the per-forward host logic of the blocks depends
only on shapes and on the host mirrors of the ranges - so one denoise forward of the whole block stack is ONE graph per
forward *signature*, replayed for every later forward with that signature (the tens of denoising steps of a stage).
What it buys is the launch-bound regime: small latent grids (BASELINE configs[0]) and high context-parallel degrees,
where ~7 launches per block through Python cost more than the kernels.

    stack = GraphedLayerStack(layers, engine_config)
    x = stack(hidden_states, condition, condition_map, y_xattn_flat, rotary_pos_emb, inference_params, meta_args)

The returned tensor is the graph's static output buffer: consume (or copy) it before the next call with the same
signature.
"""
import dataclasses
from typing import Dict, List, Optional

import numpy as np
import torch

from .common import EngineConfig, InferenceParams, ModelMetaArgs


class _Captured:
    __slots__ = ("graph", "static", "meta", "out")


def _np_key(a) -> bytes:
    return b"" if a is None else np.ascontiguousarray(a).tobytes()


class GraphedLayerStack:
    def __init__(self, layers: List[torch.nn.Module], engine_config: Optional[EngineConfig] = None):
        self.layers = list(layers)
        self.engine_config = engine_config if engine_config is not None else self.layers[0].engine_config
        self._graphs: Dict[tuple, _Captured] = {}
        self.captures = self.replays = 0
        self.disabled_reason: Optional[str] = None

    # ------------------------------------------------------------------ eager path (also the capture body)
    def _run(self, hs, cond, cmap, y, rope, inference_params, meta_args):
        x = hs
        for layer in self.layers:
            x = layer(x, cond, cmap, y, rope, inference_params, meta_args)
        return x

    @staticmethod
    def _signature(hs, cond, cmap, y, rope, inf: Optional[InferenceParams], m: ModelMetaArgs) -> tuple:
        core, cross = m.core_attn_params, m.cross_attn_params
        shape = lambda t: None if t is None else (tuple(t.shape), str(t.dtype))
        return (shape(hs), shape(cond), shape(cmap), shape(y), shape(rope), id(inf),
                None if inf is None else (inf.update_kv_cache, inf.max_batch_size, inf.max_sequence_length),
                m.slice_point, m.denoising_range_num, m.range_num, m.clip_token_nums, m.cp_pad_size,
                None if m.cp_split_sizes is None else tuple(int(v) for v in m.cp_split_sizes),
                m.extract_prefix_video_feature, m.fwd_extra_1st_chunk, m.distill_nearly_clean_chunk,
                _np_key(core.np_q_range), _np_key(core.np_k_range), core.max_seqlen_q, core.max_seqlen_k,
                shape(cross.q_ranges), shape(cross.kv_ranges), shape(cross.cu_seqlens_q), shape(cross.cu_seqlens_kv),
                cross.max_seqlen_q, cross.max_seqlen_kv)

    @staticmethod
    def _static_meta(m: ModelMetaArgs) -> ModelMetaArgs:
        clone = lambda t: None if t is None else t.clone()
        core, cross = m.core_attn_params, m.cross_attn_params
        core_s = dataclasses.replace(core, q_range=clone(core.q_range), k_range=clone(core.k_range))
        cross_s = dataclasses.replace(cross, q_ranges=clone(cross.q_ranges), kv_ranges=clone(cross.kv_ranges),
                                      cu_seqlens_q=clone(cross.cu_seqlens_q), cu_seqlens_kv=clone(cross.cu_seqlens_kv))
        return dataclasses.replace(m, core_attn_params=core_s, cross_attn_params=cross_s)

    @staticmethod
    def _load_meta(dst: ModelMetaArgs, src: ModelMetaArgs) -> None:
        pairs = ((dst.core_attn_params.q_range, src.core_attn_params.q_range),
                 (dst.core_attn_params.k_range, src.core_attn_params.k_range),
                 (dst.cross_attn_params.q_ranges, src.cross_attn_params.q_ranges),
                 (dst.cross_attn_params.kv_ranges, src.cross_attn_params.kv_ranges),
                 (dst.cross_attn_params.cu_seqlens_q, src.cross_attn_params.cu_seqlens_q),
                 (dst.cross_attn_params.cu_seqlens_kv, src.cross_attn_params.cu_seqlens_kv))
        for d, s_ in pairs:
            if d is not None and d.data_ptr() != s_.data_ptr():
                d.copy_(s_, non_blocking=True)

    def _capture(self, key, tensors, inference_params, meta_args) -> _Captured:
        c = _Captured()
        c.static = [None if t is None else t.clone() for t in tensors]
        c.meta = self._static_meta(meta_args)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up outside capture: KV caches, workspaces and cuBLAS handles get allocated
            self._run(*c.static, inference_params, c.meta)
        torch.cuda.current_stream().wait_stream(side)
        # the attention work lists built during warm-up are dropped so that their (tiny) kernels are part of the graph:
        # a replay then rebuilds them from the static range tensors it has just been given
        for holder in (c.meta.core_attn_params, c.meta.cross_attn_params):
            holder.__dict__.pop("_magi_plans", None)
        c.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(c.graph):
            c.out = self._run(*c.static, inference_params, c.meta)
        self._graphs[key] = c
        self.captures += 1
        return c

    @torch.no_grad()
    def __call__(self, hidden_states, condition, condition_map, y_xattn_flat, rotary_pos_emb,
                 inference_params: Optional[InferenceParams], meta_args: ModelMetaArgs):
        tensors = (hidden_states, condition, condition_map, y_xattn_flat, rotary_pos_emb)
        if not self.engine_config.enable_cuda_graph or self.disabled_reason is not None:
            return self._run(*tensors, inference_params, meta_args)
        key = self._signature(*tensors, inference_params, meta_args)
        c = self._graphs.get(key)
        if c is None:
            try:
                c = self._capture(key, tensors, inference_params, meta_args)
            except Exception as e:  # e.g. a collective that cannot be captured: stay correct, run eagerly from now on
                self.disabled_reason = f"{type(e).__name__}: {e}"[:300]
                torch.cuda.synchronize()
                return self._run(*tensors, inference_params, meta_args)
        else:
            for d, s_ in zip(c.static, tensors):
                if d is not None and d.data_ptr() != s_.data_ptr():
                    d.copy_(s_, non_blocking=True)
            self._load_meta(c.meta, meta_args)
        c.graph.replay()
        self.replays += 1
        return c.out
