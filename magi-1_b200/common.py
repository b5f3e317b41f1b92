"""Boundary types of the hot path, field-for-field with the reference so callers (the DiT model shell, the sampler)
do not change: inference/common/dataclass.py:22-83 and the config fields the attention block reads
(inference/common/config.py:20-107, read at dit_module.py:781-819, 1166-1244)."""
import dataclasses
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch


@dataclass(frozen=True)
class PackedCoreAttnParams:
    """dataclass.py:22-30. q_range/k_range: int32 (n, 2) device tensors; np_*: host mirrors."""
    q_range: torch.Tensor
    k_range: torch.Tensor
    np_q_range: np.ndarray
    np_k_range: np.ndarray
    max_seqlen_q: int
    max_seqlen_k: int


@dataclass(frozen=True)
class PackedCrossAttnParams:
    """dataclass.py:33-41."""
    q_ranges: torch.Tensor = None
    kv_ranges: torch.Tensor = None
    cu_seqlens_q: torch.Tensor = None
    cu_seqlens_kv: torch.Tensor = None
    max_seqlen_q: int = None
    max_seqlen_kv: int = None


@dataclass(frozen=True)
class ModelMetaArgs:
    """dataclass.py:44-59."""
    H: int
    W: int
    cp_pad_size: int
    cp_split_sizes: List[int]
    slice_point: int
    denoising_range_num: int
    range_num: int
    extract_prefix_video_feature: bool
    fwd_extra_1st_chunk: bool
    distill_nearly_clean_chunk: bool
    clip_token_nums: int
    enable_cuda_graph: bool
    core_attn_params: PackedCoreAttnParams
    cross_attn_params: PackedCrossAttnParams


class InferenceParams:
    """dataclass.py:62-83: per-generation KV-cache holder. key_value_memory_dict[layer_number] is the packed
    (max_seq * max_bs, Hkv_per_partition, 256) bf16 cache tensor, created lazily by the attention block."""

    def __init__(self, max_batch_size, max_sequence_length):
        self.max_sequence_length = max_sequence_length
        self.max_batch_size = max_batch_size
        self.sequence_len_offset = 0
        self.key_value_memory_dict = {}
        self.update_kv_cache = False


@dataclasses.dataclass
class ModelConfig:
    """The subset of config.py:20-47 the attention block reads; defaults are the 4.5B model
    (example/4.5B/4.5B_base_config.json:3-25)."""
    model_name: str = "videodit_ardf"
    num_layers: int = 34
    hidden_size: int = 3072
    ffn_hidden_size: int = 12288
    num_attention_heads: int = 24
    num_query_groups: int = 8
    kv_channels: int = 128
    layernorm_epsilon: float = 1e-6
    apply_layernorm_1p: bool = True
    params_dtype: torch.dtype = torch.bfloat16
    xattn_cond_hidden_ratio: float = 1.0
    caption_max_length: int = 800
    gated_linear_unit: bool = False
    cond_hidden_ratio: float = 0.25
    cond_gating_ratio: float = 1.0

    @classmethod
    def magi_4p5b(cls, **kw):
        return cls(**kw)

    @classmethod
    def magi_24b(cls, **kw):
        """example/24B/24B_base_config.json:3-25."""
        base = dict(num_layers=48, hidden_size=6144, ffn_hidden_size=16384, num_attention_heads=48, num_query_groups=8,
                    gated_linear_unit=True)
        base.update(kw)
        return cls(**base)


@dataclasses.dataclass
class EngineConfig:
    """config.py:88-107 (fields the hot path reads), plus `cp_ring_kv`, the strategy this build adds."""
    distributed_backend: str = "nccl"
    pp_size: int = 1
    cp_size: int = 1
    cp_strategy: str = "none"  # "none" | "cp_ulysses" | "cp_shuffle_overlap" | "cp_ring_kv" (new)
    ulysses_overlap_degree: int = 1  # head groups of the NCCL Ulysses path; the fused path overlaps by itself and ignores it
    fp8_quant: bool = False
    distill_nearly_clean_chunk_threshold: float = 0.3
    distill: bool = False
    kv_offload: bool = False
    enable_cuda_graph: bool = False
    # new: under cp_ulysses, fuse both all-to-alls into the kernels' own stores over NVLink peer (symmetric) memory
    # instead of calling NCCL (falls back to NCCL when symmetric memory is unavailable, e.g. gloo / CPU tests)
    cp_fused_all_to_all: bool = True
    # new: with cp_strategy "none", run the q / k / v / qx projections and the QK-LayerNorm + RoPE
    # prologue as ONE tensor-core GEMM whose epilogue is the prologue (magi_qkv_proj_norm_rope_fwd) instead of four
    # cuBLAS GEMMs + the prologue kernels. Measured on a B200 (profiles/r02b_gemm_bench.log): 1338 vs 1523 us per layer at
    # cfg2 (4.5B, 37 440 rows), 3938 vs 4196 us at the 24B shapes; results equal to the unfused path up to single bf16
    # ulps on 0.001 % of the elements.
    fused_qkv_proj: bool = True


def divide(numerator: int, denominator: int) -> int:
    """common_utils.py `divide`: exact integer division."""
    if numerator % denominator != 0:
        raise ValueError(f"{numerator} is not divisible by {denominator}")
    return numerator // denominator
