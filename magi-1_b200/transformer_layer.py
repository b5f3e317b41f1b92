"""Drop-in `TransformerLayer` for the reference's DiT block (inference/model/dit/dit_module.py:1252-1367), the caller
on either side of the attention hot path: same constructor
`(model_config, engine_config, layer_number)`, same sub-module / parameter names (reference state_dicts load:
`ada_modulate_layer.proj.0`, `self_attention.*`, `self_attn_post_norm`, `mlp.{layer_norm,linear_fc1,linear_fc2}`,
`mlp_post_norm`), same `forward(hidden_states, condition, condition_map, y_xattn_flat, rotary_pos_emb,
inference_params, meta_args)`.

What runs where
  * the attention hot path: hand-written CUDA (attention.py / functional.py);
  * both gating chains (`bias_modulate_add`, dit_module.py:292-307): one fused kernel launch each;
  * the out-projection input "sq b (hn n hd)": written in place by the two attention kernels (no cat / rearrange);
  * q / qx / k / v: one fused tcgen05 GEMM inside the attention module (attention.py); GELU / SwiGLU: one launch of
    magi_mlp_activation; under fp8_quant the activation quantiser in front of every fp8 linear: magi_fp8_quantize;
  * the other dense GEMMs (kv_xattn, linear_proj, MLP fc1 / fc2; cuBLASLt fp8 under fp8_quant, as in the reference), the
    hidden-size LayerNorms in front of them and the tiny adaLN projection stay plain torch / cuBLAS library calls.
"""
import torch
import torch.nn as nn

from . import functional as F
from .attention import FullyParallelAttention, HeadLayerNorm
from .common import EngineConfig, InferenceParams, ModelConfig, ModelMetaArgs
from .fp8_linear import PerChannelQuantizedFp8Linear, PerTensorQuantizedFp8Linear, uses_fp8


class AdaModulateLayer(nn.Module):
    """dit_module.py:177-195: SiLU -> Linear(hidden * cond_hidden_ratio -> hidden * cond_gating_ratio * 2)."""

    def __init__(self, model_config: ModelConfig):
        super().__init__()
        self.model_config = model_config
        self.gate_num_chunks = 2
        self.act = nn.SiLU()
        self.proj = nn.Sequential(nn.Linear(int(model_config.hidden_size * model_config.cond_hidden_ratio),
                                            int(model_config.hidden_size * model_config.cond_gating_ratio
                                                * self.gate_num_chunks),
                                            bias=True, dtype=model_config.params_dtype))

    def forward(self, c):
        return self.proj(self.act(c))


def softcap(x: torch.Tensor, cap: float):
    """dit_module.py:360-361."""
    return (cap * torch.tanh(x.float() / cap)).to(x.dtype)


class CustomMLP(nn.Module):
    """dit_module.py:493-551: LayerNorm -> fc1 -> GELU, or SwiGLU (`flashinfer.activation.silu_and_mul`:
    silu(x[..., :d]) * x[..., d:]) when gated_linear_unit -> fc2. On the layers `fp8_quant` covers fc1 is a
    PerTensorQuantizedFp8Linear and fc2 a PerChannelQuantizedFp8Linear (:520-540)."""

    def __init__(self, model_config: ModelConfig, engine_config: EngineConfig, layer_number: int):
        super().__init__()
        self.model_config = model_config
        dt, h, f = model_config.params_dtype, model_config.hidden_size, model_config.ffn_hidden_size
        self.layer_norm = nn.LayerNorm(h, eps=model_config.layernorm_epsilon, dtype=dt)
        fp8 = uses_fp8(engine_config, model_config, layer_number)
        fc1_cls = PerTensorQuantizedFp8Linear if fp8 else nn.Linear
        fc2_cls = PerChannelQuantizedFp8Linear if fp8 else nn.Linear
        self.linear_fc1 = fc1_cls(h, 2 * f if model_config.gated_linear_unit else f, bias=False, dtype=dt)
        self.linear_fc2 = fc2_cls(f, h, bias=False, dtype=dt)

    def forward(self, hidden_states):
        x = self.linear_fc1(self.layer_norm(hidden_states))
        # one launch of magi_mlp_activation: silu_and_mul (single rounding, like the flashinfer kernel the reference
        # calls) or erf-GELU
        return self.linear_fc2(F.mlp_activation(x, self.model_config.gated_linear_unit))


class TransformerLayer(nn.Module):
    def __init__(self, model_config: ModelConfig, engine_config: EngineConfig, layer_number: int = 1):
        super().__init__()
        self.model_config = model_config
        self.engine_config = engine_config
        self.layer_number = layer_number          # pp_size == 1 on this path: no layer offset (dit_module.py:1284-1295)
        self.ada_modulate_layer = AdaModulateLayer(model_config)
        self.self_attention = FullyParallelAttention(model_config, engine_config, self.layer_number)
        self.self_attn_post_norm = HeadLayerNorm(model_config, model_config.hidden_size)
        self.mlp = CustomMLP(model_config, engine_config, self.layer_number)
        self.mlp_post_norm = HeadLayerNorm(model_config, model_config.hidden_size)

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, condition: torch.Tensor, condition_map: torch.Tensor,
                y_xattn_flat: torch.Tensor, rotary_pos_emb: torch.Tensor, inference_params: InferenceParams,
                meta_args: ModelMetaArgs):
        residual = hidden_states
        attn = self.self_attention
        if attn.supports_fused_proj_input(hidden_states.shape[1], hidden_states.device):
            # both attention kernels write straight into the out-projection input (dit_module.py:1332-1338); the
            # "(n hn hd) -> (hn n hd)" groups are whole heads when Hq is a multiple of hn = 8 (24 and 48 are)
            proj_in = attn(hidden_states, y_xattn_flat, inference_params, rotary_pos_emb, meta_args,
                           fused_proj_input=True)
        else:
            core_attn_out, cross_attn_out = attn(hidden_states, y_xattn_flat, inference_params, rotary_pos_emb,
                                                 meta_args)
            proj_in = FullyParallelAttention.attn_linear_proj_input(core_attn_out, cross_attn_out)
        return self.attn_post_process_from_proj_input(proj_in, residual, condition, condition_map)

    def attn_post_process(self, core_attn_out, cross_attn_out, residual, condition, condition_map):
        """dit_module.py:1320-1330, for callers that hold the two attention outputs."""
        proj_in = FullyParallelAttention.attn_linear_proj_input(core_attn_out, cross_attn_out)
        return self.attn_post_process_from_proj_input(proj_in, residual, condition, condition_map)

    def attn_post_process_from_proj_input(self, proj_in, residual, condition, condition_map):
        # attn_linear_proj, dit_module.py:1332-1346. The reference wraps the bf16 projection in
        # torch.autocast("cuda", dtype=torch.float32), which CUDA autocast does not support (it warns and disables
        # itself), so the projection runs in the parameters' dtype, as here; the fp8 out-projection
        # (adapt_linear_quant, :1339-1340) is called as it is.
        attn_out = self.self_attention.linear_proj(proj_in)
        return self.gating_and_mlp(attn_out, residual, condition, condition_map)

    def gating_and_mlp(self, hidden_states, residual, condition, condition_map):
        """dit_module.py:1348-1367 with both bias_modulate_add chains fused."""
        gate_output = softcap(self.ada_modulate_layer(condition), 1.0)
        gate_msa, gate_mlp = gate_output.chunk(2, dim=-1)
        hidden_states = F.bias_modulate_add(hidden_states, residual, condition_map, gate_msa, self.self_attn_post_norm)
        residual = hidden_states
        hidden_states = self.mlp(hidden_states)
        return F.bias_modulate_add(hidden_states, residual, condition_map, gate_mlp, self.mlp_post_norm)


class FusedLayerNorm(HeadLayerNorm):
    """dit_module.py:323-357 as a callable module (the block's final norm): LayerNorm with `weight + 1` when
    apply_layernorm_1p. The per-head and post-attention / post-MLP instances of the reference class are parameter
    holders here (HeadLayerNorm) because their arithmetic lives inside the fused kernels."""

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        weight = self.weight + 1 if self.zero_centered_gamma else self.weight
        return torch.nn.functional.layer_norm(input, tuple(self.weight.shape), weight, self.bias, self.eps)


class TransformerBlock(nn.Module):
    """Drop-in for the reference's `TransformerBlock` (dit_module.py:1373-1441), the caller of the layers: same
    constructor `(model_config, engine_config, pre_process, post_process)`, same sub-module names (`layers.<i>.*`,
    `final_layernorm`), same `set_input_tensor` / `forward`. pp_size == 1 on this path, so the block holds all
    `num_layers` layers (`:1390`) and layer numbers carry no offset. With `engine_config.enable_cuda_graph` the layer
    stack runs as one CUDA graph per forward signature (cuda_graph.GraphedLayerStack); the final fp32 LayerNorm
    (`:1438-1439`) stays outside the graph, so the returned tensor is the caller's own."""

    def __init__(self, model_config: ModelConfig, engine_config: EngineConfig, pre_process: bool = True,
                 post_process: bool = True):
        super().__init__()
        self.model_config = model_config
        self.engine_config = engine_config
        self.pre_process = pre_process
        self.post_process = post_process
        self.input_tensor = None          # required for pipeline parallel schedules (dit_module.py:1387-1388)
        if engine_config.pp_size != 1:
            raise NotImplementedError("pipeline parallelism is outside this build (pp_size must be 1)")
        self.layers = nn.ModuleList([TransformerLayer(model_config, engine_config, layer_number=i)
                                     for i in range(model_config.num_layers)])
        if self.post_process:
            self.final_layernorm = FusedLayerNorm(model_config, model_config.hidden_size)
        self._graphed = None

    def set_input_tensor(self, input_tensor: torch.Tensor):
        """dit_module.py:1403-1411."""
        self.input_tensor = input_tensor

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, condition: torch.Tensor, condition_map: torch.Tensor,
                y_xattn_flat: torch.Tensor, rotary_pos_emb: torch.Tensor, inference_params: InferenceParams,
                meta_args: ModelMetaArgs) -> torch.Tensor:
        if not self.pre_process:
            assert self.input_tensor is not None, "please call set_input_tensor for pp"
            hidden_states = self.input_tensor
        if self.engine_config.enable_cuda_graph and hidden_states.is_cuda:
            if self._graphed is None:
                from .cuda_graph import GraphedLayerStack
                self._graphed = GraphedLayerStack(list(self.layers), self.engine_config)
            hidden_states = self._graphed(hidden_states, condition, condition_map, y_xattn_flat, rotary_pos_emb,
                                          inference_params, meta_args)
        else:
            for layer in self.layers:
                hidden_states = layer(hidden_states, condition, condition_map, y_xattn_flat, rotary_pos_emb,
                                      inference_params, meta_args)
        if self.post_process:
            hidden_states = self.final_layernorm(hidden_states.float())
        return hidden_states


def high_precision_promote(module: nn.Module) -> nn.Module:
    """The part of `_high_precision_promoter` (dit_model.py:621-638) that concerns the transformer block: the
    self-attention q / k LayerNorms (not the `_xattn` ones), both post-norms and the block's final norm hold fp32
    parameters."""
    for name, sub in module.named_modules():
        if "_xattn" in name:
            continue
        if any(k in name for k in ("q_layernorm", "k_layernorm", "self_attn_post_norm", "mlp_post_norm",
                                   "final_layernorm")):
            sub.float()
    return module
