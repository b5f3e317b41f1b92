"""Drop-in fp8 linears of the reference's `fp8_quant` mode (inference/model/dit/dit_module.py:431-488; used by the
q / qx / k / v projections :405-410, the MLP :520-540 and the attention out-projection :873-880 of every layer except
the first and the last when `engine_config.fp8_quant`, e.g. example/4.5B/4.5B_distill_quant_config.json): same
constructor arguments, same parameter names / shapes / dtypes (the calibrated `.fp8` checkpoints load), same forward.

What runs where
  * the activation quantiser `div_clamp_to` (:364-384): one launch of the hand-written kernel behind
    `magi_fp8_quantize` (3 bytes of HBM traffic per element instead of the eager chain's ~17), byte-identical;
  * the GEMM: where the reference calls the library (`flashinfer.bmm_fp8` -> cuBLASLt) this calls the library
    (`torch._scaled_mm` -> the same cuBLASLt fp8 matmul with per-tensor scale pointers) - except for q / qx / k / v,
    which `CustomLayerNormLinear.fused_operands` routes through the hand-written tcgen05 `kind::f8f6f4` GEMM whose
    epilogue is the QK-LayerNorm + RoPE prologue (attention.py).
"""
import torch
import torch.nn as nn
from torch.nn import Parameter

from . import functional as F


class PerTensorQuantizedFp8Linear(nn.Module):
    """dit_module.py:431-457. `bias` and `device` are accepted and ignored, like there."""

    def __init__(self, in_features: int, out_features: int, bias=False, dtype=torch.bfloat16, device=None) -> None:
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.finfo = torch.finfo(torch.float8_e4m3fn)
        self.output_dtype = dtype
        self.weight = Parameter(torch.empty((1, out_features, in_features), dtype=torch.float8_e4m3fn))
        self.weight_scale = Parameter(torch.empty(1, dtype=torch.float32))
        self.input_scale = Parameter(torch.empty(in_features, dtype=torch.float32))

    def quantize_input(self, input: torch.Tensor) -> torch.Tensor:
        return F.fp8_quantize(input, self.input_scale)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return F.fp8_scaled_linear(self.quantize_input(input), self.weight[0], self.input_scale, self.weight_scale,
                                   self.output_dtype)


class PerChannelQuantizedFp8Linear(nn.Module):
    """dit_module.py:462-488: the input is divided by the per-input-channel `smooth_scale` before the quantisation;
    `input_scale` and `weight_scale` are the per-tensor scalars of the GEMM."""

    def __init__(self, in_features: int, out_features: int, bias=False, dtype=torch.bfloat16, device=None) -> None:
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.output_dtype = dtype
        self.finfo = torch.finfo(torch.float8_e4m3fn)
        self.weight = Parameter(torch.empty((1, out_features, in_features), dtype=torch.float8_e4m3fn))
        self.weight_scale = Parameter(torch.empty(1, dtype=torch.float32))
        self.input_scale = Parameter(torch.empty(1, dtype=torch.float32))
        self.smooth_scale = Parameter(torch.empty(1, in_features, dtype=torch.float32))

    def quantize_input(self, x: torch.Tensor) -> torch.Tensor:
        return F.fp8_quantize(x, self.smooth_scale)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.fp8_scaled_linear(self.quantize_input(x), self.weight[0], self.input_scale, self.weight_scale,
                                   self.output_dtype)


def uses_fp8(engine_config, model_config, layer_number: int) -> bool:
    """Which layers hold fp8 linears (dit_module.py:407, 522, 535, 874-876): all but the first and the last."""
    return bool(engine_config.fp8_quant) and layer_number != 0 and layer_number != model_config.num_layers - 1
