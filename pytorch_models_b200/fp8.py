"""OPTIONAL FP8 (e4m3) variant of the linears — SURVEY §8(f) rank 4. Never used unless a caller asks for it.

The reference has no FP8 path (its linears are ``nn.Linear`` in the module's dtype, transformer.py:28-31,59,66); this
module exists because §8(f) lists an FP8 GEMM variant as headroom beyond the bf16 contract. It is not part of any
``forward`` of the model classes and not part of the headline benchmark. Scheme: per-tensor symmetric scaling to the
e4m3 range (max 448), fp32 accumulation in TMEM (``tcgen05.mma kind::f8f6f4``), the product of the two scales applied
in the GEMM epilogue together with bias / erf-GELU / residual. Stated tolerance (tests/test_gpu_fp8.py): the GEMM is
exact up to accumulation order against the dequantised operands (1e-2 relative); against the fp32 result of the
un-quantised operands the error is that of e4m3 rounding, a few per cent of the output RMS.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import ops

E4M3_MAX = 448.0


def quantize_e4m3(t: Tensor) -> tuple[Tensor, Tensor]:
    """Per-tensor symmetric quantisation: returns (e4m3 tensor, fp32 1-element dequantisation scale), no host sync."""
    amax = t.detach().abs().amax().float().clamp_min(1e-12)
    scale = (amax / E4M3_MAX).reshape(1)
    q = (t.detach().float() / scale).clamp_(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
    return q.contiguous(), scale


class Fp8Linear:
    """Kernel-ready e4m3 copy of an ``nn.Linear`` (quantised once) applied to dynamically quantised activations."""

    def __init__(self, linear: nn.Linear) -> None:
        self.w8, self.w_scale = quantize_e4m3(linear.weight)
        self.bias = None if linear.bias is None else linear.bias.detach().float().contiguous()
        self.out_features = linear.out_features

    def __call__(self, x: Tensor, *, gelu: bool = False, residual: Tensor | None = None) -> Tensor:
        """x: (M, K) bf16 CUDA -> (M, N) bf16."""
        x8, x_scale = quantize_e4m3(x)
        if x.dim() != 2:
            raise ValueError("Fp8Linear expects a (M, K) activation")
        out = torch.empty(x.shape[0], self.out_features, device=x.device, dtype=torch.bfloat16)
        return ops.linear_fp8(x8, self.w8, x_scale * self.w_scale, self.bias, out, gelu=gelu, residual=residual)


def mlp_forward_fp8(mlp: nn.Module, x: Tensor) -> Tensor:
    """``MLP`` (transformer.py:56-67) with both linears in FP8: linear1 + erf-GELU, linear2. x: (M, d) bf16."""
    if getattr(mlp, "_act_name", "gelu") != "gelu":
        raise NotImplementedError("the FP8 variant implements the erf-GELU MLP only")
    cache = mlp.__dict__.setdefault("_b200_fp8", {})
    key = (mlp.linear1.weight.data_ptr(), mlp.linear1.weight._version, mlp.linear2.weight.data_ptr(),
           mlp.linear2.weight._version)
    if cache.get("key") != key:
        cache.update(key=key, l1=Fp8Linear(mlp.linear1), l2=Fp8Linear(mlp.linear2))
    return cache["l2"](cache["l1"](x, gelu=True))
