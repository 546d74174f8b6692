"""GPU tests of the OPTIONAL FP8 (e4m3) linear variant (SURVEY §8 f rank 4). Its own tolerance, stated here:

* against fp64 on the *dequantised* operands the kernel is exact up to fp32 accumulation order: |err| <= 1e-2 * |ref| + 1e-2
  (the bf16 rounding of the output dominates);
* against the fp32 result of the original (un-quantised) operands the error is e4m3 rounding of both operands:
  relative RMS error <= 6 % for unit-normal data, measured ~3.5 %.
The bf16 path is the product; nothing here feeds the headline benchmark."""
from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (300, 768, 3072), (1000, 3072, 768), (257, 576, 208)])
@pytest.mark.parametrize("mode", ["bias", "gelu", "residual"])
def test_linear_fp8_against_dequantised_operands(M, N, K, mode):
    from pytorch_models_b200 import ops
    from pytorch_models_b200.fp8 import quantize_e4m3

    torch.manual_seed(M + N + K)
    x, w = torch.randn(M, K, device="cuda"), 0.05 * torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16() if mode == "residual" else None
    x8, sx = quantize_e4m3(x)
    w8, sw = quantize_e4m3(w)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.linear_fp8(x8, w8, sx * sw, bias, out, gelu=mode == "gelu", residual=res)
    ref = (x8.double() * sx.double()) @ (w8.double() * sw.double()).T + bias.double()
    if mode == "gelu":
        ref = F.gelu(ref)
    if res is not None:
        ref = ref + res.double()
    err = (out.double() - ref).abs()
    assert bool((err <= 1e-2 * ref.abs() + 1e-2).all()), float(err.max())


def test_mlp_fp8_error_is_e4m3_rounding():
    import pytorch_models_b200 as pm
    from pytorch_models_b200.fp8 import mlp_forward_fp8

    torch.manual_seed(0)
    mlp = pm.MLP(768, 3072).eval().cuda()
    x = torch.randn(2048, 768, device="cuda")
    with torch.no_grad():
        want = mlp.linear2(F.gelu(mlp.linear1(x)))                      # fp32 reference math
        got_bf16 = mlp(x.bfloat16()).float()                            # the product path
        got_fp8 = mlp_forward_fp8(mlp, x.bfloat16()).float()
    rms = want.pow(2).mean().sqrt()
    e_bf16 = (got_bf16 - want).pow(2).mean().sqrt() / rms
    e_fp8 = (got_fp8 - want).pow(2).mean().sqrt() / rms
    assert e_fp8 <= 0.06 and e_bf16 <= 0.01, (float(e_fp8), float(e_bf16))


def test_fp8_rejects_unsupported_combinations():
    from pytorch_models_b200 import ops
    from pytorch_models_b200.fp8 import quantize_e4m3

    x8, sx = quantize_e4m3(torch.randn(64, 72, device="cuda"))  # K = 72 is not a multiple of 16
    w8, sw = quantize_e4m3(torch.randn(64, 72, device="cuda"))
    with pytest.raises(ValueError):
        ops.linear_fp8(x8, w8, sx * sw, None, torch.empty(64, 64, device="cuda", dtype=torch.bfloat16))
