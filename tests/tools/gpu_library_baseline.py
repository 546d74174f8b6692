#!/usr/bin/env python
"""Context numbers on the B200 box (not product code): cuBLAS bf16 GEMM throughput on the encoder's own shapes, and
the img/s of the same ViT-B/16 math run through PyTorch's library kernels (cuDNN conv, cuBLASLt, FlashAttention-2 SDPA,
separate LN / GELU / add kernels) — what the reference modules dispatch to under .cuda().bfloat16().
Round 2 adds (VERDICT r01 "missing" 4): F.scaled_dot_product_attention with the flash / cuDNN / mem-efficient backends on
the four BASELINE attention shapes next to our kernel on the same box, and the same ViT math under torch.compile.
Writes gpurun_out/r2_library_baseline.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pytorch_models_b200 as pm  # noqa: E402
from bench import CONFIGS, synthetic_weights_  # noqa: E402
from oracle import oracle_torch  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


@torch.no_grad()
def main():
    out = {"gemm": {}, "model": {}}
    for name, (M, N, K) in {"qkv": (201728, 2304, 768), "out_proj": (201728, 768, 768), "fc1": (201728, 3072, 768),
                            "fc2": (201728, 768, 3072), "qkv_b128": (25216, 2304, 768)}.items():
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: torch.nn.functional.linear(a, w, b))
        out["gemm"][name] = dict(ms=ms, tflops=2.0 * M * N * K / ms * 1e-9)
        print("cublas", name, out["gemm"][name], flush=True)
        del a, w
    cfg = CONFIGS["c2"]
    torch.manual_seed(0)
    m = cfg["make"](pm).eval()
    synthetic_weights_(m, 100)
    sd = {k: v.cuda().bfloat16() for k, v in m.state_dict().items()}
    for B in (256, 1024):
        x = torch.randn(B, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: oracle_torch.vit_forward(sd, x, 12, "cls_token"), iters=5, warm=2)
        out["model"][f"torch_eager_bf16_b{B}"] = dict(ms=ms, img_s=B / ms * 1e3)
        print("torch eager bf16 ViT-B/16 batch", B, out["model"][f"torch_eager_bf16_b{B}"], flush=True)
        del x
    # ---- the reference's math under torch.compile (README.md:7 promises compile support; Inductor + Triton here)
    try:
        x = torch.randn(1024, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
        fn = torch.compile(lambda t: oracle_torch.vit_forward(sd, t, 12, "cls_token"))
        ms = timeit(lambda: fn(x), iters=5, warm=3)
        out["model"]["torch_compile_bf16_b1024"] = dict(ms=ms, img_s=1024 / ms * 1e3)
        del x
    except Exception as e:  # noqa: BLE001 - a missing host compiler / Triton failure must not hide the other numbers
        out["model"]["torch_compile_bf16_b1024"] = dict(error=f"{type(e).__name__}: {str(e)[:300]}")
    print("torch.compile", out["model"]["torch_compile_bf16_b1024"], flush=True)
    # ---- ours on the same batch
    x = torch.randn(1024, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
    mg = m.cuda().bfloat16()
    ms = timeit(lambda: mg(x), iters=10, warm=3)
    out["model"]["b200enc_bf16_b1024"] = dict(ms=ms, img_s=1024 / ms * 1e3)
    print("ours", out["model"]["b200enc_bf16_b1024"], flush=True)
    del x

    # ---- attention: F.scaled_dot_product_attention (transformer.py:52) per backend vs b200enc_attention, fused-QKV
    # layout exactly as the reference hands it over: (B, H, L, 64) views with strides (L*3d, 64, 3d, 1)
    from torch.nn.attention import SDPBackend, sdpa_kernel

    from pytorch_models_b200 import ops

    out["attention"] = {}
    shapes = {"c2_L197": (1024, 12, 197), "c3_L576": (256, 16, 576), "c4_L1370": (128, 16, 1370), "c5_L1500": (64, 20, 1500)}
    backends = {"flash": SDPBackend.FLASH_ATTENTION, "cudnn": SDPBackend.CUDNN_ATTENTION,
                "efficient": SDPBackend.EFFICIENT_ATTENTION}
    for name, (B, H, L) in shapes.items():
        D = H * 64
        qkv = torch.randn(B, L, 3 * D, device="cuda", dtype=torch.bfloat16)
        q, k, v = (qkv[:, :, i * D:(i + 1) * D] for i in range(3))
        heads = lambda t: t.unflatten(-1, (H, 64)).transpose(1, 2)  # noqa: E731
        fl = 4.0 * B * H * L * L * 64
        by = 8.0 * B * L * D
        row = {}
        o = torch.empty(B, L, D, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.attention(q, k, v, o, H, 0.125), iters=10)
        row["b200enc"] = dict(ms=ms, tflops=fl / ms * 1e-9, gbs=by / ms * 1e-6)
        ref = None
        for bname, be in backends.items():
            try:
                with sdpa_kernel(be):
                    fn = lambda: torch.nn.functional.scaled_dot_product_attention(heads(q), heads(k), heads(v))  # noqa: E731
                    y = fn()
                    ms = timeit(fn, iters=10)
                row[bname] = dict(ms=ms, tflops=fl / ms * 1e-9, gbs=by / ms * 1e-6)
                if ref is None:
                    ref = y.transpose(1, 2).flatten(-2)
                    row["max_abs_diff_vs_" + bname] = float((ref.float() - o.float()).abs().max())
            except Exception as e:  # noqa: BLE001
                row[bname] = dict(error=f"{type(e).__name__}: {str(e)[:200]}")
        # the layout-friendly case for the library: contiguous (B, H, L, 64) tensors (a copy the reference never makes)
        qc, kc, vc = (heads(t).contiguous() for t in (q, k, v))
        for bname, be in backends.items():
            try:
                with sdpa_kernel(be):
                    fn = lambda: torch.nn.functional.scaled_dot_product_attention(qc, kc, vc)  # noqa: E731
                    fn()
                    ms = timeit(fn, iters=10)
                row[bname + "_contiguous_heads"] = dict(ms=ms, tflops=fl / ms * 1e-9)
            except Exception as e:  # noqa: BLE001
                row[bname + "_contiguous_heads"] = dict(error=f"{type(e).__name__}: {str(e)[:200]}")
        best = min((v_["ms"] for k_, v_ in row.items() if isinstance(v_, dict) and "ms" in v_ and k_ != "b200enc"), default=None)
        row["speedup_vs_best_sdpa"] = None if best is None else best / row["b200enc"]["ms"]
        out["attention"][name] = row
        print("attention", name, json.dumps(row), flush=True)
        del qkv, qc, kc, vc, o
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_library_baseline.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
