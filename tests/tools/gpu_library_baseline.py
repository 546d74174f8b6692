#!/usr/bin/env python
"""Context numbers on the B200 box (not product code): cuBLAS bf16 GEMM throughput on the encoder's own shapes, and
the img/s of the same ViT-B/16 math run through PyTorch's library kernels (cuDNN conv, cuBLASLt, FlashAttention-2 SDPA,
separate LN / GELU / add kernels) — what the reference modules dispatch to under .cuda().bfloat16().
Writes gpurun_out/library_baseline.json."""
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
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "library_baseline.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
