#!/usr/bin/env python
"""Drive every HBM-bound row kernel of libb200enc once per realistic shape (for `ncu` captures and CUDA-event timing):
layernorm, row_stats, mean_tokens, patch_rows (bf16 / fp32, p=16 / p=14), cls_rows, time_rows, embed_rows, whisper_logmel.
Prints achieved algorithmic GB/s per kernel and writes gpurun_out/r2_row_kernels.json. Not product code."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pytorch_models_b200 import ops  # noqa: E402

dev = "cuda"
bf = torch.bfloat16


def timeit(fn, iters=10, warm=2):
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
    iters = int(os.environ.get("ROW_ITERS", "10"))
    out = {}

    def rec(name, nbytes, fn):
        ms = timeit(fn, iters)
        out[name] = dict(ms=ms, algorithmic_bytes=nbytes, gbs=nbytes / ms * 1e-6)
        print(f"{name:34s} {ms:8.4f} ms  {out[name]['gbs']:8.1f} GB/s (algorithmic)", flush=True)

    # final norm of a gap / siglip model and BERT post-norm: all token rows (C2: 1024 x 197 rows of 768)
    M, d = 1024 * 197, 768
    x = torch.randn(M, d, device=dev).to(bf)
    y = torch.empty_like(x)
    g, b = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    st = torch.empty(M, 2, device=dev)
    rec("layernorm_201728x768", 4.0 * M * d, lambda: ops.layernorm(x, g, b, 1e-6, y))
    rec("row_stats_201728x768", 2.0 * M * d, lambda: ops.row_stats(x, 1e-6, st))
    x3 = x.view(1024, 197, d)
    pooled = torch.empty(1024, d, device=dev, dtype=bf)
    rec("mean_tokens_1024x197x768", 2.0 * M * d, lambda: ops.mean_tokens(x3, pooled))
    rec("layernorm_cls_rows_1024x768", 4.0 * 1024 * d, lambda: ops.layernorm(x3[:, 0, :], g, b, 1e-6, pooled))
    cls = torch.randn(1, 1, d, device=dev).to(bf)
    rec("cls_rows_1024x768", 2.0 * 1024 * d, lambda: ops.cls_rows(cls, x3))
    del x, y, x3
    # patch rows (the im2col view of the patch-embedding conv)
    for name, B, HW, p, dt in (("patch_rows_bf16_p16_b1024", 1024, 224, 16, bf), ("patch_rows_f32_p16_b1024", 1024, 224, 16, torch.float32),
                               ("patch_rows_bf16_p14_b128_518", 128, 518, 14, bf)):
        img = torch.randn(B, 3, HW, HW, device=dev).to(dt)
        kp = (3 * p * p + 7) // 8 * 8
        rows = torch.empty(B, (HW // p) ** 2, kp, device=dev, dtype=bf)
        rec(name, img.numel() * img.element_size() + rows.numel() * 2.0, lambda: ops.patch_rows(img, p, kp, rows))
        del img, rows
    # whisper stem input transpose, token embedding, audio front end
    mel = torch.randn(64, 128, 3000, device=dev)
    tr = torch.empty(64, 3002, 128, device=dev, dtype=bf)
    rec("time_rows_64x128x3000_f32", mel.numel() * 4.0 + tr.numel() * 2.0, lambda: ops.time_rows(mel, tr))
    ids = torch.randint(0, 50257, (32, 1024), device=dev)
    tok, pos = torch.randn(50257, 768, device=dev).to(bf), torch.randn(1024, 768, device=dev).to(bf)
    emb = torch.empty(32, 1024, 768, device=dev, dtype=bf)
    rec("embed_rows_32x1024x768", 3 * 2.0 * emb.numel(), lambda: ops.embed_rows(ids, tok, pos, emb))
    audio = 0.3 * torch.randn(16, 480000, device=dev)
    import pytorch_models_b200 as pm

    pre = pm.WhisperPreprocessor("large-v3").cuda()
    rec("whisper_logmel_16x30s", audio.numel() * 4.0 + 16 * 128 * 3000 * 4.0 * 2, lambda: pre(audio))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_row_kernels.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
