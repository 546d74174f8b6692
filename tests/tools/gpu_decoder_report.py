#!/usr/bin/env python
"""Decoder-path report on the GPU box (DecoderLayer / Decoder, GPT-2, Whisper decoder):
parity of the sm_100a path and of PyTorch's own bf16 forward against the fp32 CPU oracle on identical seeded weights,
and tokens/s of both on the GPU. Test infrastructure (uses oracle/). Writes gpurun_out/decoder_report.json.

    gpt2          GPT-2 small (12 x 768, vocab 50257), 1024-token sequences       gpt2.py:11-27
    whisper_dec   Whisper large-v3 decoder (32 x 1280, vocab 51866), 448 tokens   whisper.py:37-53
                  against a 1500-frame encoder memory
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pytorch_models_b200 as pm  # noqa: E402
from bench import synthetic_weights_  # noqa: E402
from oracle import oracle_torch  # noqa: E402


def timeit(fn, iters: int, warm: int = 2) -> float:
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rel_err(got: torch.Tensor, want: torch.Tensor) -> dict:
    """Logits are not unit-RMS: report the error relative to the RMS of the reference logits, and cosine per row."""
    got, want = got.double().flatten(0, -2), want.double().flatten(0, -2)
    rms = want.pow(2).mean().sqrt().item()
    cos = torch.nn.functional.cosine_similarity(got, want, dim=-1).min().item()
    agree = (got.argmax(-1) == want.argmax(-1)).double().mean().item()
    return dict(max_abs_over_rms=(got - want).abs().max().item() / rms, rms=rms, min_cos=cos, argmax_agreement=agree)


@torch.no_grad()
def main() -> None:
    report = {}
    dev = "cuda"
    cases = {
        "gpt2": dict(make=lambda: pm.GPT2.from_hf("gpt2"), batch=32, L=1024, parity_batch=2, parity_L=256),
        "whisper_dec": dict(make=lambda: pm.WhisperDecoder(51866, 32, 1280), batch=64, L=448, Lm=1500,
                            parity_batch=1, parity_L=64),
    }
    for name in sys.argv[1:] or list(cases):
        c = cases[name]
        torch.manual_seed(0)
        m = c["make"]().eval()
        synthetic_weights_(m, 100)
        torch.nn.init.normal_(m.pos_embs, std=0.02)
        torch.nn.init.normal_(m.token_embs.weight, std=0.05)
        vocab, d = m.token_embs.weight.shape
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        n_layers = len(m.layers)
        # ---- parity on a small slice: fp32 CPU oracle vs (a) this library, (b) torch bf16 on the GPU
        torch.manual_seed(1)
        ids = torch.randint(0, vocab, (c["parity_batch"], c["parity_L"]))
        mem = torch.randn(c["parity_batch"], 200, d) if "Lm" in c else None
        fwd = (lambda s, i, mm: oracle_torch.gpt2_forward(s, i)) if name == "gpt2" else \
              (lambda s, i, mm: oracle_torch.whisper_decoder_forward(s, i, mm))
        want = fwd(sd, ids, mem)
        mg = m.to(dev).bfloat16()
        got = mg(ids.to(dev)) if mem is None else mg(ids.to(dev), mem.to(dev).bfloat16())
        sd_g = {k: v.to(dev).bfloat16() for k, v in sd.items()}
        lib = fwd(sd_g, ids.to(dev), None if mem is None else mem.to(dev).bfloat16())
        entry = dict(config=dict(n_layers=n_layers, d_model=d, vocab=vocab, batch=c["batch"], L=c["L"], Lm=c.get("Lm")),
                     parity=dict(ours=rel_err(got.float().cpu(), want), torch_bf16=rel_err(lib.float().cpu(), want)))
        # ---- throughput at the full shape
        ids = torch.randint(0, vocab, (c["batch"], c["L"]), device=dev)
        mem = torch.randn(c["batch"], c["Lm"], d, device=dev, dtype=torch.bfloat16) if "Lm" in c else None
        run_ours = (lambda: mg(ids)) if mem is None else (lambda: mg(ids, mem))
        run_lib = lambda: fwd(sd_g, ids, mem)  # noqa: E731
        ms_ours, ms_lib = timeit(run_ours, 5), timeit(run_lib, 3)
        tok = c["batch"] * c["L"]
        L, Lm = c["L"], c.get("Lm", 0)
        per_layer = 24.0 * L * d * d + 2.0 * L * L * d  # causal self-attention does half of 4 L^2 d
        if Lm:
            per_layer += 4.0 * L * d * d + 4.0 * Lm * d * d + 4.0 * L * Lm * d  # q/out proj, k/v proj of memory, attention
        flops = c["batch"] * (n_layers * per_layer + 2.0 * L * d * vocab)
        entry["throughput"] = dict(ours_ms=ms_ours, ours_tokens_per_s=tok / ms_ours * 1e3,
                                   ours_model_tflops=flops / ms_ours * 1e-9, torch_bf16_ms=ms_lib,
                                   torch_bf16_tokens_per_s=tok / ms_lib * 1e3, speedup=ms_lib / ms_ours)
        report[name] = entry
        print(name, json.dumps(entry), flush=True)
        del m, mg, sd_g, got, lib
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "decoder_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
