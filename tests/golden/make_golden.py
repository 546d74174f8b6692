"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) in fp32 on CPU.

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Each fixture holds the seeded weights (the reference's state_dict after `randomize_`), the input, the reference's
output and the hyper-parameters needed to rebuild the model. The big C1 fixture (ViT-Ti/16, batch 8, BASELINE
configs[0]) stores only seeds + outputs: its weights are reproduced from `torch.manual_seed` because the product
modules construct their parameters in the same order as the reference (checked in tests/test_host.py).

ViT with a class token: the reference's forward only works at batch 1 (vit.py:80-81 has no .expand), so such
fixtures are produced per sample and concatenated — exactly how a reference user would have to call it.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))

from pytorch_models.audio2text.whisper import Whisper, WhisperEncoder, WhisperPreprocessor  # noqa: E402
from pytorch_models.image import ViT  # noqa: E402
from pytorch_models.text import BERT, GPT, GPT2  # noqa: E402
from pytorch_models.transformer import Decoder  # noqa: E402

from oracle.oracle_torch import randomize_  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def per_sample(m: torch.nn.Module, x: torch.Tensor) -> torch.Tensor:
    return torch.cat([m(x[i:i + 1]) for i in range(x.shape[0])])


def vit_token_outputs(m: ViT, x: torch.Tensor) -> torch.Tensor:
    """Final-norm token embeddings (before the pooler) from the reference modules, per sample."""
    outs = []
    for i in range(x.shape[0]):
        t = m.patch_embed(x[i:i + 1]).flatten(-2).transpose(-1, -2) + m.pe
        if m.cls_token is not None:
            t = torch.cat([m.cls_token, t], dim=-2)
        outs.append(m.norm(m.layers(t)))
    return torch.cat(outs)


def save(name: str, model: torch.nn.Module, hyper: dict, inputs: torch.Tensor, outputs: dict, weights: bool = True,
         extra: dict | None = None):
    arrays = {f"out.{k}": v.detach().numpy() for k, v in outputs.items()}
    arrays["input"] = inputs.numpy()
    arrays.update({f"in.{k}": v.numpy() for k, v in (extra or {}).items()})
    if weights:
        arrays.update({f"sd.{k}": v.detach().numpy() for k, v in model.state_dict().items()})
    arrays["hyper"] = np.frombuffer(json.dumps(hyper).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, outputs " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in outputs.items()))


@torch.no_grad()
def main() -> None:
    torch.set_num_threads(8)

    def vit_case(name, seed, batch, img, **kw):
        torch.manual_seed(seed)
        m = ViT(**kw, img_size=img).eval()
        randomize_(m.state_dict(), seed + 100)
        x = torch.randn(batch, 3, img, img)
        out = per_sample(m, x) if m.cls_token is not None else m(x)
        save(name, m, dict(kind="vit", img_size=img, **kw), x, dict(pooled=out, tokens=vit_token_outputs(m, x)))

    vit_case("vit_cls", 1, 3, 64, n_layers=2, d_model=128, n_heads=2, patch_size=16)
    vit_case("vit_gap", 2, 2, 64, n_layers=1, d_model=128, n_heads=2, patch_size=16, pool_type="gap")
    vit_case("vit_siglip", 3, 3, 64, n_layers=2, d_model=128, n_heads=2, patch_size=16, cls_token=False, pool_type="mha")
    vit_case("vit_p14", 4, 2, 56, n_layers=1, d_model=64, n_heads=1, patch_size=14)
    vit_case("vit_long", 5, 1, 192, n_layers=1, d_model=64, n_heads=1, patch_size=8)  # 577 tokens: streaming KV path

    torch.manual_seed(6)
    w = WhisperEncoder(2, 128, 80).eval()
    randomize_(w.state_dict(), 106)
    x = torch.randn(2, 80, 120)
    save("whisper", w, dict(kind="whisper", n_layers=2, d_model=128, n_mels=80), x, dict(tokens=w(x)))

    torch.manual_seed(7)
    b = BERT(1000, 2, 128).eval()
    randomize_(b.state_dict(), 107)
    torch.nn.init.normal_(b.pos_embs, std=0.02)
    t = torch.randint(3, 1000, (2, 16))
    save("bert", b, dict(kind="bert", vocab_size=1000, n_layers=2, d_model=128), t, dict(tokens=b(t)))

    decoder_cases()
    audio_cases()

    # C1 = BASELINE configs[0]: ViT-Ti/16 augreg 224, batch 8, random-init weights, fp32 CPU forward (reference path)
    torch.manual_seed(0)
    m = ViT.from_google("Ti/16").eval()
    randomize_(m.state_dict(), 100)
    torch.manual_seed(1)
    x = torch.randn(8, 3, 224, 224)
    save("c1_vit_ti16", m, dict(kind="vit_seeded", tag="Ti/16", weight_seed=0, noise_seed=100, input_seed=1, batch=8),
         torch.zeros(1), dict(pooled=per_sample(m, x), tokens=vit_token_outputs(m, x)), weights=False)


def decoder_cases() -> None:
    """DecoderLayer / Decoder and the models built on it (transformer.py:70-105,152-176)."""
    def noise_(m: torch.nn.Module, seed: int) -> None:
        randomize_(m.state_dict(), seed)

    # post-norm decoder with cross-attention: not reachable through any model class, so pinned directly
    torch.manual_seed(8)
    dec = Decoder(2, 128, cross_attn=True, pre_norm=False).eval()
    noise_(dec, 108)
    x, mem = torch.randn(2, 20, 128), torch.randn(2, 37, 128)
    save("decoder_postnorm_cross", dec, dict(kind="decoder", n_layers=2, d_model=128, cross_attn=True, pre_norm=False),
         x, dict(tokens=dec(x, mem)), extra=dict(memory=mem))

    # 300 tokens: the causal mask crosses 128-row tile and 128-column block boundaries
    torch.manual_seed(9)
    dec = Decoder(1, 64).eval()
    noise_(dec, 109)
    x = torch.randn(1, 300, 64)
    save("decoder_causal_long", dec, dict(kind="decoder", n_layers=1, d_model=64, cross_attn=False, pre_norm=True),
         x, dict(tokens=dec(x)))

    # full Whisper (encoder + decoder with cross-attention); odd vocabulary like the real 51865 / 51866
    torch.manual_seed(10)
    w = Whisper(1001, 2, 128, 80).eval()
    noise_(w, 110)
    torch.nn.init.normal_(w.decoder.pos_embs, std=0.02)
    torch.nn.init.normal_(w.decoder.token_embs.weight, std=0.05)
    x, tgt = torch.randn(2, 80, 120), torch.randint(0, 1001, (2, 12))
    save("whisper_full", w, dict(kind="whisper_full", vocab_size=1001, n_layers=2, d_model=128, n_mels=80), x,
         dict(logits=w(x, tgt), memory=w.encoder(x)), extra=dict(targets=tgt))

    # GPT-2 / GPT with a reduced vocabulary (vocab_size is a class attribute in the reference: gpt2.py:12, gpt.py:15)
    for name, base, seed in (("gpt2", GPT2, 11), ("gpt", GPT, 12)):
        cls = type(f"{base.__name__}Small", (base,), dict(vocab_size=777))
        torch.manual_seed(seed)
        m = cls(2, 64).eval()
        noise_(m, 100 + seed)
        torch.nn.init.normal_(m.pos_embs, std=0.02)
        torch.nn.init.normal_(m.token_embs.weight, std=0.05)
        t = torch.randint(0, 777, (2, 33))
        save(name, m, dict(kind=name, vocab_size=777, n_layers=2, d_model=64), t, dict(logits=m(t)))


def audio_cases() -> None:
    """the Whisper audio front end (whisper.py:138-148, audio/spectrogram.py)."""
    torch.manual_seed(13)
    for name, variant, shape in (("logmel_tiny", "tiny", (2, 16000)), ("logmel_large_v3", "large-v3", (1, 3333))):
        pre = WhisperPreprocessor(variant).eval()
        t = torch.arange(shape[1]) / 16000.0
        x = 0.1 * torch.randn(shape) + 0.5 * torch.sin(2 * torch.pi * 440.0 * t) * torch.linspace(0, 1, shape[1])
        x[0, : shape[1] // 5] = 0.0  # a stretch of digital silence: log10(0) = -inf meets the max - 8 floor
        save(name, pre, dict(kind="logmel", variant=variant), x, dict(logmel=pre(x)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "audio":
        with torch.no_grad():
            torch.set_num_threads(8)
            audio_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "decoder":  # add the decoder fixtures without rewriting the others
        with torch.no_grad():
            torch.set_num_threads(8)
            decoder_cases()
    else:
        main()
