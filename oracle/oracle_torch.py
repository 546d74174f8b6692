"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product package (pytorch_models_b200/).

The same restatement as ``oracle_np.py`` but written on ``torch.nn.functional`` CPU ops in fp32 — i.e. on the very ATen
operators the reference dispatches to (native_layer_norm, addmm, scaled_dot_product_attention, gelu,
convolution), so it is bit-identical to the reference modules and multi-threaded. Used (a) as the expected value of
the larger GPU parity tests, (b) as the CPU arm of ``bench.py`` (``cpu_baseline`` and ``--impl reference``), where
it stands in for the reference's own CPU forward (kind = "port"; /root/reference does not exist on the GPU box).

Parity status: PINNED against the reference itself through ``tests/golden/*.npz`` (see oracle_np.py header).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm (transformer.py:87,93; vit.py:69; whisper.py:27; bert.py:31)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def mha(sd: dict, prefix: str, q_in: Tensor, kv_in: Tensor | None, n_heads: int, causal: bool = False,
        attn_bias: Tensor | None = None) -> Tensor:
    """MHA.forward (transformer.py:36-53), k = v = kv_in (self-attention when None)."""
    kv_in = q_in if kv_in is None else kv_in

    def proj(name: str, t: Tensor) -> Tensor:
        y = F.linear(t, sd[f"{prefix}{name}.weight"], sd.get(f"{prefix}{name}.bias"))
        return y.unflatten(-1, (n_heads, -1)).transpose(-2, -3)

    o = F.scaled_dot_product_attention(proj("q_proj", q_in), proj("k_proj", kv_in), proj("v_proj", kv_in),
                                       attn_mask=attn_bias, is_causal=causal)
    return F.linear(o.transpose(-2, -3).flatten(-2), sd[prefix + "out_proj.weight"], sd.get(prefix + "out_proj.bias"))


def mlp(sd: dict, prefix: str, x: Tensor, act: str = "gelu") -> Tensor:
    """MLP (transformer.py:56-67): linear1 -> GELU (exact, or tanh form for act="approximate_gelu", :61-62) -> linear2."""
    h = F.gelu(F.linear(x, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"]),
               approximate="tanh" if act == "approximate_gelu" else "none")
    return F.linear(h, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])


def encoder_layer(sd: dict, prefix: str, x: Tensor, n_heads: int, pre_norm: bool, eps: float) -> Tensor:
    """EncoderLayer.forward (transformer.py:123-130)."""
    def ln(name: str, t: Tensor) -> Tensor:
        return layer_norm(t, sd[f"{prefix}{name}.weight"], sd[f"{prefix}{name}.bias"], eps)

    if pre_norm:
        x = x + mha(sd, prefix + "sa.", ln("sa_norm", x), None, n_heads)
        return x + mlp(sd, prefix + "mlp.", ln("mlp_norm", x))
    x = ln("sa_norm", x + mha(sd, prefix + "sa.", x, None, n_heads))
    return ln("mlp_norm", x + mlp(sd, prefix + "mlp.", x))


def decoder_layer(sd: dict, prefix: str, x: Tensor, memory: Tensor | None, n_heads: int, pre_norm: bool, eps: float,
                  act: str = "gelu") -> Tensor:
    """DecoderLayer.forward (transformer.py:95-105); cross-attention iff the state dict holds ``ca.*``."""
    def ln(name: str, t: Tensor) -> Tensor:
        return layer_norm(t, sd[f"{prefix}{name}.weight"], sd[f"{prefix}{name}.bias"], eps)

    cross = f"{prefix}ca.q_proj.weight" in sd
    if pre_norm:
        x = x + mha(sd, prefix + "sa.", ln("sa_norm", x), None, n_heads, causal=True)
        if cross:
            x = x + mha(sd, prefix + "ca.", ln("ca_norm", x), memory, n_heads)
        return x + mlp(sd, prefix + "mlp.", ln("mlp_norm", x), act)
    x = ln("sa_norm", x + mha(sd, prefix + "sa.", x, None, n_heads, causal=True))
    if cross:
        x = ln("ca_norm", x + mha(sd, prefix + "ca.", x, memory, n_heads))
    return ln("mlp_norm", x + mlp(sd, prefix + "mlp.", x, act))


def decoder(sd: dict, x: Tensor, memory: Tensor | None, n_heads: int, pre_norm: bool, eps: float,
            act: str = "gelu", prefix: str = "layers.") -> Tensor:
    """Decoder.forward (transformer.py:173-176)."""
    for i in range(n_layers_of(sd, prefix)):
        x = decoder_layer(sd, f"{prefix}{i}.", x, memory, n_heads, pre_norm, eps, act)
    return x


def n_layers_of(sd: dict, prefix: str = "layers.") -> int:
    return 1 + max(int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix))


def encoder(sd: dict, x: Tensor, n_heads: int, pre_norm: bool, eps: float, prefix: str = "layers.",
            taps: list | None = None) -> Tensor:
    """Encoder (transformer.py:133-149); ``taps`` collects the residual stream after every layer."""
    for i in range(n_layers_of(sd, prefix)):
        x = encoder_layer(sd, f"{prefix}{i}.", x, n_heads, pre_norm, eps)
        if taps is not None:
            taps.append(x)
    return x


def vit_tokens(sd: dict, imgs: Tensor) -> Tensor:
    """vit.py:78-81 with the class token expanded over the batch (the reference itself only runs batch 1 here)."""
    p = sd["patch_embed.weight"].shape[-1]
    x = F.conv2d(imgs, sd["patch_embed.weight"], sd["patch_embed.bias"], stride=p).flatten(-2).transpose(-1, -2)
    x = x + sd["pe"]
    if "cls_token" in sd:
        x = torch.cat([sd["cls_token"].expand(x.shape[0], -1, -1), x], dim=-2)
    return x


def vit_forward(sd: dict, imgs: Tensor, n_heads: int, pool: str = "cls_token", eps: float = 1e-6,
                return_tokens: bool = False, taps: list | None = None) -> Tensor:
    """ViT.forward (vit.py:77-85) + poolers (vit.py:20-43)."""
    x = encoder(sd, vit_tokens(sd, imgs), n_heads, True, eps, taps=taps)
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], eps)
    if return_tokens:
        return x
    if pool == "cls_token":
        return x[:, 0]
    if pool == "gap":
        return x.mean(1)
    y = mha(sd, "pooler.attn.", sd["pooler.probe"], x, n_heads).squeeze(1)
    return y + mlp(sd, "pooler.mlp.", layer_norm(y, sd["pooler.norm.weight"], sd["pooler.norm.bias"], eps))


def whisper_encoder_forward(sd: dict, x: Tensor, eps: float = 1e-5) -> Tensor:
    """WhisperEncoder.forward (whisper.py:29-34)."""
    h = F.gelu(F.conv1d(x, sd["stem.0.weight"], sd["stem.0.bias"], stride=1, padding=1))
    h = F.gelu(F.conv1d(h, sd["stem.2.weight"], sd["stem.2.bias"], stride=2, padding=1))
    h = h.transpose(1, 2)
    h = h + sd["pos_embs"][: h.shape[1]]
    h = encoder(sd, h, h.shape[-1] // 64, True, eps)
    return layer_norm(h, sd["norm.weight"], sd["norm.bias"], eps)


def bert_forward(sd: dict, tokens: Tensor, eps: float = 1e-12) -> Tensor:
    """BERT.forward (bert.py:34-39)."""
    x = F.embedding(tokens, sd["token_embs.weight"]) + sd["pos_embs"][: tokens.shape[-1]]
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], eps)
    return encoder(sd, x, x.shape[-1] // 64, False, eps)


def sub_dict(sd: dict, prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def whisper_decoder_forward(sd: dict, tokens: Tensor, memory: Tensor, eps: float = 1e-5) -> Tensor:
    """WhisperDecoder.forward (whisper.py:46-52): embeddings, Decoder(cross_attn=True), LayerNorm, tied logits."""
    x = F.embedding(tokens, sd["token_embs.weight"]) + sd["pos_embs"][: tokens.shape[1]]
    x = decoder(sd, x, memory, x.shape[-1] // 64, True, eps)
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], eps)
    return x @ sd["token_embs.weight"].T


def whisper_forward(sd: dict, x: Tensor, targets: Tensor) -> Tensor:
    """Whisper.forward (whisper.py:62-63)."""
    memory = whisper_encoder_forward(sub_dict(sd, "encoder."), x)
    return whisper_decoder_forward(sub_dict(sd, "decoder."), targets, memory)


def gpt2_forward(sd: dict, tokens: Tensor, eps: float = 1e-5) -> Tensor:
    """GPT2.forward (gpt2.py:21-27): pre-norm causal Decoder with tanh-GELU, final LayerNorm, tied logits."""
    x = F.embedding(tokens, sd["token_embs.weight"]) + sd["pos_embs"][: tokens.shape[-1]]
    x = decoder(sd, x, None, x.shape[-1] // 64, True, eps, "approximate_gelu")
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], eps)
    return x @ sd["token_embs.weight"].T


def gpt_forward(sd: dict, tokens: Tensor, eps: float = 1e-5) -> Tensor:
    """GPT.forward (gpt.py:24-29): post-norm causal Decoder with tanh-GELU, no final norm, tied logits."""
    x = F.embedding(tokens, sd["token_embs.weight"]) + sd["pos_embs"][: tokens.shape[-1]]
    x = decoder(sd, x, None, x.shape[-1] // 64, False, eps, "approximate_gelu")
    return x @ sd["token_embs.weight"].T


def whisper_logmel(audio: Tensor, filters: Tensor) -> Tensor:
    """WhisperPreprocessor.forward (whisper.py:143-148) on top of MelSpectrogram / Spectrogram
    (audio/spectrogram.py:15-16,44-45); ``filters`` is the module's (n_mels, 201) buffer."""
    spec = torch.stft(audio, 400, 160, window=torch.hann_window(400), return_complex=True).abs().square()
    x = (filters @ spec)[..., :-1]
    x = x.clamp(0).log10()
    x = x.maximum(x.flatten(-2).max(-1, keepdim=True)[0].unsqueeze(-1) - 8)
    return (x + 4) / 4


def randomize_(sd: dict, seed: int) -> dict:
    """Seeded noise into the tensors the reference zero/one-initialises (cls_token, pe, pos_embs, probe, LayerNorm
    affine: vit.py:65-66, whisper.py:24), so those code paths are exercised (SURVEY §8(d) recipe)."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if not v.is_floating_point():
            continue
        last = k.split(".")[-1]
        is_norm = "norm" in k.split(".")[-2] if "." in k else False
        if k in ("cls_token", "pe", "pos_embs", "pooler.probe") or k.endswith(".pos_embs"):
            v.copy_(0.02 * torch.randn(v.shape, generator=g))
        elif is_norm and last == "weight":
            v.copy_(1.0 + 0.1 * torch.randn(v.shape, generator=g))
        elif is_norm and last == "bias":
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
    return sd
