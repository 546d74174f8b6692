"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product package (pytorch_models_b200/).

Plain numpy (fp32, scipy.special.erf for the exact GELU) restatement of the reference's encoder hot path, one
function per reference call site. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may use it, and only as the checker / the CPU arm — never as a fallback of the CUDA path.

Where the arithmetic really lives: the reference delegates every op to PyTorch (un-pinned dependency; CI uses
torch==2.1.2, this image 2.11.0), so the algorithm restated here is the published definition of those ops
(nn.LayerNorm, nn.Linear, nn.GELU, F.scaled_dot_product_attention, nn.Conv2d / nn.Conv1d with stride == the
reference's), anchored on the reference's own call sites.

Parity status: PINNED. The reference has no golden vectors of its own (its parity tests need network + timm /
openai-whisper / HF), so this oracle is pinned against outputs of the *reference itself* run in the build container:
``tests/golden/make_golden.py`` imports /root/reference, runs its modules in fp32 on seeded weights/inputs and
stores weights + inputs + outputs under ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks both oracles against
every fixture (max-abs 2e-5, the tolerance the reference holds itself to: tests/image/test_vit.py:45).

Models are described by a ``state_dict`` (name -> fp32 array with the reference's keys) plus a few hyper-parameters.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf

F32 = np.float32


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float) -> np.ndarray:
    """nn.LayerNorm over the last dim, biased variance (transformer.py:87,93; vit.py:69; whisper.py:27; bert.py:31)."""
    mu = x.mean(-1, keepdims=True, dtype=F32)
    var = ((x - mu) ** 2).mean(-1, keepdims=True, dtype=F32)
    return ((x - mu) / np.sqrt(var + F32(eps)) * w + b).astype(F32)


def linear(x: np.ndarray, w: np.ndarray, b: np.ndarray | None) -> np.ndarray:
    """nn.Linear: x @ w.T + b (transformer.py:28-31,59,66)."""
    y = x @ w.T
    return (y + b).astype(F32) if b is not None else y.astype(F32)


def gelu(x: np.ndarray) -> np.ndarray:
    """nn.GELU() exact erf form (transformer.py:61)."""
    return (0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))).astype(F32)


def gelu_tanh(x: np.ndarray) -> np.ndarray:
    """nn.GELU(approximate="tanh") (transformer.py:62)."""
    return (0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))).astype(F32)


def sdpa(q: np.ndarray, k: np.ndarray, v: np.ndarray, causal: bool = False) -> np.ndarray:
    """F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0, is_causal=causal) (transformer.py:52):
    softmax(q k^T / sqrt(head_dim)) v over (..., heads, L, head_dim); with is_causal, query i sees keys j <= i
    (lower-triangular mask aligned at the top-left corner)."""
    s = (q @ np.swapaxes(k, -1, -2)) * F32(1.0 / np.sqrt(q.shape[-1]))
    if causal:
        Lq, Lk = s.shape[-2:]
        s = np.where(np.tril(np.ones((Lq, Lk), dtype=bool)), s, F32(-np.inf))
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(-1, keepdims=True, dtype=F32)
    return (p @ v).astype(F32)


def mha(sd: dict, prefix: str, q_in: np.ndarray, kv_in: np.ndarray | None, n_heads: int,
        causal: bool = False) -> np.ndarray:
    """MHA.forward (transformer.py:36-53) with k = v = kv_in (or q_in when kv_in is None), attn_bias None."""
    kv_in = q_in if kv_in is None else kv_in

    def heads(t: np.ndarray) -> np.ndarray:  # (*, L, h*hd) -> (*, h, L, hd)   transformer.py:47-49
        return np.swapaxes(t.reshape(*t.shape[:-1], n_heads, t.shape[-1] // n_heads), -2, -3)

    q = heads(linear(q_in, sd[prefix + "q_proj.weight"], sd.get(prefix + "q_proj.bias")))
    k = heads(linear(kv_in, sd[prefix + "k_proj.weight"], sd.get(prefix + "k_proj.bias")))
    v = heads(linear(kv_in, sd[prefix + "v_proj.weight"], sd.get(prefix + "v_proj.bias")))
    o = sdpa(q, k, v, causal)
    o = np.swapaxes(o, -2, -3)
    o = o.reshape(*o.shape[:-2], -1)  # transformer.py:53
    return linear(o, sd[prefix + "out_proj.weight"], sd.get(prefix + "out_proj.bias"))


def mlp(sd: dict, prefix: str, x: np.ndarray, act: str = "gelu") -> np.ndarray:
    """MLP: linear1 -> GELU -> linear2 -> dropout(eval = identity) (transformer.py:56-67)."""
    fn = gelu_tanh if act == "approximate_gelu" else gelu
    h = fn(linear(x, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"]))
    return linear(h, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])


def encoder_layer(sd: dict, prefix: str, x: np.ndarray, n_heads: int, pre_norm: bool, eps: float) -> np.ndarray:
    """EncoderLayer.forward (transformer.py:123-130)."""
    def ln(name: str, t: np.ndarray) -> np.ndarray:
        return layer_norm(t, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"], eps)

    if pre_norm:  # transformer.py:125-126
        x = x + mha(sd, prefix + "sa.", ln("sa_norm", x), None, n_heads)
        x = x + mlp(sd, prefix + "mlp.", ln("mlp_norm", x))
    else:  # transformer.py:128-129
        x = ln("sa_norm", x + mha(sd, prefix + "sa.", x, None, n_heads))
        x = ln("mlp_norm", x + mlp(sd, prefix + "mlp.", x))
    return x.astype(F32)


def decoder_layer(sd: dict, prefix: str, x: np.ndarray, memory: np.ndarray | None, n_heads: int, pre_norm: bool,
                  eps: float, act: str = "gelu") -> np.ndarray:
    """DecoderLayer.forward (transformer.py:95-105); cross-attention iff the state dict holds ``ca.*``."""
    def ln(name: str, t: np.ndarray) -> np.ndarray:
        return layer_norm(t, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"], eps)

    cross = prefix + "ca.q_proj.weight" in sd
    if pre_norm:  # transformer.py:97-99
        x = x + mha(sd, prefix + "sa.", ln("sa_norm", x), None, n_heads, causal=True)
        if cross:
            x = x + mha(sd, prefix + "ca.", ln("ca_norm", x), memory, n_heads)
        x = x + mlp(sd, prefix + "mlp.", ln("mlp_norm", x), act)
    else:  # transformer.py:101-103
        x = ln("sa_norm", x + mha(sd, prefix + "sa.", x, None, n_heads, causal=True))
        if cross:
            x = ln("ca_norm", x + mha(sd, prefix + "ca.", x, memory, n_heads))
        x = ln("mlp_norm", x + mlp(sd, prefix + "mlp.", x, act))
    return x.astype(F32)


def decoder(sd: dict, x: np.ndarray, memory: np.ndarray | None, n_heads: int, pre_norm: bool, eps: float,
            act: str = "gelu", prefix: str = "layers.") -> np.ndarray:
    """Decoder.forward (transformer.py:173-176)."""
    for i in range(n_layers_of(sd, prefix)):
        x = decoder_layer(sd, f"{prefix}{i}.", x, memory, n_heads, pre_norm, eps, act)
    return x


def n_layers_of(sd: dict, prefix: str = "layers.") -> int:
    return 1 + max(int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix))


def encoder(sd: dict, x: np.ndarray, n_heads: int, pre_norm: bool, eps: float, prefix: str = "layers.") -> np.ndarray:
    """Encoder = nn.Sequential of EncoderLayer (transformer.py:133-149)."""
    for i in range(n_layers_of(sd, prefix)):
        x = encoder_layer(sd, f"{prefix}{i}.", x, n_heads, pre_norm, eps)
    return x


def patch_embed(imgs: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """nn.Conv2d(3, d, p, stride=p) then flatten(-2).transpose(-1,-2) (vit.py:64,78): (N,3,H,W) -> (N, P, d)."""
    d, c, p, _ = w.shape
    n, _, hh, ww = imgs.shape
    x = imgs.reshape(n, c, hh // p, p, ww // p, p).transpose(0, 2, 4, 1, 3, 5).reshape(n, (hh // p) * (ww // p), c * p * p)
    return (x @ w.reshape(d, -1).T + b).astype(F32)


def vit_tokens(sd: dict, imgs: np.ndarray) -> np.ndarray:
    """vit.py:78-81: patch embedding + pe, class token (no pe on it) prepended for EVERY image. The reference omits
    the .expand and therefore only runs at batch 1; expanding is bit-identical to per-sample calls."""
    x = patch_embed(imgs, sd["patch_embed.weight"], sd["patch_embed.bias"]) + sd["pe"]
    if "cls_token" in sd:
        cls = np.broadcast_to(sd["cls_token"], (x.shape[0], 1, x.shape[2]))
        x = np.concatenate([cls, x], axis=1)
    return x.astype(F32)


def vit_forward(sd: dict, imgs: np.ndarray, n_heads: int, pool: str = "cls_token", eps: float = 1e-6,
                return_tokens: bool = False) -> np.ndarray:
    """ViT.forward (vit.py:77-85); pool in {"cls_token", "gap", "mha"} (vit.py:20-43)."""
    x = encoder(sd, vit_tokens(sd, imgs), n_heads, True, eps)
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], eps)
    if return_tokens:
        return x
    if pool == "cls_token":
        return x[:, 0]
    if pool == "gap":
        return x.mean(1, dtype=F32)
    # MHAPooling.forward (vit.py:40-43)
    y = mha(sd, "pooler.attn.", sd["pooler.probe"], x, n_heads)[:, 0]
    return (y + mlp(sd, "pooler.mlp.", layer_norm(y, sd["pooler.norm.weight"], sd["pooler.norm.bias"], eps))).astype(F32)


def conv1d_k3(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int) -> np.ndarray:
    """nn.Conv1d(C, d, 3, stride, padding=1) (whisper.py:17,19): (N, C, T) -> (N, d, T_out)."""
    n, c, t = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1)))
    t_out = (t + 2 - 3) // stride + 1
    cols = np.stack([xp[:, :, k:k + stride * (t_out - 1) + 1:stride] for k in range(3)], axis=-1)  # (N, C, T_out, 3)
    return (np.einsum("nctk,dck->ndt", cols, w) + b[None, :, None]).astype(F32)


def whisper_encoder_forward(sd: dict, x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """WhisperEncoder.forward (whisper.py:29-34); Encoder defaults: head_dim 64, pre-norm, eps 1e-5 (whisper.py:26)."""
    h = gelu(conv1d_k3(x, sd["stem.0.weight"], sd["stem.0.bias"], 1))
    h = gelu(conv1d_k3(h, sd["stem.2.weight"], sd["stem.2.bias"], 2))
    h = np.swapaxes(h, 1, 2)
    h = h + sd["pos_embs"][: h.shape[1]]
    d = h.shape[-1]
    h = encoder(sd, h.astype(F32), d // 64, True, eps)
    return layer_norm(h, sd["norm.weight"], sd["norm.bias"], eps)


def bert_forward(sd: dict, tokens: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """BERT.forward (bert.py:34-39): embedding + positions, embedding LayerNorm, post-norm encoder."""
    x = sd["token_embs.weight"][tokens] + sd["pos_embs"][: tokens.shape[-1]]
    x = layer_norm(x.astype(F32), sd["norm.weight"], sd["norm.bias"], eps)
    return encoder(sd, x, x.shape[-1] // 64, False, eps)


def sub_dict(sd: dict, prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def lm_embed(sd: dict, tokens: np.ndarray) -> np.ndarray:
    """token_embs(x) + pos_embs[:L] (gpt2.py:22-23, gpt.py:25-26, whisper.py:47-48)."""
    return (sd["token_embs.weight"][tokens] + sd["pos_embs"][: tokens.shape[-1]]).astype(F32)


def whisper_decoder_forward(sd: dict, tokens: np.ndarray, memory: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """WhisperDecoder.forward (whisper.py:46-52)."""
    x = lm_embed(sd, tokens)
    x = decoder(sd, x, memory, x.shape[-1] // 64, True, eps)
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], eps)
    return (x @ sd["token_embs.weight"].T).astype(F32)


def whisper_forward(sd: dict, x: np.ndarray, targets: np.ndarray) -> np.ndarray:
    """Whisper.forward (whisper.py:62-63)."""
    memory = whisper_encoder_forward(sub_dict(sd, "encoder."), x)
    return whisper_decoder_forward(sub_dict(sd, "decoder."), targets, memory)


def gpt2_forward(sd: dict, tokens: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """GPT2.forward (gpt2.py:21-27)."""
    x = lm_embed(sd, tokens)
    x = decoder(sd, x, None, x.shape[-1] // 64, True, eps, "approximate_gelu")
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], eps)
    return (x @ sd["token_embs.weight"].T).astype(F32)


def gpt_forward(sd: dict, tokens: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """GPT.forward (gpt.py:24-29)."""
    x = lm_embed(sd, tokens)
    x = decoder(sd, x, None, x.shape[-1] // 64, False, eps, "approximate_gelu")
    return (x @ sd["token_embs.weight"].T).astype(F32)


def whisper_logmel(audio: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """WhisperPreprocessor.forward (whisper.py:143-148): torch.stft(x, 400, 160, hann, center=True, reflect)
    (audio/spectrogram.py:15-16) restated as explicit framing + rfft in float64, mel projection (:44-45), drop of the
    last frame, log10, per-sample floor at max - 8 and the (x + 4) / 4 rescale."""
    n_fft, hop = 400, 160
    x = np.asarray(audio, dtype=np.float64)
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(n_fft // 2, n_fft // 2)], mode="reflect")
    n_frames = 1 + x.shape[-1] // hop
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)  # torch.hann_window(400): periodic
    power = np.abs(np.fft.rfft(xp[..., idx] * window, axis=-1)) ** 2       # (..., frames, 201)
    mel = np.swapaxes(power @ np.asarray(filters, dtype=np.float64).T, -1, -2)[..., :-1]
    with np.errstate(divide="ignore"):
        logmel = np.log10(np.maximum(mel, 0.0))
    floor = logmel.reshape(*logmel.shape[:-2], -1).max(-1)[..., None, None] - 8.0
    return ((np.maximum(logmel, floor) + 4.0) / 4.0).astype(F32)
