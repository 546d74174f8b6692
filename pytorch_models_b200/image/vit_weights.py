"""Pretrained-weight converters for ``ViT`` (host-side, run once at load time; reference: image/vit.py:122-200 for the
Google flax ``.npz`` checkpoints, image/vit.py:241-306 for the Facebook ``.pth`` state dicts).

Parameters are written IN PLACE into the already-constructed modules, like the reference does, so the packed-weight
caches of the kernels notice (``Tensor._version``) and rebuild on the next forward. Downloads need network access.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch import Tensor, nn

_AUGREG = {
    "Ti/16": "Ti_16-i21k-300ep-lr_0.001-aug_none-wd_0.03-do_0.0-sd_0.0.npz",
    "S/32": "S_32-i21k-300ep-lr_0.001-aug_none-wd_0.1-do_0.0-sd_0.0.npz",
    "S/16": "S_16-i21k-300ep-lr_0.001-aug_light1-wd_0.03-do_0.0-sd_0.0.npz",
    "B/32": "B_32-i21k-300ep-lr_0.001-aug_light1-wd_0.1-do_0.0-sd_0.0.npz",
    "B/16": "B_16-i21k-300ep-lr_0.001-aug_medium1-wd_0.1-do_0.0-sd_0.0.npz",
    "L/16": "L_16-i21k-300ep-lr_0.001-aug_strong1-wd_0.1-do_0.0-sd_0.0.npz",
}
_SIGLIP = {
    ("B/16", 224): "webli_en_b16_224_63724782.npz",
    ("B/16", 256): "webli_en_b16_256_60500360.npz",
    ("B/16", 384): "webli_en_b16_384_68578854.npz",
    ("B/16", 512): "webli_en_b16_512_68580893.npz",
    ("L/16", 256): "webli_en_l16_256_60552751.npz",
    ("L/16", 384): "webli_en_l16_384_63634585.npz",
}


def hub_download(url: str, subdir: str | None = None) -> str:
    """Cached download into ``torch.hub.get_dir()`` (reference: pytorch_models/utils.py:6-16)."""
    import requests

    parts = [torch.hub.get_dir()] + ([subdir] if subdir else []) + [os.path.basename(url)]
    path = os.path.join(*parts)
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        resp = requests.get(url)
        resp.raise_for_status()
        with open(path, "wb") as f:
            f.write(resp.content)
    return path


# ------------------------------------------------------------------------------------------------ Google (flax .npz)
def _ln(norm: nn.LayerNorm, w: dict, key: str) -> None:
    norm.weight.copy_(w.pop(f"{key}/scale"))
    norm.bias.copy_(w.pop(f"{key}/bias"))


def _dense(lin: nn.Linear, w: dict, key: str) -> None:
    """flax kernels are (in..., out...) — flatten to (in, out) and transpose into nn.Linear's (out, in)."""
    n_out, n_in = lin.weight.shape
    lin.weight.copy_(w.pop(f"{key}/kernel").reshape(n_in, n_out).T)
    if lin.bias is not None:
        lin.bias.copy_(w.pop(f"{key}/bias").reshape(-1))


def _attention(mha, w: dict, key: str) -> None:
    for ours, theirs in (("q_proj", "query"), ("k_proj", "key"), ("v_proj", "value"), ("out_proj", "out")):
        _dense(getattr(mha, ours), w, f"{key}/{theirs}")


@torch.no_grad()
def load_flax_arrays(model, arrays: dict, *, big_vision: bool = False) -> dict:
    """Copy a flax parameter tree (flat ``{"a/b/c": array}``) into ``model``; returns the entries left unused."""
    w = {k: torch.as_tensor(np.asarray(v)) for k, v in arrays.items()}
    # module numbering differs between google-research/vision_transformer and google-research/big_vision
    attn, ln2, mlp = (("MultiHeadDotProductAttention_0", "LayerNorm_1", "MlpBlock_0") if big_vision
                      else ("MultiHeadDotProductAttention_1", "LayerNorm_2", "MlpBlock_3"))
    if model.cls_token is not None:
        model.cls_token.copy_(w.pop("cls"))
    if big_vision:
        model.pe.copy_(w.pop("pos_embedding"))
    else:
        pos = w.pop("Transformer/posembed_input/pos_embedding")
        model.cls_token.add_(pos[:, 0])  # the class token's position embedding is folded into the token itself
        model.pe.copy_(pos[:, 1:])
    model.patch_embed.weight.copy_(w.pop("embedding/kernel").permute(3, 2, 0, 1))  # HWIO -> OIHW
    if model.patch_embed.bias is not None:
        model.patch_embed.bias.copy_(w.pop("embedding/bias"))
    _ln(model.norm, w, "Transformer/encoder_norm")
    for i, layer in enumerate(model.layers):
        block = f"Transformer/encoderblock_{i}"
        _ln(layer.sa_norm, w, f"{block}/LayerNorm_0")
        _attention(layer.sa, w, f"{block}/{attn}")
        _ln(layer.mlp_norm, w, f"{block}/{ln2}")
        _dense(layer.mlp.linear1, w, f"{block}/{mlp}/Dense_0")
        _dense(layer.mlp.linear2, w, f"{block}/{mlp}/Dense_1")
    pooler = model.pooler
    if hasattr(pooler, "probe"):  # SigLIP MAP head
        pooler.probe.copy_(w.pop("MAPHead_0/probe"))
        _attention(pooler.attn, w, "MAPHead_0/MultiHeadDotProductAttention_0")
        _ln(pooler.norm, w, "MAPHead_0/LayerNorm_0")
        _dense(pooler.mlp.linear1, w, "MAPHead_0/MlpBlock_0/Dense_0")
        _dense(pooler.mlp.linear2, w, "MAPHead_0/MlpBlock_0/Dense_1")
    return w


def load_google(model, tag: str, weights: str, img_size: int = 224) -> None:
    if weights == "augreg":
        path = hub_download(f"https://storage.googleapis.com/vit_models/augreg/{_AUGREG[tag]}")
        prefix, big_vision = "", False
    elif weights == "siglip":
        path = hub_download(f"https://storage.googleapis.com/big_vision/siglip/{_SIGLIP[(tag, img_size)]}")
        prefix, big_vision = "params/img/", True
    else:
        raise ValueError(f"Unsupported weights={weights}")
    arrays = {k[len(prefix):]: v for k, v in np.load(path).items() if k.startswith(prefix)}
    left = load_flax_arrays(model, arrays, big_vision=big_vision)
    if left:
        print(left.keys())


# ------------------------------------------------------------------------------------------------ Facebook (.pth)
@torch.no_grad()
def load_facebook_state_dict(model, state_dict: dict[str, Tensor]) -> dict:
    """DeiT-3 / DINO / DINOv2 checkpoints (timm-style keys). LayerScale (``gamma_1/2`` or ``ls1/2.gamma``) is folded
    into ``out_proj`` / ``linear2`` in place, as the reference does (vit.py:290-304). Returns unused entries."""
    sd = dict(state_dict)

    def take(module, key: str) -> None:
        module.weight.copy_(sd.pop(f"{key}.weight").view(module.weight.shape))
        module.bias.copy_(sd.pop(f"{key}.bias"))

    def fold_layerscale(lin: nn.Linear, *names: str) -> None:
        for name in names:
            gamma = sd.pop(name, None)
            if gamma is not None:
                lin.weight.mul_(gamma.view(-1, 1))
                lin.bias.mul_(gamma)
                return

    take(model.patch_embed, "patch_embed.proj")
    pos = sd.pop("pos_embed")
    n = model.pe.shape[1]
    model.pe.copy_(pos[:, -n:])
    model.cls_token.copy_(sd.pop("cls_token"))
    if pos.shape[1] > n:  # checkpoints that carry a position embedding for the class token
        model.cls_token.add_(pos[:, 0])
    take(model.norm, "norm")
    d = model.norm.normalized_shape[0]
    for i, layer in enumerate(model.layers):
        blk = f"blocks.{i}"
        take(layer.sa_norm, f"{blk}.norm1")
        take(layer.mlp_norm, f"{blk}.norm2")
        qkv_w, qkv_b = sd.pop(f"{blk}.attn.qkv.weight"), sd.pop(f"{blk}.attn.qkv.bias")
        for j, proj in enumerate((layer.sa.q_proj, layer.sa.k_proj, layer.sa.v_proj)):
            proj.weight.copy_(qkv_w[j * d:(j + 1) * d])
            proj.bias.copy_(qkv_b[j * d:(j + 1) * d])
        take(layer.sa.out_proj, f"{blk}.attn.proj")
        fold_layerscale(layer.sa.out_proj, f"{blk}.gamma_1", f"{blk}.ls1.gamma")
        take(layer.mlp.linear1, f"{blk}.mlp.fc1")
        take(layer.mlp.linear2, f"{blk}.mlp.fc2")
        fold_layerscale(layer.mlp.linear2, f"{blk}.gamma_2", f"{blk}.ls2.gamma")
    return sd


def load_facebook(model, size: str, patch: int, weights: str, img_size: int) -> None:
    if weights == "deit3":
        assert patch == 16
        name = dict(S="small", M="medium", B="base", L="large", H="huge")[size]
        url = f"https://dl.fbaipublicfiles.com/deit/deit_3_{name}_{img_size}_21k.pth"
    elif weights == "dino":
        assert patch in (16, 8)
        tag = f"dino_deit{dict(S='small', B='base')[size]}{patch}_pretrain"
        url = f"https://dl.fbaipublicfiles.com/dino/{tag}/{tag}.pth"
    elif weights == "dinov2":
        assert patch == 14
        tag = f"dinov2_vit{size.lower()}{patch}"
        url = f"https://dl.fbaipublicfiles.com/dinov2/{tag}/{tag}_pretrain.pth"
    else:
        raise ValueError(f"Unsupported {weights}")
    sd = torch.hub.load_state_dict_from_url(url)
    left = load_facebook_state_dict(model, sd.get("model", sd))
    if left:
        print(left.keys())
