"""B200-native drop-in for ``pytorch_models/image/vit.py``: same ``ViT`` constructor, ``from_google`` /
``from_facebook`` tags, ``resize_pe`` and ``state_dict`` layout; ``forward`` runs on libb200enc kernels.

Forward (reference vit.py:77-85):
    patch_rows      images NCHW -> [N*P, Kpad] bf16                         (the im2col view of Conv2d, vit.py:64,78)
    linear          rows x patch_embed.weight.view(d, 3p²) + bias + pe      (vit.py:78-79), written at token offset 1
    cls_rows        token 0 of every image = cls_token                      (vit.py:80-81, batch > 1 handled)
    Encoder         n_layers fused blocks                                   (vit.py:82)
    layernorm+pool  final norm only on the rows the pooler consumes         (vit.py:83-84)
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .. import ops, plans
from ..compile import compilable, compilable_module, float_like
from ..transformer import partial_stats_of, MHA, MLP, Encoder, _Packed, norm_vectors, pack_folded

_SIZES = dict(Ti=(12, 192, 3), S=(12, 384, 6), M=(12, 512, 8), B=(12, 768, 12), L=(24, 1024, 16), H=(32, 1280, 16))


class ClassTokenPooling(nn.Module):
    """x[:, 0] (vit.py:20-22). Inside ``ViT.forward`` the final norm is applied to these rows only."""

    def forward(self, x: Tensor) -> Tensor:
        return x[:, 0]


class GlobalAveragePooling(nn.Module):
    """x.mean(1) (vit.py:25-27) as one reduction kernel."""

    def forward(self, x: Tensor) -> Tensor:
        if x.dtype != torch.bfloat16 or not x.is_cuda:
            raise RuntimeError("GlobalAveragePooling expects a CUDA bfloat16 (N, L, d) tensor")
        out = torch.empty(x.shape[0], x.shape[2], device=x.device, dtype=torch.bfloat16)
        return ops.mean_tokens(x, out)


class MHAPooling(nn.Module):
    """SigLIP MAP head (vit.py:30-43): one learned query attends over the tokens, then a residual MLP block."""

    def __init__(
        self, d_model: int, n_heads: int, bias: bool = True, mlp_ratio: float = 4.0, norm_eps: float = 1e-6
    ) -> None:
        super().__init__()
        self.probe = nn.Parameter(torch.zeros(1, 1, d_model))
        self.attn = MHA(d_model, n_heads=n_heads, bias=bias)
        self.norm = nn.LayerNorm(d_model, norm_eps)
        self.mlp = MLP(d_model, int(d_model * mlp_ratio))

    def forward(self, x: Tensor) -> Tensor:
        if x.dtype != torch.bfloat16 or not x.is_contiguous():
            x = x.to(torch.bfloat16).contiguous()
        probe = self.probe.detach().to(torch.bfloat16).contiguous()  # parameter-derived (plan-safe)
        pooled = self.attn.run(probe, x).squeeze(1)  # (N, d) bf16: library launches only, no dtype round trip
        N, d = pooled.shape
        # x + mlp(norm(x)) with the LayerNorm folded into linear1 (vit.py:42)
        self.mlp.check_supported()
        p1, p2 = self.mlp.pack1(self.norm), self.mlp.pack2()
        stats = torch.empty(N, 2, device=pooled.device, dtype=torch.float32)
        hidden = torch.empty(N, self.mlp.linear1.out_features, device=pooled.device, dtype=torch.bfloat16)
        out = torch.empty_like(pooled)
        ops.row_stats(pooled, self.norm.eps, stats)
        ops.linear(pooled, p1.w, p1.bias, hidden, colsum=p1.colsum, rowstats=stats, gelu=True)
        ops.linear(hidden, p2.w, p2.bias, out, residual=pooled)
        return out


@compilable_module
class ViT(nn.Module):
    norm_eps = 1e-6

    def __init__(
        self,
        n_layers: int,
        d_model: int,
        n_heads: int,
        patch_size: int,
        img_size: int = 224,
        cls_token: bool = True,
        pool_type: str = "cls_token",
        dropout: float = 0.0,
    ) -> None:
        assert img_size % patch_size == 0
        super().__init__()
        grid = img_size // patch_size
        # nn.Conv2d is kept as the parameter container so state_dict / resize_pe / the loaders see (d, 3, p, p)
        self.patch_embed = nn.Conv2d(3, d_model, patch_size, patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, d_model)) if cls_token else None
        self.pe = nn.Parameter(torch.zeros(1, grid * grid, d_model))
        self.layers = Encoder(n_layers, d_model, n_heads=n_heads, dropout=dropout, norm_eps=self.norm_eps)
        self.norm = nn.LayerNorm(d_model, self.norm_eps)
        if pool_type == "cls_token":
            self.pooler = ClassTokenPooling()
        elif pool_type == "gap":
            self.pooler = GlobalAveragePooling()
        elif pool_type == "mha":
            self.pooler = MHAPooling(d_model, n_heads, norm_eps=self.norm_eps)
        else:
            raise KeyError(pool_type)
        self._pembed = _Packed()

    # -- packing -------------------------------------------------------------------------------
    def _pack_embed(self) -> SimpleNamespace:
        conv, cls = self.patch_embed, self.cls_token

        def build() -> SimpleNamespace:
            d = conv.out_channels
            k = conv.weight[0].numel()
            kpad = (k + 7) // 8 * 8
            w = torch.zeros(d, kpad, device=conv.weight.device, dtype=torch.bfloat16)
            w[:, :k] = conv.weight.detach().reshape(d, k).to(torch.bfloat16)
            bias = (conv.bias.detach().float() if conv.bias is not None
                    else torch.zeros(d, device=w.device)).contiguous()
            return SimpleNamespace(
                w=w, bias=bias, kpad=kpad,
                pe=self.pe.detach().to(torch.bfloat16).contiguous(),
                cls=None if cls is None else cls.detach().to(torch.bfloat16).contiguous(),
                cls_stats=None if cls is None else partial_stats_of(cls),
            )

        return self._pembed.get((conv.weight, conv.bias, self.pe, cls), build)

    # -- forward -------------------------------------------------------------------------------
    def embed(self, imgs: Tensor, with_stats: bool = False):
        """(N, 3, H, W) -> contiguous bf16 tokens (N, L, d) with the class token (if any) at position 0.
        ``with_stats``: also return the rows' partial LayerNorm statistics (N*L, ceil(d/128), 2), written by the
        patch-embedding GEMM's epilogue (and, for the class token, copied from a cached vector), or None."""
        if not imgs.is_cuda:
            raise RuntimeError("pytorch_models_b200 runs only on CUDA (sm_100a) tensors; there is no CPU fallback")
        if imgs.dtype not in (torch.bfloat16, torch.float32):
            imgs = imgs.float()
        imgs = imgs.contiguous()
        N, _, H, W = imgs.shape
        p = self.patch_embed.kernel_size[0]
        pk = self._pack_embed()
        P = (H // p) * (W // p)
        if H % p or W % p or P != self.pe.shape[1]:
            raise ValueError(f"image {H}x{W} gives {P} patches but pe has {self.pe.shape[1]}; call resize_pe first")
        d = self.patch_embed.out_channels
        off = 0 if self.cls_token is None else 1
        tokens = torch.empty(N, P + off, d, device=imgs.device, dtype=torch.bfloat16)
        stats = None
        if with_stats and self.layers.wants_stats():
            stats = torch.empty(N, P + off, (d + 127) // 128, 2, device=imgs.device, dtype=torch.float32)
        if p == 16 and imgs.dtype == torch.bfloat16 and tuple(self.patch_embed.kernel_size) == (16, 16) and N * 3 * (H // 16) < 2 ** 31:
            # im2col-free: the GEMM's A operand is the NCHW image itself (5-D tensor map, csrc/gemm.cuh kPatch)
            ops.patch_embed16(imgs, pk.w, pk.bias, pk.pe.view(P, d), tokens[:, off:, :], stats_out=stats,
                              stats_rows=P + off, stats_row_offset=off)
        else:  # other patch sizes (DINOv2: 14) and fp32 images: materialised patch rows (the kernel converts fp32 -> bf16)
            rows = torch.empty(N, P, pk.kpad, device=imgs.device, dtype=torch.bfloat16)
            ops.patch_rows(imgs, p, pk.kpad, rows)
            ops.linear(rows, pk.w, pk.bias, tokens[:, off:, :], residual=pk.pe, stats_out=stats, stats_rows=P + off,
                       stats_row_offset=off)
        if off:
            ops.cls_rows(pk.cls, tokens)
            if stats is not None:
                ops.broadcast_row(pk.cls_stats, stats)
        if with_stats:
            return tokens, (None if stats is None else stats.view(N * (P + off), -1, 2))
        return tokens

    @compilable(lambda self, x, extra: ((x.shape[0], self.norm.normalized_shape[0]), float_like(x)))
    def forward(self, imgs: Tensor) -> Tensor:
        out_dtype = imgs.dtype if imgs.dtype in (torch.bfloat16, torch.float32) else torch.float32
        if imgs.shape[0] == 0:
            return torch.empty(0, self.norm.normalized_shape[0], device=imgs.device, dtype=out_dtype)
        if not imgs.is_cuda:
            raise RuntimeError("pytorch_models_b200 runs only on CUDA (sm_100a) tensors; there is no CPU fallback")
        if imgs.dtype not in (torch.bfloat16, torch.float32):
            imgs = imgs.float()
        # the launches below are recorded once per (image shape, dtype, stream, weights) and replayed by one C-ABI
        # call afterwards (plans.py): 65 ctypes calls -> 1
        pooled = plans.run(self, (imgs.contiguous(),), self._forward_launches)
        return pooled if out_dtype == torch.bfloat16 else pooled.to(out_dtype)

    def _forward_launches(self, imgs: Tensor) -> Tensor:
        """bf16 / fp32 contiguous CUDA images -> pooled bf16 (N, d); libb200enc launches only (plan-recordable)."""
        tokens, stats = self.embed(imgs, with_stats=True)
        x = self.layers.run(tokens, stats)
        N, L, d = x.shape
        gamma, beta = norm_vectors(self.norm)
        if isinstance(self.pooler, ClassTokenPooling):
            # the pooler only reads token 0, so only those N rows are normalised (strided gather in the kernel)
            pooled = torch.empty(N, d, device=x.device, dtype=torch.bfloat16)
            ops.layernorm(x[:, 0, :], gamma, beta, self.norm.eps, pooled)
        else:
            normed = torch.empty_like(x)
            ops.layernorm(x.view(N * L, d), gamma, beta, self.norm.eps, normed.view(N * L, d))
            pooled = self.pooler(normed)
        return pooled

    @torch.no_grad()
    def resize_pe(self, size: int, interpolation_mode: str = "bicubic") -> None:
        """Interpolate the positional grid to ``size`` pixels (vit.py:87-94); host-side, rare, stays a PyTorch op."""
        old = int(self.pe.shape[1] ** 0.5)
        new = size // self.patch_embed.weight.shape[2]
        grid = self.pe.unflatten(1, (old, old)).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, (new, new), mode=interpolation_mode)
        self.pe = nn.Parameter(grid.permute(0, 2, 3, 1).flatten(1, 2))

    # -- weight loading (host side; mirrors load_flax_ckpt / load_facebook_state_dict, vit.py:151-200,257-306) ------
    def load_flax_arrays(self, arrays: dict, *, big_vision: bool = False) -> None:
        from .vit_weights import load_flax_arrays

        load_flax_arrays(self, arrays, big_vision=big_vision)

    def load_facebook_state_dict(self, state_dict: dict) -> None:
        from .vit_weights import load_facebook_state_dict

        load_facebook_state_dict(self, state_dict)

    # -- constructors --------------------------------------------------------------------------
    @staticmethod
    def _parse(model_tag: str, default_weights: str) -> tuple[str, int, str]:
        tag, _, weights = model_tag.partition("_")
        size, patch = tag.split("/")
        return size, int(patch), weights or default_weights

    @staticmethod
    def from_google(model_tag: str, *, pretrained: bool = False, **kwargs) -> "ViT":
        """Tags as in vit.py:96-119: "B/16", "B/16_augreg", "L/16_siglip" ..."""
        size, patch, weights = ViT._parse(model_tag, "augreg")
        n_layers, d_model, n_heads = _SIZES[size]
        extra = dict(cls_token=False, pool_type="mha") if weights == "siglip" else {}
        m = ViT(n_layers, d_model, n_heads, patch, **extra, **kwargs)
        if pretrained:
            from .vit_weights import load_google

            load_google(m, f"{size}/{patch}", weights, kwargs.get("img_size", 224))
        return m

    @staticmethod
    def from_facebook(model_tag: str, *, pretrained: bool = False, **kwargs) -> "ViT":
        """Tags as in vit.py:202-239: "S/16_deit3", "B/16_dino", "L/14_dinov2" (dinov2 defaults to 518 px)."""
        size, patch, weights = ViT._parse(model_tag, "deit3")
        if weights in ("deit3", "dino"):
            kwargs["img_size"] = kwargs.get("img_size", 224)
        elif weights == "dinov2":
            kwargs["img_size"] = kwargs.get("img_size", 518)
        else:
            raise ValueError(f"Unsupported {weights}")
        n_layers, d_model, n_heads = _SIZES[size]
        m = ViT(n_layers, d_model, n_heads, patch, **kwargs)
        if pretrained:
            from .vit_weights import load_facebook

            load_facebook(m, size, patch, weights, kwargs["img_size"])
        return m
