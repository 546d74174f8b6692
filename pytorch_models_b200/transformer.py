"""B200-native drop-in for the encoder half of ``pytorch_models/transformer.py`` (reference lines cited inline).

Same class names, constructor signatures, attribute names and ``state_dict`` keys as the reference, so
``new.load_state_dict(ref.state_dict())`` is strict-clean and the reference's weight loaders (which write in place
into ``layer.sa.q_proj.weight`` etc.) keep working. The arithmetic does not run in PyTorch: ``forward`` sequences
hand-written sm_100a kernels from ``libb200enc.so``:

    pre-norm layer (transformer.py:125-126), 5 launches (+1 row_stats for the first layer of a stack)
        linear [3·inner, d], LayerNorm folded -> fused q|k|v                  (transformer.py:87,47-49)
        attention                             -> softmax(q kᵀ/√64) v          (transformer.py:52)
        linear out_proj + bias + residual     -> x1 (+ partial LN statistics) (transformer.py:53,125)
        linear1, LayerNorm folded, erf-GELU   -> hidden                       (transformer.py:93,59-61)
        linear2 + bias + residual             -> x2 (+ partial LN statistics) (transformer.py:66,126)

    The row statistics (mean, rstd) a folded LayerNorm needs are produced by the epilogue of the GEMM that wrote
    its input (per-128-column (mean, M2) partials, combined in a fixed order by the consumer), so no separate pass
    over the residual stream is needed after the first layer.

Only what the kernels implement is accepted (self/cross attention without mask, head_dim 64, exact GELU, eval mode);
anything else raises ``NotImplementedError`` — there is no PyTorch fallback.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
from torch import Tensor, nn

from . import ops

_SUPPORTED_HEAD_DIM = 64
_MAX_STAT_PARTS = 12  # libb200enc combines at most 12 partial statistics per row (d <= 1536)


# ----------------------------------------------------------------------------------------------- weight packing
def _key(*params: Tensor | None) -> tuple:
    return tuple(None if p is None else (p.data_ptr(), p._version, p.dtype, p.device) for p in params)


class _Packed:
    """Kernel-ready copies of a module's parameters, rebuilt whenever a source tensor is replaced or mutated in place
    (the reference loaders do both: vit.py:172-197 copy_, vit.py:290-304 mul_)."""

    def __init__(self) -> None:
        self.key: tuple | None = None
        self.t: SimpleNamespace | None = None

    def get(self, params: tuple, build) -> SimpleNamespace:
        key = _key(*params)
        if key != self.key:
            with torch.no_grad():
                self.t = build()
            self.key = key
        return self.t


def _cat_bias(linears: list[nn.Linear]) -> Tensor:
    parts = []
    for lin in linears:
        if lin.bias is None:
            parts.append(torch.zeros(lin.out_features, device=lin.weight.device, dtype=torch.float32))
        else:
            parts.append(lin.bias.detach().float())
    return torch.cat(parts)


def pack_plain(linears: list[nn.Linear]) -> SimpleNamespace:
    """bf16 [sum N, K] weight + fp32 bias of one or more linears that share their input."""
    w = torch.cat([lin.weight.detach() for lin in linears]).to(torch.bfloat16).contiguous()
    return SimpleNamespace(w=w, bias=_cat_bias(linears).contiguous(), colsum=None)


def pack_folded(linears: list[nn.Linear], norm: nn.LayerNorm) -> SimpleNamespace:
    """LayerNorm folded into the following linear:  LN(x) Wᵀ + b = rstd·(x W'ᵀ − mean·s) + c  with
    W' = W ⊙ gamma (bf16), s = row sums of the *rounded* W' (so the mean term cancels exactly against what the
    tensor cores accumulate) and c = W beta + b."""
    w32 = torch.cat([lin.weight.detach().float() for lin in linears])
    gamma, beta = norm.weight.detach().float(), norm.bias.detach().float()
    wg = (w32 * gamma[None, :]).to(torch.bfloat16).contiguous()
    colsum = wg.float().sum(dim=1).contiguous()
    c = (w32 @ beta + _cat_bias(linears)).contiguous()
    return SimpleNamespace(w=wg, bias=c, colsum=colsum)


def _as_tokens(x: Tensor, d: int) -> tuple[Tensor, tuple]:
    """(*, L, d) any dtype/strides -> contiguous bf16 (B, L, d) plus what is needed to restore the caller's view."""
    if not x.is_cuda:
        raise RuntimeError(
            "pytorch_models_b200 runs only on CUDA (sm_100a) tensors; there is no CPU fallback "
            f"(got a {x.device} tensor)"
        )
    if x.dim() < 2 or x.shape[-1] != d:
        raise ValueError(f"expected (*, L, {d}) input, got {tuple(x.shape)}")
    lead = x.shape[:-2]
    x3 = x.reshape(-1, x.shape[-2], d) if x.dim() != 3 else x
    if x3.dtype != torch.bfloat16 or not x3.is_contiguous():
        x3 = x3.to(torch.bfloat16).contiguous()
    return x3, (lead, x.dtype)


def _restore(y3: Tensor, meta: tuple) -> Tensor:
    lead, dtype = meta
    y = y3 if len(lead) == 1 else y3.reshape(*lead, y3.shape[-2], y3.shape[-1])
    return y if dtype == torch.bfloat16 else y.to(dtype)


# ----------------------------------------------------------------------------------------------- modules
class MHA(nn.Module):
    """Multi-head attention, reference ``MHA`` (transformer.py:9-53)."""

    def __init__(
        self,
        d_model: int,
        n_heads: int | None = None,
        head_dim: int | None = None,
        bias: bool = True,
        dropout: float = 0.0,
    ) -> None:
        super().__init__()
        # same defaulting rule as transformer.py:20-26
        if n_heads is None and head_dim is None:
            head_dim = 64
        if head_dim is None:
            head_dim = d_model // n_heads
        if n_heads is None:
            n_heads = d_model // head_dim
        inner = n_heads * head_dim
        self.q_proj = nn.Linear(d_model, inner, bias)
        self.k_proj = nn.Linear(d_model, inner, bias)
        self.v_proj = nn.Linear(d_model, inner, bias)
        self.out_proj = nn.Linear(inner, d_model, bias)
        self.n_heads = n_heads
        self.head_dim = head_dim
        self.dropout = dropout
        self._packs = {name: _Packed() for name in ("qkv", "q", "kv", "k", "v", "out")}

    # -- packing -------------------------------------------------------------------------------
    def _pack(self, name: str, linears: list[nn.Linear]) -> SimpleNamespace:
        params = tuple(p for lin in linears for p in (lin.weight, lin.bias))
        return self._packs[name].get(params, lambda: pack_plain(linears))

    def check_supported(self, attn_bias: Tensor | None = None, causal: bool = False) -> None:
        if attn_bias is not None:
            raise NotImplementedError("attn_bias is not supported by the sm_100a attention kernel")
        if causal:
            raise NotImplementedError("causal attention is not supported by the sm_100a attention kernel")
        if self.head_dim != _SUPPORTED_HEAD_DIM:
            raise NotImplementedError(f"head_dim={self.head_dim}: the sm_100a attention kernel is specialised on 64")
        if self.training and self.dropout > 0.0:
            raise NotImplementedError("attention dropout (training mode) is not supported; call .eval()")

    @property
    def scale(self) -> float:
        return 1.0 / math.sqrt(self.head_dim)  # F.scaled_dot_product_attention default (transformer.py:52)

    def forward(
        self,
        q: Tensor,
        k: Tensor | None = None,
        v: Tensor | None = None,
        attn_bias: Tensor | None = None,
        causal: bool = False,
    ) -> Tensor:
        self.check_supported(attn_bias, causal)
        d_in = self.q_proj.in_features
        inner = self.n_heads * self.head_dim
        q3, meta = _as_tokens(q, d_in)
        B, Lq, _ = q3.shape
        dev = q3.device
        if k is None and v is None:
            pk = self._pack("qkv", [self.q_proj, self.k_proj, self.v_proj])
            qkv = torch.empty(B, Lq, 3 * inner, device=dev, dtype=torch.bfloat16)
            ops.linear(q3.view(B * Lq, d_in), pk.w, pk.bias, qkv.view(B * Lq, 3 * inner))
            qv, kv_k, kv_v = qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:]
        else:
            k = q if k is None else k
            k3, _ = _as_tokens(k, d_in)
            Bk, Lkv, _ = k3.shape
            pq = self._pack("q", [self.q_proj])
            qv = torch.empty(B, Lq, inner, device=dev, dtype=torch.bfloat16)
            ops.linear(q3.view(B * Lq, d_in), pq.w, pq.bias, qv.view(B * Lq, inner))
            if v is None or v is k:
                pkv = self._pack("kv", [self.k_proj, self.v_proj])
                kvbuf = torch.empty(Bk, Lkv, 2 * inner, device=dev, dtype=torch.bfloat16)
                ops.linear(k3.view(Bk * Lkv, d_in), pkv.w, pkv.bias, kvbuf.view(Bk * Lkv, 2 * inner))
            else:
                v3, _ = _as_tokens(v, d_in)
                kvbuf = torch.empty(Bk, Lkv, 2 * inner, device=dev, dtype=torch.bfloat16)
                pkk, pvv = self._pack("k", [self.k_proj]), self._pack("v", [self.v_proj])
                ops.linear(k3, pkk.w, pkk.bias, kvbuf[:, :, :inner])
                ops.linear(v3, pvv.w, pvv.bias, kvbuf[:, :, inner:])
            kv_k, kv_v = kvbuf[:, :, :inner], kvbuf[:, :, inner:]
            if B != Bk:  # broadcast query (the MAP-pooling probe, vit.py:35,41)
                if B != 1:
                    raise ValueError("query batch must be 1 or match the key batch")
                qv = qv.expand(Bk, Lq, inner).contiguous()
                B = Bk
        att = torch.empty(B, Lq, inner, device=dev, dtype=torch.bfloat16)
        ops.attention(qv, kv_k, kv_v, att, self.n_heads, self.scale)
        po = self._pack("out", [self.out_proj])
        out = torch.empty(B, Lq, self.out_proj.out_features, device=dev, dtype=torch.bfloat16)
        ops.linear(att.view(B * Lq, inner), po.w, po.bias, out.view(B * Lq, -1))
        if B != q3.shape[0]:
            meta = ((B,), meta[1])
        return _restore(out, meta)


_ACTS = dict(
    gelu=nn.GELU,
    approximate_gelu=lambda: nn.GELU(approximate="tanh"),
    relu=lambda: nn.ReLU(inplace=True),
    silu=nn.SiLU,
)


class MLP(nn.Sequential):
    """``linear1 -> act -> linear2 -> dropout`` with the reference's child names/order (transformer.py:56-67)."""

    def __init__(self, in_dim: int, hidden_dim: float, dropout: float = 0.0, act: str = "gelu") -> None:
        super().__init__()
        self.linear1 = nn.Linear(in_dim, hidden_dim)
        self.act = _ACTS[act]()
        self.linear2 = nn.Linear(hidden_dim, in_dim)
        self.dropout = nn.Dropout(dropout)
        self._act_name = act
        self._p1, self._p2 = _Packed(), _Packed()

    def check_supported(self) -> None:
        if self._act_name != "gelu":
            raise NotImplementedError(f"act={self._act_name!r}: only exact (erf) GELU is fused into the sm_100a GEMM epilogue")
        if self.training and self.dropout.p > 0.0:
            raise NotImplementedError("MLP dropout (training mode) is not supported; call .eval()")

    def pack1(self, norm: nn.LayerNorm | None) -> SimpleNamespace:
        lin = self.linear1
        if norm is None:
            return self._p1.get((lin.weight, lin.bias), lambda: pack_plain([lin]))
        return self._p1.get((lin.weight, lin.bias, norm.weight, norm.bias), lambda: pack_folded([lin], norm))

    def pack2(self) -> SimpleNamespace:
        lin = self.linear2
        return self._p2.get((lin.weight, lin.bias), lambda: pack_plain([lin]))

    def forward(self, x: Tensor) -> Tensor:
        self.check_supported()
        d = self.linear1.in_features
        x3, meta = _as_tokens(x.unsqueeze(0) if x.dim() == 2 else x, d)
        B, L, _ = x3.shape
        p1, p2 = self.pack1(None), self.pack2()
        hidden = torch.empty(B * L, self.linear1.out_features, device=x3.device, dtype=torch.bfloat16)
        out = torch.empty(B, L, d, device=x3.device, dtype=torch.bfloat16)
        ops.linear(x3.view(B * L, d), p1.w, p1.bias, hidden, gelu=True)
        ops.linear(hidden, p2.w, p2.bias, out.view(B * L, d))
        y = _restore(out, meta)
        return y.squeeze(0) if x.dim() == 2 else y


class EncoderLayer(nn.Module):
    """Reference ``EncoderLayer`` (transformer.py:108-130; fields created at :84-94 with ``cross_attn=False``)."""

    def __init__(
        self,
        d_model: int,
        n_heads: int | None = None,
        head_dim: int | None = None,
        bias: bool = True,
        mlp_ratio: float = 4.0,
        dropout: float = 0.0,
        act: str = "gelu",
        pre_norm: bool = True,
        norm_eps: float = 1e-5,
    ) -> None:
        super().__init__()
        self.pre_norm = pre_norm
        self.sa_norm = nn.LayerNorm(d_model, norm_eps)
        self.sa = MHA(d_model, n_heads, head_dim, bias, dropout)
        self.ca_norm = None  # attribute kept for parity with the reference's DecoderLayer base (transformer.py:90-91)
        self.ca = None
        self.mlp_norm = nn.LayerNorm(d_model, norm_eps)
        self.mlp = MLP(d_model, int(d_model * mlp_ratio), dropout, act)
        self._pqkv = _Packed()

    # -- packing -------------------------------------------------------------------------------
    def _pack_qkv(self) -> SimpleNamespace:
        sa = self.sa
        lins = [sa.q_proj, sa.k_proj, sa.v_proj]
        params = tuple(p for lin in lins for p in (lin.weight, lin.bias))
        if self.pre_norm:
            params += (self.sa_norm.weight, self.sa_norm.bias)
            return self._pqkv.get(params, lambda: pack_folded(lins, self.sa_norm))
        return self._pqkv.get(params, lambda: pack_plain(lins))

    def workspace(self, B: int, L: int, device: torch.device) -> SimpleNamespace:
        d = self.sa_norm.normalized_shape[0]
        inner = self.sa.n_heads * self.sa.head_dim
        M = B * L
        e = lambda *s, dt=torch.bfloat16: torch.empty(*s, device=device, dtype=dt)  # noqa: E731
        parts = (d + 127) // 128
        fused = self.pre_norm and parts <= _MAX_STAT_PARTS
        return SimpleNamespace(
            qkv=e(B, L, 3 * inner), att=e(B, L, inner), hidden=e(M, self.mlp.linear1.out_features), mid=e(M, d),
            tmp=None if self.pre_norm else e(M, d), stats=e(M, 2, dt=torch.float32),
            parts_mid=e(M, parts, 2, dt=torch.float32) if fused else None,
            parts_out=e(M, parts, 2, dt=torch.float32) if fused else None,
        )

    def run(self, x3: Tensor, out3: Tensor, ws: SimpleNamespace, stats_in: Tensor | None = None,
            want_stats: bool = False) -> Tensor | None:
        """x3 (B, L, d) bf16 contiguous -> out3 (same shape, must not alias x3).

        ``stats_in``: partial LayerNorm statistics of x3 written by the producing GEMM (else a row_stats pass runs).
        Returns the partial statistics of out3 when ``want_stats`` (for the next layer's sa_norm), else None."""
        sa, mlp = self.sa, self.mlp
        sa.check_supported()
        mlp.check_supported()
        B, L, d = x3.shape
        M = B * L
        inner = sa.n_heads * sa.head_dim
        x2, out2 = x3.view(M, d), out3.view(M, d)
        pq = self._pack_qkv()
        po = sa._pack("out", [sa.out_proj])
        p1 = mlp.pack1(self.mlp_norm if self.pre_norm else None)
        p2 = mlp.pack2()
        qkv2 = ws.qkv.view(M, 3 * inner)
        q, k, v = ws.qkv[:, :, :inner], ws.qkv[:, :, inner:2 * inner], ws.qkv[:, :, 2 * inner:]
        if self.pre_norm:
            if stats_in is None:
                stats_in = ops.row_stats(x2, self.sa_norm.eps, ws.stats)
            ops.linear(x2, pq.w, pq.bias, qkv2, colsum=pq.colsum, rowstats=stats_in, ln_eps=self.sa_norm.eps)
            ops.attention(q, k, v, ws.att, sa.n_heads, sa.scale)
            ops.linear(ws.att.view(M, inner), po.w, po.bias, ws.mid, residual=x2, stats_out=ws.parts_mid)
            mid_stats = ws.parts_mid
            if mid_stats is None:
                mid_stats = ops.row_stats(ws.mid, self.mlp_norm.eps, ws.stats)
            ops.linear(ws.mid, p1.w, p1.bias, ws.hidden, colsum=p1.colsum, rowstats=mid_stats,
                       ln_eps=self.mlp_norm.eps, gelu=True)
            out_stats = ws.parts_out if want_stats else None
            ops.linear(ws.hidden, p2.w, p2.bias, out2, residual=ws.mid, stats_out=out_stats)
            return out_stats
        else:  # post-norm (BERT): transformer.py:128-129
            g1, b1 = norm_vectors(self.sa_norm)
            g2, b2 = norm_vectors(self.mlp_norm)
            ops.linear(x2, pq.w, pq.bias, qkv2)
            ops.attention(q, k, v, ws.att, sa.n_heads, sa.scale)
            ops.linear(ws.att.view(M, inner), po.w, po.bias, ws.tmp, residual=x2)
            ops.layernorm(ws.tmp, g1, b1, self.sa_norm.eps, ws.mid)
            ops.linear(ws.mid, p1.w, p1.bias, ws.hidden, gelu=True)
            ops.linear(ws.hidden, p2.w, p2.bias, ws.tmp, residual=ws.mid)
            ops.layernorm(ws.tmp, g2, b2, self.mlp_norm.eps, out2)
        return None

    def forward(self, x: Tensor) -> Tensor:
        d = self.sa_norm.normalized_shape[0]
        x3, meta = _as_tokens(x, d)
        out3 = torch.empty_like(x3)
        self.run(x3, out3, self.workspace(x3.shape[0], x3.shape[1], x3.device))
        return _restore(out3, meta)


def norm_vectors(norm: nn.LayerNorm) -> tuple[Tensor, Tensor]:
    """fp32 contiguous (gamma, beta) of a LayerNorm, cached on the module until the parameters change."""
    key = _key(norm.weight, norm.bias)
    hit = norm.__dict__.get("_b200_vectors")
    if hit is None or hit[0] != key:
        with torch.no_grad():
            hit = (key, (norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous()))
        norm.__dict__["_b200_vectors"] = hit
    return hit[1]


class Encoder(nn.Sequential):
    """Reference ``Encoder`` (transformer.py:133-149): an ``nn.Sequential`` of ``EncoderLayer`` — iteration, ``len``
    and indexing behave the same; ``forward`` additionally shares one workspace across the layers."""

    def __init__(
        self,
        n_layers: int,
        d_model: int,
        n_heads: int | None = None,
        head_dim: int | None = None,
        bias: bool = True,
        mlp_ratio: float = 4.0,
        dropout: float = 0.0,
        act: str = "gelu",
        pre_norm: bool = True,
        norm_eps: float = 1e-5,
    ) -> None:
        super().__init__()
        for _ in range(n_layers):
            self.append(EncoderLayer(d_model, n_heads, head_dim, bias, mlp_ratio, dropout, act, pre_norm, norm_eps))
        self.d_model = d_model

    def run(self, x3: Tensor) -> Tensor:
        """bf16 contiguous (B, L, d) -> new tensor of the same shape; x3 is left untouched."""
        layers = list(self)
        if not layers or x3.numel() == 0:
            return x3.clone()
        B, L, _ = x3.shape
        ws = layers[0].workspace(B, L, x3.device)
        bufs = [torch.empty_like(x3), torch.empty_like(x3) if len(layers) > 1 else None]
        cur, stats = x3, None
        for i, layer in enumerate(layers):
            stats = layer.run(cur, bufs[i % 2], ws, stats_in=stats, want_stats=i + 1 < len(layers))
            cur = bufs[i % 2]
        return cur

    def forward(self, x: Tensor) -> Tensor:
        x3, meta = _as_tokens(x, self.d_model)
        return _restore(self.run(x3), meta)
