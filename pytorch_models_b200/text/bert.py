"""B200-native drop-in for ``BERT`` (reference ``pytorch_models/text/bert.py:17-39``): token + position embedding,
embedding LayerNorm, then a *post-norm* encoder (``pre_norm=False``, ``norm_eps=1e-12``) on libb200enc kernels.

The embedding gather stays a PyTorch indexing op (host-side glue outside the hot path); everything
from the embedding LayerNorm on runs in the hand-written kernels.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor, nn

from .. import ops
from ..transformer import Encoder, norm_vectors


class BERT(nn.Module):
    def __init__(
        self,
        vocab_size: int,
        n_layers: int,
        d_model: int,
        max_seq_len: int = 512,
        dropout: float = 0.0,
        norm_eps: float = 1e-12,
    ) -> None:
        super().__init__()
        vocab_size = math.ceil(vocab_size / 64) * 64  # padded to a multiple of 64 like bert.py:28
        self.token_embs = nn.Embedding(vocab_size, d_model)
        self.pos_embs = nn.Parameter(torch.zeros(max_seq_len, d_model))
        self.norm = nn.LayerNorm(d_model, norm_eps)
        self.layers = Encoder(n_layers, d_model, dropout=dropout, pre_norm=False, norm_eps=norm_eps)

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("pytorch_models_b200 runs only on CUDA (sm_100a) tensors; there is no CPU fallback")
        out_dtype = self.token_embs.weight.dtype
        L = x.shape[-1]
        emb = (self.token_embs(x) + self.pos_embs[:L]).to(torch.bfloat16)
        emb3 = emb.reshape(-1, L, emb.shape[-1]).contiguous()
        B, _, d = emb3.shape
        gamma, beta = norm_vectors(self.norm)
        h = torch.empty_like(emb3)
        ops.layernorm(emb3.view(B * L, d), gamma, beta, self.norm.eps, h.view(B * L, d))
        y = self.layers.run(h).reshape(*x.shape, d)
        return y if out_dtype == torch.bfloat16 else y.to(out_dtype)
