"""B200-native drop-in for ``WhisperEncoder`` (reference ``pytorch_models/audio2text/whisper.py:11-34``).

Same constructor, ``state_dict`` keys (``stem.0/2``, ``pos_embs`` buffer, ``layers.*``, ``norm``) and call signature.
The conv stem runs as two GEMMs on the same tcgen05 kernel as the linears: the log-mel input is rewritten once as
zero-padded time-major rows, after which a k=3 convolution over time is a GEMM over an *overlapping* strided view
(row t starts at element stride·t·C and is 3·C long), with GELU — and for the second conv the positional embedding —
fused into the epilogue (whisper.py:16-21,30-31).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import Tensor, nn

from .. import ops
from ..transformer import Encoder, _Packed, norm_vectors


class WhisperEncoder(nn.Module):
    max_seq_len = 3000

    def __init__(self, n_layers: int, d_model: int, n_mels: int = 80, dropout: float = 0.0) -> None:
        super().__init__()
        # parameter containers with the reference's layout; the arithmetic does not go through nn.Conv1d.forward
        self.stem = nn.Sequential(
            nn.Conv1d(n_mels, d_model, 3, 1, 1),
            nn.GELU(),
            nn.Conv1d(d_model, d_model, 3, 2, 1),
            nn.GELU(),
        )
        self.register_buffer("pos_embs", torch.zeros(self.max_seq_len // 2, d_model))
        self.pos_embs: Tensor
        self.layers = Encoder(n_layers, d_model, dropout=dropout)
        self.norm = nn.LayerNorm(d_model)
        self._pstem = _Packed()

    def _pack_stem(self) -> SimpleNamespace:
        c1, c2 = self.stem[0], self.stem[2]

        def conv_as_linear(conv: nn.Conv1d) -> tuple[Tensor, Tensor]:
            # w[n][k*C + c] = weight[n][c][k]: matches rows that concatenate time steps t-1, t, t+1
            w = conv.weight.detach().permute(0, 2, 1).reshape(conv.out_channels, -1).to(torch.bfloat16).contiguous()
            b = (conv.bias.detach().float() if conv.bias is not None
                 else torch.zeros(conv.out_channels, device=w.device)).contiguous()
            return w, b

        def build() -> SimpleNamespace:
            w1, b1 = conv_as_linear(c1)
            w2, b2 = conv_as_linear(c2)
            return SimpleNamespace(w1=w1, b1=b1, w2=w2, b2=b2, pos=self.pos_embs.detach().to(torch.bfloat16).contiguous())

        return self._pstem.get((c1.weight, c1.bias, c2.weight, c2.bias, self.pos_embs), build)

    def embed(self, x: Tensor) -> Tensor:
        """(N, n_mels, T) log-mel -> contiguous bf16 tokens (N, ceil(T/2), d) incl. positional embedding."""
        if not x.is_cuda:
            raise RuntimeError("pytorch_models_b200 runs only on CUDA (sm_100a) tensors; there is no CPU fallback")
        if x.dtype not in (torch.bfloat16, torch.float32):
            x = x.float()
        x = x.contiguous()
        N, C, T = x.shape
        d = self.stem[0].out_channels
        if C % 8 != 0:
            raise NotImplementedError(f"n_mels={C} must be a multiple of 8 for 16-byte aligned rows")
        T2 = (T + 2 - 3) // 2 + 1  # Conv1d(k=3, stride=2, pad=1) output length
        if T2 > self.pos_embs.shape[0]:
            raise ValueError(f"{T} frames give {T2} tokens but pos_embs has {self.pos_embs.shape[0]} rows")
        pk = self._pack_stem()
        dev = x.device
        rows = torch.empty(N, T + 2, C, device=dev, dtype=torch.bfloat16)
        ops.time_rows(x, rows)
        # conv1 (stride 1): output step t reads rows[t : t+3]; written at rows 1..T of a zero-padded buffer
        h1 = torch.empty(N, T + 3, d, device=dev, dtype=torch.bfloat16)
        h1[:, 0].zero_()
        h1[:, T + 1:].zero_()
        a1 = rows.as_strided((N, T, 3 * C), ((T + 2) * C, C, 1))
        ops.linear(a1, pk.w1, pk.b1, h1[:, 1:T + 1, :], gelu=True)
        # conv2 (stride 2): output step t reads h1 rows[2t : 2t+3]  (= time steps 2t-1, 2t, 2t+1)
        a2 = h1.as_strided((N, T2, 3 * d), ((T + 3) * d, 2 * d, 1))
        tokens = torch.empty(N, T2, d, device=dev, dtype=torch.bfloat16)
        ops.linear(a2, pk.w2, pk.b2, tokens, gelu=True, residual=pk.pos[:T2].unsqueeze(0))
        return tokens

    def forward(self, x: Tensor) -> Tensor:
        out_dtype = x.dtype if x.dtype in (torch.bfloat16, torch.float32) else torch.float32
        h = self.layers.run(self.embed(x))
        N, L, d = h.shape
        gamma, beta = norm_vectors(self.norm)
        out = torch.empty_like(h)
        ops.layernorm(h.view(N * L, d), gamma, beta, self.norm.eps, out.view(N * L, d))
        return out if out_dtype == torch.bfloat16 else out.to(out_dtype)

    @torch.no_grad()
    def load_openai_state_dict(self, state_dict: dict) -> None:
        """Encoder half of an openai-whisper checkpoint (``model_state_dict``), cf. whisper.py:97-135."""
        sd = {k[len("encoder."):]: v for k, v in state_dict.items() if k.startswith("encoder.")}

        def take(module, key: str) -> None:
            module.weight.copy_(sd.pop(f"{key}.weight"))
            if module.bias is not None:
                module.bias.copy_(sd.pop(f"{key}.bias", 0))  # the key projection has no bias in the checkpoint

        take(self.stem[0], "conv1")
        take(self.stem[2], "conv2")
        self.pos_embs.copy_(sd.pop("positional_embedding"))
        for i, layer in enumerate(self.layers):
            pre = f"blocks.{i}"
            take(layer.sa.q_proj, f"{pre}.attn.query")
            take(layer.sa.k_proj, f"{pre}.attn.key")
            take(layer.sa.v_proj, f"{pre}.attn.value")
            take(layer.sa.out_proj, f"{pre}.attn.out")
            take(layer.sa_norm, f"{pre}.attn_ln")
            take(layer.mlp.linear1, f"{pre}.mlp.0")
            take(layer.mlp.linear2, f"{pre}.mlp.2")
            take(layer.mlp_norm, f"{pre}.mlp_ln")
        take(self.norm, "ln_post")
