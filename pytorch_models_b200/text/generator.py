"""Greedy / top-k text generation around a decoder-only model — reference ``DecoderGenerator``
(``pytorch_models/text/generator.py:11-39``): the whole prefix is re-run for every new token (no KV cache, as in the
reference); the forward itself is the sm_100a path of ``GPT2`` / ``GPT``."""
from __future__ import annotations

import torch
from torch import nn


class DecoderGenerator:
    def __init__(self, model: nn.Module, tokenizer) -> None:
        self.model = model
        self.tokenizer = tokenizer

    @torch.inference_mode()
    def generate(self, prompt: str, max_tokens: int = 100, topk: int = 1) -> str:
        device = next(self.model.parameters()).device
        tokens = self.tokenizer.encode(prompt)
        n = len(tokens)
        while len(tokens) - n < max_tokens:
            logits = self.model(torch.tensor(tokens, device=device))[-1]
            if topk == 1:  # greedy decoding
                token = logits.argmax(-1).item()
            else:  # top-k sampling (generator.py:30-32)
                top, indices = logits.float().topk(topk)
                token = indices[torch.multinomial(top.softmax(-1), 1).item()].item()
            tokens.append(token)
            if tokens[-1] == self.tokenizer.eos_token_id:
                break
        return self.tokenizer.decode(tokens)
