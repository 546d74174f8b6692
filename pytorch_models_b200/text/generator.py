"""Text generation around a decoder-only language model, API-compatible with the reference's ``DecoderGenerator``
(``pytorch_models/text/generator.py:11-39``): ``DecoderGenerator(model, tokenizer).generate(prompt, max_tokens, topk)``.

Behaviour kept from the reference: the prompt is encoded with ``tokenizer.encode``, one token is appended per model
call on the whole prefix (no KV cache), ``topk == 1`` is greedy, otherwise the next token is sampled from the softmax of
the ``topk`` largest logits, and generation stops after ``max_tokens`` new tokens or at ``tokenizer.eos_token_id``.
The model call is the sm_100a forward of ``GPT2`` / ``GPT`` (1-D token tensor in, ``(L, vocab)`` logits out).
"""
from __future__ import annotations

import torch
from torch import Tensor, nn


def _pick(last_logits: Tensor, topk: int) -> int:
    """Next token id from the logits of the last position."""
    scores = last_logits.float()
    if topk <= 1:
        return int(scores.argmax(dim=-1))
    best, ids = scores.topk(topk)
    choice = torch.multinomial(torch.softmax(best, dim=-1), num_samples=1)
    return int(ids[choice])


class DecoderGenerator:
    def __init__(self, model: nn.Module, tokenizer) -> None:
        self.model = model
        self.tokenizer = tokenizer

    @torch.inference_mode()
    def generate(self, prompt: str, max_tokens: int = 100, topk: int = 1) -> str:
        device = next(self.model.parameters()).device
        eos = self.tokenizer.eos_token_id
        ids = list(self.tokenizer.encode(prompt))
        for _ in range(max_tokens):
            prefix = torch.tensor(ids, device=device, dtype=torch.long)
            ids.append(_pick(self.model(prefix)[-1], topk))
            if ids[-1] == eos:
                break
        return self.tokenizer.decode(ids)
