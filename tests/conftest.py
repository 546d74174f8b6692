"""Shared test helpers. GPU tests are marked `@pytest.mark.gpu`; everything else must pass on a CPU-only box."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

FIXTURES = ["vit_cls", "vit_gap", "vit_siglip", "vit_p14", "vit_long", "whisper", "bert",
            "decoder_postnorm_cross", "decoder_causal_long", "whisper_full", "gpt2", "gpt"]
AUDIO_FIXTURES = ["logmel_tiny", "logmel_large_v3"]  # fp32 front end: own (much tighter) tolerance


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


class Golden:
    """One tests/golden/<name>.npz: reference weights (sd), input, reference outputs (out) and hyper-parameters."""

    def __init__(self, name: str) -> None:
        z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        self.name = name
        self.hyper = json.loads(bytes(z["hyper"]).decode())
        self.input = z["input"]
        self.sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
        self.out = {k[4:]: z[k] for k in z.files if k.startswith("out.")}
        self.extra = {k[3:]: z[k] for k in z.files if k.startswith("in.")}  # further inputs (memory, targets)

    def torch_sd(self):
        import torch

        return {k: torch.from_numpy(np.array(v)) for k, v in self.sd.items()}


@pytest.fixture(scope="session")
def golden():
    cache: dict[str, Golden] = {}

    def get(name: str) -> Golden:
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]

    return get


def build_model(g: Golden):
    """Instantiate the product module that corresponds to a fixture and load the reference weights into it."""
    import pytorch_models_b200 as pm

    h = dict(g.hyper)
    kind = h.pop("kind")
    if kind == "vit":
        m = pm.ViT(**h)
    elif kind == "whisper":
        m = pm.WhisperEncoder(h["n_layers"], h["d_model"], h["n_mels"])
    elif kind == "bert":
        m = pm.BERT(h["vocab_size"], h["n_layers"], h["d_model"])
    elif kind == "decoder":
        m = pm.Decoder(h["n_layers"], h["d_model"], cross_attn=h["cross_attn"], pre_norm=h["pre_norm"])
    elif kind == "whisper_full":
        m = pm.Whisper(h["vocab_size"], h["n_layers"], h["d_model"], h["n_mels"])
    elif kind == "logmel":
        m = pm.WhisperPreprocessor(h["variant"])
    elif kind in ("gpt2", "gpt"):
        base = pm.GPT2 if kind == "gpt2" else pm.GPT
        # vocab_size is a class attribute, as in the reference (gpt2.py:12, gpt.py:15)
        m = type(f"{base.__name__}Small", (base,), dict(vocab_size=h["vocab_size"]))(h["n_layers"], h["d_model"])
    else:
        raise KeyError(kind)
    m.load_state_dict(g.torch_sd(), strict=True)
    return m.eval()


def error_stats(actual, expected):
    """max-abs error and the minimum per-sample cosine similarity (both in fp64)."""
    a = np.asarray(actual, dtype=np.float64).reshape(expected.shape[0], -1)
    e = np.asarray(expected, dtype=np.float64).reshape(expected.shape[0], -1)
    cos = (a * e).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(e, axis=1) + 1e-30)
    return float(np.abs(a - e).max()), float(cos.min())
