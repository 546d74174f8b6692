"""The oracles (numpy and torch-functional restatements) against the reference's own outputs in tests/golden/.

Tolerance 2e-5 abs/rel = what the reference holds itself to against timm (tests/image/test_vit.py:45)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import AUDIO_FIXTURES, FIXTURES, Golden
from oracle import oracle_np, oracle_torch

TOL = 2e-5


def run_np(g: Golden) -> dict:
    h = g.hyper
    if h["kind"] == "vit":
        pool = h.get("pool_type", "cls_token")
        return dict(
            pooled=oracle_np.vit_forward(g.sd, g.input, h["n_heads"], pool),
            tokens=oracle_np.vit_forward(g.sd, g.input, h["n_heads"], pool, return_tokens=True),
        )
    if h["kind"] == "whisper":
        return dict(tokens=oracle_np.whisper_encoder_forward(g.sd, g.input))
    if h["kind"] == "decoder":
        return dict(tokens=oracle_np.decoder(g.sd, g.input, g.extra.get("memory"), h["d_model"] // 64, h["pre_norm"],
                                             1e-5, prefix=""))
    if h["kind"] == "whisper_full":
        return dict(logits=oracle_np.whisper_forward(g.sd, g.input, g.extra["targets"]),
                    memory=oracle_np.whisper_encoder_forward(oracle_np.sub_dict(g.sd, "encoder."), g.input))
    if h["kind"] == "gpt2":
        return dict(logits=oracle_np.gpt2_forward(g.sd, g.input))
    if h["kind"] == "gpt":
        return dict(logits=oracle_np.gpt_forward(g.sd, g.input))
    return dict(tokens=oracle_np.bert_forward(g.sd, g.input))


@torch.no_grad()
def run_torch(g: Golden) -> dict:
    h, sd, x = g.hyper, g.torch_sd(), torch.from_numpy(np.array(g.input))
    if h["kind"] == "vit":
        pool = h.get("pool_type", "cls_token")
        return dict(
            pooled=oracle_torch.vit_forward(sd, x, h["n_heads"], pool).numpy(),
            tokens=oracle_torch.vit_forward(sd, x, h["n_heads"], pool, return_tokens=True).numpy(),
        )
    if h["kind"] == "whisper":
        return dict(tokens=oracle_torch.whisper_encoder_forward(sd, x).numpy())
    extra = {k: torch.from_numpy(np.array(v)) for k, v in g.extra.items()}
    if h["kind"] == "decoder":
        return dict(tokens=oracle_torch.decoder(sd, x, extra.get("memory"), h["d_model"] // 64, h["pre_norm"], 1e-5,
                                                prefix="").numpy())
    if h["kind"] == "whisper_full":
        return dict(logits=oracle_torch.whisper_forward(sd, x, extra["targets"]).numpy(),
                    memory=oracle_torch.whisper_encoder_forward(oracle_torch.sub_dict(sd, "encoder."), x).numpy())
    if h["kind"] == "gpt2":
        return dict(logits=oracle_torch.gpt2_forward(sd, x).numpy())
    if h["kind"] == "gpt":
        return dict(logits=oracle_torch.gpt_forward(sd, x).numpy())
    return dict(tokens=oracle_torch.bert_forward(sd, x).numpy())


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("impl", ["numpy", "torch"])
def test_oracle_matches_reference(golden, name, impl):
    g = golden(name)
    got = (run_np if impl == "numpy" else run_torch)(g)
    for key, expected in g.out.items():
        np.testing.assert_allclose(got[key], expected, rtol=TOL, atol=TOL, err_msg=f"{name}.{key} ({impl})")


@torch.no_grad()
def test_oracle_c1_vit_ti16(golden):
    """BASELINE configs[0]: ViT-Ti/16 224, batch 8, fp32 CPU — weights rebuilt from the recorded seeds."""
    import pytorch_models_b200 as pm

    g = golden("c1_vit_ti16")
    h = g.hyper
    torch.manual_seed(h["weight_seed"])
    sd = pm.ViT.from_google(h["tag"]).state_dict()
    oracle_torch.randomize_(sd, h["noise_seed"])
    torch.manual_seed(h["input_seed"])
    x = torch.randn(h["batch"], 3, 224, 224)
    pooled = oracle_torch.vit_forward(sd, x, 3)
    np.testing.assert_allclose(pooled.numpy(), g.out["pooled"], rtol=TOL, atol=TOL)
    tokens = oracle_torch.vit_forward(sd, x, 3, return_tokens=True)
    np.testing.assert_allclose(tokens.numpy(), g.out["tokens"], rtol=TOL, atol=TOL)
    # the numpy oracle on the same weights (2 samples: it is single-threaded)
    sd_np = {k: v.numpy() for k, v in sd.items()}
    np.testing.assert_allclose(oracle_np.vit_forward(sd_np, x[:2].numpy(), 3), g.out["pooled"][:2], rtol=TOL, atol=TOL)


def test_oracle_edge_cases():
    """Single token, single sample and non-square leading dims run through both oracles identically."""
    rng = np.random.default_rng(0)
    d, h = 64, 1
    sd = {}
    for name, shape in [("sa.q_proj", (d, d)), ("sa.k_proj", (d, d)), ("sa.v_proj", (d, d)), ("sa.out_proj", (d, d)),
                        ("mlp.linear1", (4 * d, d)), ("mlp.linear2", (d, 4 * d))]:
        sd[f"layers.0.{name}.weight"] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
        sd[f"layers.0.{name}.bias"] = (rng.standard_normal(shape[0]) * 0.1).astype(np.float32)
    for name in ("sa_norm", "mlp_norm"):
        sd[f"layers.0.{name}.weight"] = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
        sd[f"layers.0.{name}.bias"] = (0.1 * rng.standard_normal(d)).astype(np.float32)
    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    for shape in [(1, 1, d), (2, 3, 5, d), (1, 130, d)]:
        x = rng.standard_normal(shape).astype(np.float32)
        for pre in (True, False):
            a = oracle_np.encoder(sd, x, h, pre, 1e-5)
            b = oracle_torch.encoder(sd_t, torch.from_numpy(x), h, pre, 1e-5).numpy()
            np.testing.assert_allclose(a, b, rtol=TOL, atol=TOL)


@pytest.mark.parametrize("name", AUDIO_FIXTURES)
def test_logmel_oracles_match_reference(golden, name):
    """WhisperPreprocessor (whisper.py:138-148): numpy restatement (explicit framing + rfft) and the torch.stft one
    against the reference's output, including a stretch of digital silence (log10(0) = -inf under the max - 8 floor)."""
    g = golden(name)
    want = g.out["logmel"]
    got_np = oracle_np.whisper_logmel(g.input, g.sd["filters"])
    np.testing.assert_allclose(got_np, want, rtol=0, atol=2e-5)
    got_t = oracle_torch.whisper_logmel(torch.from_numpy(np.array(g.input)), torch.from_numpy(np.array(g.sd["filters"])))
    np.testing.assert_allclose(got_t.numpy(), want, rtol=0, atol=2e-5)
    assert want.shape == (g.input.shape[0], g.sd["filters"].shape[0], g.input.shape[1] // 160)
