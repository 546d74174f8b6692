"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C-ABI, against
 (a) the committed outputs of the reference itself (tests/golden/*.npz), and
 (b) the fp32 CPU oracle on seeded inputs at the BASELINE configs' shapes.

Stated tolerance (SURVEY §8(c) / BASELINE.md §4), for post-LayerNorm outputs with RMS ~ 1, bf16 kernels vs the fp32
reference: max-abs <= 0.125 for <= 12 layers, <= 0.15 for 24-32 layers, per-sample cosine >= 0.9999 — the numbers
BASELINE.md §4 / SURVEY §8(c) state. Calibration (profiles/r01/parity_report.json, measured on the B200): PyTorch's own
bf16 forward of the same math has max-abs 0.052 / 0.011 / 0.047 / 0.135 on C2 / C3 / C4 / C5 against the same fp32
oracle (ours: 0.049 / 0.009 / 0.047 / 0.135); a larger error is a bug, not "bf16 noise". The bound is applied as stated
to every fixture whose outputs have unit scale; only BERT's post-norm stack (outputs up to |x| ~ 12, SURVEY §8(c):
the reference's own bf16 error there is 0.31) scales it with the output range.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import AUDIO_FIXTURES, FIXTURES, build_model, error_stats
from oracle import oracle_torch

pytestmark = pytest.mark.gpu

MAX_ABS_12, MAX_ABS_32, MIN_COS = 0.125, 0.15, 0.9999


def _run_fixture(g, m):
    x = torch.from_numpy(np.array(g.input)).cuda()
    kind = g.hyper["kind"]
    with torch.no_grad():
        if kind == "vit":
            tokens = m.layers.run(m.embed(x))
            gamma, beta = m.norm.weight.float(), m.norm.bias.float()
            normed = torch.empty_like(tokens)
            from pytorch_models_b200 import ops

            ops.layernorm(tokens.view(-1, tokens.shape[-1]), gamma.contiguous(), beta.contiguous(), m.norm.eps,
                          normed.view(-1, tokens.shape[-1]))
            return dict(pooled=m(x).float().cpu().numpy(), tokens=normed.float().cpu().numpy())
        extra = {k: torch.from_numpy(np.array(v)).cuda() for k, v in g.extra.items()}
        if kind == "decoder":
            return dict(tokens=m(x, extra.get("memory")).float().cpu().numpy())
        if kind == "whisper_full":
            return dict(logits=m(x, extra["targets"]).float().cpu().numpy(),
                        memory=m.encoder(x).float().cpu().numpy())
        if kind in ("gpt2", "gpt"):
            return dict(logits=m(x).float().cpu().numpy())
        return dict(tokens=m(x).float().cpu().numpy())


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("param_dtype", ["fp32", "bf16"])
def test_against_reference_golden(golden, name, param_dtype):
    g = golden(name)
    m = build_model(g).cuda()
    if param_dtype == "bf16":
        m = m.bfloat16()
    got = _run_fixture(g, m)
    for key, expected in g.out.items():
        max_abs, min_cos = error_stats(got[key], expected)
        # unit-scale outputs take the stated bound as is; BERT's post-norm outputs reach |x| ~ 12 and logits are not
        # normalised at all, so those two kinds scale the bound with the output range (relative bf16 error)
        scaled = g.hyper["kind"] == "bert" or key == "logits"
        scale = max(1.0, float(np.abs(expected).max()) / 4.0) if scaled else 1.0
        assert max_abs <= MAX_ABS_12 * scale and min_cos >= MIN_COS, f"{name}.{key}: max_abs={max_abs} cos={min_cos}"


def test_c1_vit_ti16_batch8(golden):
    """BASELINE configs[0] against the reference's recorded fp32 output."""
    import pytorch_models_b200 as pm

    g = golden("c1_vit_ti16")
    h = g.hyper
    torch.manual_seed(h["weight_seed"])
    m = pm.ViT.from_google(h["tag"]).eval()
    oracle_torch.randomize_(m.state_dict(), h["noise_seed"])
    torch.manual_seed(h["input_seed"])
    x = torch.randn(h["batch"], 3, 224, 224)
    with torch.no_grad():
        out = m.cuda()(x.cuda()).float().cpu().numpy()
    max_abs, min_cos = error_stats(out, g.out["pooled"])
    assert max_abs <= MAX_ABS_12 and min_cos >= MIN_COS, (max_abs, min_cos)


def _config_case(make, n_heads, pool, batch, shape, kind, max_abs_tol):
    torch.manual_seed(0)
    m = make().eval()
    sd = oracle_torch.randomize_(m.state_dict(), 100)
    torch.manual_seed(1)
    x = torch.randn(batch, *shape)
    with torch.no_grad():
        if kind == "vit":
            want = oracle_torch.vit_forward(sd, x, n_heads, pool)
        else:
            want = oracle_torch.whisper_encoder_forward(sd, x)
        got = m.cuda().bfloat16()(x.cuda().bfloat16()).float().cpu()
    max_abs, min_cos = error_stats(got.numpy(), want.numpy())
    assert max_abs <= max_abs_tol and min_cos >= MIN_COS, (max_abs, min_cos)


def test_c2_vit_b16_224():
    import pytorch_models_b200 as pm

    _config_case(lambda: pm.ViT.from_google("B/16"), 12, "cls_token", 4, (3, 224, 224), "vit", MAX_ABS_12)


def test_c3_vit_l16_siglip_384():
    import pytorch_models_b200 as pm

    _config_case(lambda: pm.ViT.from_google("L/16_siglip", img_size=384), 16, "mha", 2, (3, 384, 384), "vit", MAX_ABS_32)


def test_c4_dinov2_l14_518():
    import pytorch_models_b200 as pm

    _config_case(lambda: pm.ViT.from_facebook("L/14_dinov2"), 16, "cls_token", 1, (3, 518, 518), "vit", MAX_ABS_32)


def test_c5_whisper_large_v3_encoder():
    import pytorch_models_b200 as pm

    _config_case(lambda: pm.WhisperEncoder(32, 1280, 128), 20, None, 1, (128, 3000), "whisper", MAX_ABS_32)


def test_batch_shard_is_bit_identical():
    """per-sample results must not depend on how the batch is sharded (no cross-sample math)."""
    import pytorch_models_b200 as pm

    torch.manual_seed(0)
    m = pm.ViT.from_google("Ti/16").eval()
    oracle_torch.randomize_(m.state_dict(), 100)
    m = m.cuda().bfloat16()
    x = torch.randn(8, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        whole = m(x)
        halves = torch.cat([m(x[:4]), m(x[4:])])
        singles = torch.cat([m(x[i:i + 1]) for i in range(8)])
    assert torch.equal(whole, halves) and torch.equal(whole, singles)


def test_resize_pe_and_inplace_weight_mutation(golden):
    g = golden("vit_cls")
    m = build_model(g).cuda()
    x = torch.from_numpy(np.array(g.input)).cuda()
    with torch.no_grad():
        y0 = m(x)
        # LayerScale-style in-place fold (vit.py:290-304) must invalidate the packed-weight cache
        m.layers[0].sa.out_proj.weight.mul_(0.5)
        m.layers[1].mlp_norm.bias.add_(0.25)
        y1 = m(x)
        fresh = build_model(g).cuda()
        fresh.load_state_dict(m.state_dict())
        y2 = fresh(x)
    assert not torch.equal(y0, y1) and torch.equal(y1, y2)
    with torch.no_grad():
        m.resize_pe(96)  # test_vit.py:21-26 (224 -> 256 there)
        assert m.pe.shape == (1, 36, 128)
        assert m(torch.randn(2, 3, 96, 96, device="cuda")).shape == (2, 128)
        with pytest.raises(ValueError):
            m(x)  # 64 px no longer matches pe


def test_encoder_accepts_leading_dims_and_strided_input():
    """Encoder.forward takes (*, L, d) with arbitrary strides (MobileViT passes 4-D, SigLIP a permuted view)."""
    import pytorch_models_b200 as pm

    torch.manual_seed(3)
    enc = pm.Encoder(2, 128).eval()
    sd = oracle_torch.randomize_(enc.state_dict(), 5)
    x = torch.randn(2, 3, 128, 40).transpose(-1, -2)  # (2, 3, 40, 128) non-contiguous
    with torch.no_grad():
        want = oracle_torch.encoder({f"layers.{k}": v for k, v in sd.items()}, x, 2, True, 1e-5)
        got = enc.cuda()(x.cuda())
    assert got.shape == x.shape and got.dtype == torch.float32
    max_abs, min_cos = error_stats(got.cpu().numpy().reshape(6, -1), want.numpy().reshape(6, -1))
    assert max_abs <= 0.06 and min_cos >= MIN_COS, (max_abs, min_cos)


def test_cuda_graph_replay_is_bit_identical(golden):
    """The whole forward is capturable (no host sync / allocation inside the library) and replays bit-exactly."""
    from pytorch_models_b200.graphs import GraphedForward

    g = golden("vit_cls")
    m = build_model(g).cuda().bfloat16()
    x = torch.from_numpy(np.array(g.input)).cuda().bfloat16()
    with torch.no_grad():
        eager = m(x)
        graphed = GraphedForward(m, x)
        y1 = graphed(x)
        x2 = torch.randn_like(x)
        y2 = graphed(x2)
        assert torch.equal(y1, eager) and torch.equal(y2, m(x2))
    with pytest.raises(ValueError):
        graphed(x[:1])


def test_decoder_generator_greedy_matches_oracle(golden):
    """DecoderGenerator (generator.py:16-39) on the GPT-2 fixture: greedy continuation of a prompt, token by token,
    against the fp32 oracle fed the same growing prefix. 1-D token input, as the reference's generator passes it."""
    from pytorch_models_b200.text import DecoderGenerator

    class Tok:  # stand-in tokenizer: "3 5 8" <-> [3, 5, 8]
        eos_token_id = -1

        def encode(self, s):
            return [int(t) for t in s.split()]

        def decode(self, ids):
            return " ".join(str(i) for i in ids)

    g = golden("gpt2")
    m = build_model(g).cuda()
    out = DecoderGenerator(m, Tok()).generate("5 17 300 2 41", max_tokens=6)
    got = [int(t) for t in out.split()]
    assert len(got) == 11 and got[:5] == [5, 17, 300, 2, 41]
    sd = g.torch_sd()
    want = got[:5]
    with torch.no_grad():
        for step in range(6):
            logits = oracle_torch.gpt2_forward(sd, torch.tensor(want))[-1]
            top2 = logits.topk(2).values
            nxt = int(logits.argmax())
            if nxt != got[5 + step]:
                # a bf16 near-tie may legitimately flip the argmax: accept only if the oracle's margin is tiny
                assert float(top2[0] - top2[1]) < 0.05 * float(logits.std()), (step, nxt, got[5 + step])
                nxt = got[5 + step]
            want.append(nxt)
    assert want == got


@pytest.mark.parametrize("name", AUDIO_FIXTURES)
def test_whisper_preprocessor_against_reference_golden(golden, name):
    """fp32 front end (STFT + mel + log + normalisation in one kernel): absolute tolerance 2e-4 on values in [-0.6, 1.5]
    (direct 400-point DFT vs the reference's FFT: different summation order, nothing else)."""
    g = golden(name)
    m = build_model(g).cuda()
    x = torch.from_numpy(np.array(g.input)).cuda()
    with torch.no_grad():
        got = m(x)
    assert got.dtype == torch.float32 and tuple(got.shape) == g.out["logmel"].shape
    np.testing.assert_allclose(got.cpu().numpy(), g.out["logmel"], rtol=0, atol=2e-4)
    # leading dims and strided input are accepted like any (*, L) tensor
    with torch.no_grad():
        again = m(x.unsqueeze(0))[0]
    assert torch.equal(again, got)


def test_whisper_audio_to_encoder_end_to_end():
    """Raw audio -> WhisperPreprocessor -> WhisperEncoder on the GPU against the fp32 oracle chain (30 s would be
    480000 samples; 2 s here)."""
    import pytorch_models_b200 as pm

    torch.manual_seed(0)
    pre, enc = pm.WhisperPreprocessor("tiny"), pm.WhisperEncoder(2, 128, 80).eval()
    sd = oracle_torch.randomize_(enc.state_dict(), 100)
    audio = 0.3 * torch.randn(2, 32000)
    with torch.no_grad():
        want = oracle_torch.whisper_encoder_forward(sd, oracle_torch.whisper_logmel(audio, pre.filters))
        got = enc.cuda()(pre.cuda()(audio.cuda())).float().cpu()
    max_abs, min_cos = error_stats(got.numpy(), want.numpy())
    assert max_abs <= MAX_ABS_12 and min_cos >= MIN_COS, (max_abs, min_cos)


def test_torch_compile_fullgraph_matches_eager(golden):
    """Reference guarantee (README.md:7, tests/*/test_*.py `torch.compile(m, fullgraph=True)`): the compiled model is
    one b200enc::module_forward node and returns exactly what the eager call returns."""
    with torch.no_grad():
        g = golden("vit_cls")
        m = build_model(g).cuda()
        x = torch.from_numpy(np.array(g.input)).cuda()
        want = m(x)
        for backend in ("aot_eager", "inductor"):
            torch._dynamo.reset()
            got = torch.compile(m, fullgraph=True, backend=backend)(x)
            assert torch.equal(got, want), backend
        torch._dynamo.reset()
        g = golden("whisper_full")
        w = build_model(g).cuda()
        audio = torch.from_numpy(np.array(g.input)).cuda()
        targets = torch.from_numpy(np.array(g.extra["targets"])).cuda()
        assert torch.equal(torch.compile(w, fullgraph=True, backend="aot_eager")(audio, targets), w(audio, targets))
        torch._dynamo.reset()
        g = golden("gpt2")
        lm = build_model(g).cuda().bfloat16()
        ids = torch.from_numpy(np.array(g.input)).cuda()
        assert torch.equal(torch.compile(lm, fullgraph=True, backend="aot_eager")(ids), lm(ids))
