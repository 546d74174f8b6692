"""GPU tests (`-m gpu`) for the corners round 1 left open (VERDICT r01 "What's weak" 1):

* attention with an additive bias whose FIRST key block is fully masked for some rows (sliding-window, left padding),
  key-broadcast biases, cross-attention layers called without a memory;
* LayerNorm folded into the GEMMs under trained-like statistics — row mean >> row std, a few massive channels, widely
  spread gamma — for the folded QKV / FC1 GEMMs, the fused-statistics chain and a 12-layer stack, against the fp32
  oracle and against PyTorch's own bf16 execution of the same math on the same GPU;
* the host-side caches and the device-side status word.

Tolerances are stated per test: "no worse than 1.5x what PyTorch's bf16 kernels show on the same inputs, plus a small
absolute floor" — a bound that scales with the problem instead of a fixed number tuned to unit-variance inputs.
"""
from __future__ import annotations

import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import error_stats
from oracle import oracle_torch

pytestmark = pytest.mark.gpu


def _close(got, want, atol, rtol):
    got, want = got.float(), want.float()
    bad = (got - want).abs() > atol + rtol * want.abs()
    assert not bool(bad.any()), f"{int(bad.sum())} / {bad.numel()} mismatches, max abs {(got - want).abs().max().item():.4g}"


# ------------------------------------------------------------------------------------------- attention masks
@pytest.mark.parametrize("kind", ["sliding_window", "left_padding", "block_diagonal"])
@pytest.mark.parametrize("L", [300, 700])
def test_attention_bias_first_blocks_fully_masked(kind, L):
    """Rows whose first 128-key block(s) hold no visible key (these returned NaN). SDPA gives finite rows."""
    from pytorch_models_b200 import ops

    torch.manual_seed(L)
    B, H = 2, 2
    qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
    i = torch.arange(L, device="cuda")
    if kind == "sliding_window":
        vis = (i[:, None] - i[None, :]).abs() <= 40
        vis = vis.expand(B, 1, L, L)
    elif kind == "left_padding":
        pad = torch.tensor([0, L - 90], device="cuda")            # sample 1: only the last 90 keys are real
        vis = (i[None, :] >= pad[:, None])[:, None, None, :].expand(B, 1, L, L)
    else:
        blk = i // 150
        vis = (blk[:, None] == blk[None, :]).expand(B, 1, L, L)
    bias = torch.zeros(B, 1, L, L, device="cuda").masked_fill(~vis, float("-inf"))
    bias = bias + 0.5 * torch.randn(B, 1, L, L, device="cuda")
    out = torch.empty(B, L, H * 64, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[:, :, : H * 64], qkv[:, :, H * 64: 2 * H * 64], qkv[:, :, 2 * H * 64:]
    ops.attention(q, k, v, out, H, 0.125, bias=bias.expand(B, H, L, L))
    heads = lambda t: t.float().unflatten(-1, (H, 64)).transpose(1, 2)  # noqa: E731
    want = F.scaled_dot_product_attention(heads(q), heads(k), heads(v), attn_mask=bias).transpose(1, 2).flatten(-2)
    assert bool(torch.isfinite(out).all())
    _close(out, want, 0.03, 0.03)


def test_attention_key_broadcast_bias():
    """attn_bias of shape (..., Lq, 1) — one value per query row, broadcast over the keys."""
    import pytorch_models_b200 as pm

    torch.manual_seed(0)
    B, L, d = 2, 150, 128
    mha = pm.MHA(d).eval().cuda()
    x = torch.randn(B, L, d, device="cuda")
    bias = torch.randn(L, 1, device="cuda")
    with torch.no_grad():
        got = mha(x, attn_bias=bias)
        want = mha(x)  # a per-row constant does not change the softmax
    _close(got, want, 0.02, 0.02)


@pytest.mark.parametrize("pre_norm", [True, False])
def test_cross_attention_layer_without_memory(pre_norm):
    """Decoder(cross_attn=True)(x, None): the reference's MHA takes k = v = q when no memory is given
    (transformer.py:44-45), i.e. un-masked self-attention with the ca weights (this raised)."""
    import pytorch_models_b200 as pm

    torch.manual_seed(5)
    dec = pm.Decoder(2, 128, cross_attn=True, pre_norm=pre_norm).eval()
    sd = oracle_torch.randomize_(dec.state_dict(), 9)
    x = torch.randn(3, 70, 128)
    with torch.no_grad():
        want = oracle_torch.decoder(sd, x, None, 2, pre_norm, 1e-5, prefix="")  # Decoder is an nn.ModuleList: keys "0.sa..."
        got = dec.cuda()(x.cuda()).float().cpu()
        one = dec[0](x.cuda()).float().cpu()
        want_one = oracle_torch.decoder_layer(sd, "0.", x, None, 2, pre_norm, 1e-5)
    for g, w in ((got, want), (one, want_one)):
        max_abs, min_cos = error_stats(g.numpy(), w.numpy())
        assert max_abs <= 0.08 and min_cos >= 0.9999, (max_abs, min_cos)


# ------------------------------------------------------------------------------------------- trained-like statistics
def _trained_like(rows: int, d: int, kind: str, seed: int) -> torch.Tensor:
    """Residual-stream rows with the statistics trained ViT / DINOv2 streams show (fp32, CPU)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, d, generator=g)
    if kind == "mean50":
        x = x + 50.0                                        # row mean 50, row std 1
    elif kind == "outliers":
        idx = torch.randperm(d, generator=g)[:4]
        x[:, idx] = torch.tensor([500.0, -500.0, 350.0, -420.0]) + 5.0 * torch.randn(rows, 4, generator=g)
    elif kind == "first_outlier":
        x[:, 0::128] = 500.0 + torch.randn(rows, len(range(0, d, 128)), generator=g)  # the shift element of every slice
    elif kind == "mixed":
        x = 3.0 * x + 20.0
        idx = torch.randperm(d, generator=g)[:3]
        x[:, idx] += torch.tensor([300.0, -250.0, 600.0])
    return x


def _wide_gamma(d: int, seed: int) -> tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    gamma = torch.exp(1.5 * torch.randn(d, generator=g))   # spread over ~2 decades
    gamma[torch.randperm(d, generator=g)[:8]] = 0.01
    beta = torch.randn(d, generator=g)
    return gamma, beta


@pytest.mark.parametrize("kind", ["mean50", "outliers", "first_outlier", "mixed"])
@pytest.mark.parametrize("gelu", [False, True])
def test_folded_layernorm_gemm_trained_like(kind, gelu):
    """LN(x) W^T + b with the LayerNorm folded into the GEMM epilogue (QKV: plain; FC1: erf-GELU) on rows with a large
    common mean / massive channels and a wide gamma, row statistics from the row_stats kernel."""
    from types import SimpleNamespace

    from pytorch_models_b200 import ops
    from pytorch_models_b200.transformer import pack_folded

    M, d, N = 384, 768, 2304
    x = _trained_like(M, d, kind, 11).bfloat16()            # the stream is stored in bf16: quantise BEFORE the oracle
    gamma, beta = _wide_gamma(d, 12)
    torch.manual_seed(13)
    lin = torch.nn.Linear(d, N)
    norm = torch.nn.LayerNorm(d, 1e-6)
    with torch.no_grad():
        norm.weight.copy_(gamma), norm.bias.copy_(beta)
        want = F.linear(F.layer_norm(x.float(), (d,), gamma, beta, 1e-6), lin.weight, lin.bias)
        want = F.gelu(want) if gelu else want
        # what PyTorch's own bf16 kernels give for the same math on this GPU (LN -> round -> GEMM -> round [-> GELU])
        xb = x.cuda()
        lib = F.linear(F.layer_norm(xb, (d,), gamma.cuda().bfloat16(), beta.cuda().bfloat16(), 1e-6),
                       lin.weight.cuda().bfloat16(), lin.bias.cuda().bfloat16())
        lib = (F.gelu(lib) if gelu else lib).float().cpu()
        pk = pack_folded([lin.cuda()], norm.cuda())
        stats = torch.empty(M, 2, device="cuda", dtype=torch.float32)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.row_stats(xb, 1e-6, stats)
        ops.linear(xb, pk.w, pk.bias, out, colsum=pk.colsum, rowstats=stats, gelu=gelu)
    got = out.float().cpu()
    err_ours = (got - want).abs()
    err_lib = (lib - want).abs()
    scale = want.abs().max().item()
    # bound: 1.5x the library's own worst error on these inputs + 0.4 % of the output range (two bf16 roundings)
    bound = 1.5 * err_lib.max().item() + 0.004 * scale
    assert err_ours.max().item() <= bound, (kind, gelu, err_ours.max().item(), err_lib.max().item(), scale)
    assert err_ours.mean().item() <= 1.5 * err_lib.mean().item() + 1e-3 * scale


@pytest.mark.parametrize("kind", ["mean50", "outliers", "first_outlier", "mixed"])
def test_fused_statistics_chain_trained_like(kind):
    """residual GEMM -> partial (mean, M2) per 128 columns in its epilogue -> consumed by the next folded GEMM, when the
    rows the statistics describe have a large mean / massive channels (the shifted sums must not cancel)."""
    from pytorch_models_b200 import ops
    from pytorch_models_b200.transformer import pack_folded, pack_plain

    M, d, N = 300, 768, 1536
    torch.manual_seed(21)
    res = _trained_like(M, d, kind, 22).bfloat16()          # the residual stream carries the statistics' difficulty
    a = torch.randn(M, d).bfloat16()
    lin0, lin1, norm = torch.nn.Linear(d, d), torch.nn.Linear(d, N), torch.nn.LayerNorm(d, 1e-6)
    gamma, beta = _wide_gamma(d, 23)
    with torch.no_grad():
        norm.weight.copy_(gamma), norm.bias.copy_(beta)
        y = (F.linear(a.float(), lin0.weight.bfloat16().float(), lin0.bias) + res.float()).bfloat16()  # as stored
        want = F.linear(F.layer_norm(y.float(), (d,), gamma, beta, 1e-6), lin1.weight, lin1.bias)
        yb = y.cuda()
        lib = F.linear(F.layer_norm(yb, (d,), gamma.cuda().bfloat16(), beta.cuda().bfloat16(), 1e-6),
                       lin1.weight.cuda().bfloat16(), lin1.bias.cuda().bfloat16()).float().cpu()
        p0, p1 = pack_plain([lin0.cuda()]), pack_folded([lin1.cuda()], norm.cuda())
        mid = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
        parts = torch.empty(M, d // 128, 2, device="cuda", dtype=torch.float32)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.linear(a.cuda(), p0.w, p0.bias, mid, residual=res.cuda(), stats_out=parts)
        ops.linear(mid, p1.w, p1.bias, out, colsum=p1.colsum, rowstats=parts, ln_eps=1e-6)
        # the statistics themselves, against fp64 on the stored rows
        m64 = mid.double().cpu()
        mean = m64.mean(1)
        var = m64.var(1, unbiased=False)
        pm_ = parts.double().cpu()
        cnt = 128.0
        gmean = pm_[:, :, 0].mean(1)
        gm2 = pm_[:, :, 1].sum(1) + (cnt * (pm_[:, :, 0] - gmean[:, None]) ** 2).sum(1)
    np.testing.assert_allclose(gmean.numpy(), mean.numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose((gm2 / d).numpy(), var.numpy(), rtol=2e-4, atol=1e-5)
    got = out.float().cpu()
    # `want` was computed from y (CPU rounding of the first GEMM); compare on rows where mid == y bit for bit would be
    # too strict, so the bound carries the library's error on y plus one bf16 ulp of the stream propagated through LN
    scale = want.abs().max().item()
    bound = 1.5 * (lib - want).abs().max().item() + 0.01 * scale
    assert (got - want).abs().max().item() <= bound, (kind, (got - want).abs().max().item(), bound)


@pytest.mark.parametrize("kind", ["mean50", "mixed"])
def test_twelve_layer_stack_trained_like(kind):
    """A 12-layer pre-norm Encoder whose input stream has trained-like statistics and whose LayerNorms have a wide
    gamma: ours vs the fp32 oracle, calibrated by PyTorch's bf16 execution of the same stack on the same GPU."""
    import pytorch_models_b200 as pm

    torch.manual_seed(31)
    L, d, heads = 197, 256, 4
    enc = pm.Encoder(12, d, n_heads=heads).eval()
    sd = enc.state_dict()
    for k, v in sd.items():
        if "norm.weight" in k:
            v.copy_(_wide_gamma(d, hash(k) % 1000)[0].clamp(0.05, 8.0))
        elif "norm.bias" in k:
            v.copy_(0.3 * torch.randn(d))
        elif k.endswith("out_proj.weight") or k.endswith("linear2.weight"):
            v.mul_(0.3)  # keep the stream's statistics dominated by the input, like LayerScale-d trained nets
    x = _trained_like(4 * L, d, kind, 33).view(4, L, d).bfloat16()
    with torch.no_grad():
        want = oracle_torch.encoder(sd, x.float(), heads, True, 1e-5, prefix="")  # Encoder is an nn.Sequential: "0.sa..."
        sd_b = {k: v.cuda().bfloat16() for k, v in sd.items()}
        lib = oracle_torch.encoder(sd_b, x.cuda(), heads, True, 1e-5, prefix="").float().cpu()
        got = enc.cuda()(x.cuda()).float().cpu()
    # compare what a consumer sees: the stream after a final LayerNorm (unit scale)
    ln = lambda t: F.layer_norm(t, (d,))  # noqa: E731
    e_ours, cos_ours = error_stats(ln(got).numpy(), ln(want).numpy())
    e_lib, cos_lib = error_stats(ln(lib).numpy(), ln(want).numpy())
    # the angle to the fp32 result may exceed the bf16 library's by a tenth (at mean 50 the bf16 stream itself is the error:
    # both land at cosine 0.934, a different rounding order apart)
    assert e_ours <= 1.5 * e_lib + 0.02 and (1 - cos_ours) <= 1.1 * (1 - cos_lib) + 1e-4, (kind, e_ours, e_lib, cos_ours, cos_lib)


# ------------------------------------------------------------------------------------------- host-side behaviour
def test_invalidate_packed_after_write_through_data(golden):
    """Writes through ``param.data`` leave no trace PyTorch could report; `invalidate_packed` is the
    documented way to make the next forward re-pack."""
    from conftest import build_model
    from pytorch_models_b200.transformer import invalidate_packed

    g = golden("vit_cls")
    m = build_model(g).cuda()
    x = torch.from_numpy(np.array(g.input)).cuda()
    with torch.no_grad():
        y0 = m(x)
        m.layers[0].mlp.linear1.weight.data.mul_(0.5)   # e.g. an EMA swap
        invalidate_packed()
        y1 = m(x)
        fresh = build_model(g).cuda()
        fresh.load_state_dict(m.state_dict())
        y2 = fresh(x)
    assert not torch.equal(y0, y1) and torch.equal(y1, y2)


def test_parameters_created_under_inference_mode(golden):
    """Inference tensors do not track a version counter; building the cache key must not raise."""
    import pytorch_models_b200 as pm

    with torch.inference_mode():
        enc = pm.Encoder(1, 128).eval().cuda()
        y = enc(torch.randn(2, 10, 128, device="cuda"))
    assert y.shape == (2, 10, 128)


def test_tensor_map_cache_and_status_word(golden):
    from conftest import build_model
    from pytorch_models_b200 import _lib

    lib = _lib.load()
    g = golden("vit_cls")
    m = build_model(g).cuda().bfloat16()
    x = torch.from_numpy(np.array(g.input)).cuda().bfloat16()
    h, ms = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    with torch.no_grad():
        m(x)
        torch.cuda.synchronize()
        for _ in range(3):
            m(x)  # the caching allocator hands the same blocks back: every map of these forwards is a cache hit
        lib.b200enc_tensor_map_cache_stats(ctypes.byref(h), ctypes.byref(ms))
        h0, m0 = h.value, ms.value
        m(x)
        lib.b200enc_tensor_map_cache_stats(ctypes.byref(h), ctypes.byref(ms))
    assert h.value > h0 and ms.value == m0, (h0, m0, h.value, ms.value)
    torch.cuda.synchronize()
    assert lib.b200enc_async_status(0) == 0  # no kernel ever gave up on a barrier wait


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_a_non_current_device():
    """model.to('cuda:1') without torch.cuda.set_device(1): the launch must follow the tensors."""
    import pytorch_models_b200 as pm

    torch.manual_seed(0)
    enc = pm.Encoder(2, 128).eval()
    x = torch.randn(2, 50, 128)
    with torch.no_grad():
        y0 = enc.to("cuda:0")(x.to("cuda:0")).cpu()
        assert torch.cuda.current_device() == 0
        y1 = enc.to("cuda:1")(x.to("cuda:1")).cpu()
    assert torch.equal(y0, y1)
    from pytorch_models_b200 import ops

    with pytest.raises(RuntimeError):
        ops.row_stats(torch.zeros(4, 64, device="cuda:0", dtype=torch.bfloat16), 1e-5,
                      torch.zeros(4, 2, device="cuda:1"))
