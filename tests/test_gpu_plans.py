"""Launch plans on the GPU: a replayed forward (one b200enc_run_ops call) is bit-identical to the per-launch path, for
new inputs, after weight changes, for every pooling head / patch path, for Encoder and Decoder."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import build_model

import pytorch_models_b200 as pm
from pytorch_models_b200 import ops, plans

pytestmark = pytest.mark.gpu


def _per_launch(fn):
    prev = plans.enable(False)
    try:
        with torch.no_grad():
            return fn()
    finally:
        plans.enable(prev)


def _check_replay(m, make_input, n_rounds=4):
    """forward #1 runs plainly (a shape seen once is not worth a plan), #2 records, #3.. replay on fresh inputs; each
    must equal the per-launch path bit for bit."""
    plans.clear(m)
    before = dict(plans.STATS)
    with torch.no_grad():
        for r in range(n_rounds):
            x = make_input(r)
            got = m(*x)
            want = _per_launch(lambda: m(*x))
            assert got.shape == want.shape and got.dtype == want.dtype
            assert torch.equal(got, want), f"round {r}: replayed forward differs from the per-launch path"
    assert plans.STATS["recorded"] == before["recorded"] + 1
    assert plans.STATS["replayed"] == before["replayed"] + n_rounds - 2


@pytest.mark.parametrize("name", ["vit_cls", "vit_gap", "vit_siglip", "vit_p14"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vit_replay_matches_per_launch_path(golden, name, dtype):
    g = golden(name)
    m = build_model(g).cuda().to(dtype)
    shape = g.input.shape

    def make(r):
        torch.manual_seed(100 + r)
        return (torch.randn(*shape, device="cuda", dtype=dtype),)

    _check_replay(m, make)


def test_vit_replay_against_golden(golden):
    g = golden("vit_cls")
    m = build_model(g).cuda()
    x = torch.from_numpy(np.array(g.input)).cuda()
    plans.clear(m)
    with torch.no_grad():
        m(torch.randn_like(x))      # first sight
        m(torch.randn_like(x))      # records on other data
        got = m(x)                  # replay
    assert plans.STATS["replayed"] >= 1
    assert float((got.float().cpu() - torch.from_numpy(g.out["pooled"])).abs().max()) <= 0.06


def test_replay_is_one_library_call_and_counts_launches(golden):
    m = build_model(golden("vit_cls")).cuda().bfloat16()
    x = torch.randn(3, *golden("vit_cls").input.shape[1:], device="cuda", dtype=torch.bfloat16)
    plans.clear(m)
    with torch.no_grad():
        n0 = ops.LAUNCHES
        m(x)
        per_forward = ops.LAUNCHES - n0
        m(x)                                        # second sight: recorded
        n1 = ops.LAUNCHES
        m(x)
        assert ops.LAUNCHES - n1 == per_forward     # the replay launches (and counts) the same kernels
    plan = next(iter(plans.plans_of(m).values()))
    assert plan is not None and plan.n == per_forward


def test_weight_changes_invalidate_the_plan(golden):
    g = golden("vit_cls")
    m = build_model(g).cuda()
    x = torch.from_numpy(np.array(g.input)).cuda()
    plans.clear(m)
    with torch.no_grad():
        y0 = m(x)
        assert torch.equal(m(x), y0) and torch.equal(m(x), y0) and plans.STATS["replayed"] >= 1
        m.layers[1].mlp.linear2.weight.mul_(1.5)          # in place
        y1 = m(x)
        assert not torch.equal(y1, y0) and torch.equal(y1, _per_launch(lambda: m(x)))
        m.norm.weight.data.mul_(2.0)                      # through .data: invisible to PyTorch, documented
        pm.invalidate_packed()
        y2 = m(x)
        assert not torch.equal(y2, y1) and torch.equal(y2, _per_launch(lambda: m(x)))
        m = m.bfloat16()                                  # dtype move re-packs and re-records
        y3 = m(x.bfloat16())
        assert torch.equal(y3, _per_launch(lambda: m(x.bfloat16())))


def test_shapes_and_streams_get_their_own_plans(golden):
    g = golden("vit_cls")
    m = build_model(g).cuda().bfloat16()
    plans.clear(m)
    c, h, w = g.input.shape[1:]
    with torch.no_grad():
        for n in (1, 5, 1, 5, 1, 5):
            x = torch.randn(n, c, h, w, device="cuda", dtype=torch.bfloat16)
            assert torch.equal(m(x), _per_launch(lambda: m(x)))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            x = torch.randn(5, c, h, w, device="cuda", dtype=torch.bfloat16)
            y = m(x)
            y_again = m(x)
            y_third = m(x)
        side.synchronize()
        assert torch.equal(y, y_again) and torch.equal(y, y_third) and torch.equal(y, _per_launch(lambda: m(x)))
    assert len(plans.plans_of(m)) == 3


def test_outputs_of_successive_replays_do_not_alias(golden):
    g = golden("vit_cls")
    m = build_model(g).cuda().bfloat16()
    c, h, w = g.input.shape[1:]
    with torch.no_grad():
        xs = [torch.randn(4, c, h, w, device="cuda", dtype=torch.bfloat16) for _ in range(5)]
        ys = [m(x) for x in xs]
        torch.cuda.synchronize()
        for x, y in zip(xs, ys):
            assert torch.equal(y, _per_launch(lambda: m(x)))
    assert len({y.data_ptr() for y in ys}) == 5


def test_encoder_and_decoder_replay():
    torch.manual_seed(0)
    enc = pm.Encoder(3, 128).eval().cuda()
    _check_replay(enc, lambda r: (torch.randn(2, 77, 128, device="cuda", generator=None),))
    dec = pm.Decoder(2, 128, cross_attn=True).eval().cuda()

    def make(r):
        torch.manual_seed(r)
        return torch.randn(2, 33, 128, device="cuda"), torch.randn(2, 50, 128, device="cuda")

    _check_replay(dec, make)
    post = pm.Encoder(2, 128, pre_norm=False).eval().cuda().bfloat16()
    _check_replay(post, lambda r: (torch.randn(3, 40, 128, device="cuda", dtype=torch.bfloat16),))


def test_whisper_encoder_and_bert_replay(golden):
    g = golden("whisper")
    m = build_model(g).cuda()
    shape = g.input.shape

    def mel(r):
        torch.manual_seed(50 + r)
        return (torch.randn(*shape, device="cuda"),)

    _check_replay(m, mel)
    _check_replay(m.bfloat16(), lambda r: (mel(r)[0].bfloat16(),))
    g = golden("bert")
    b = build_model(g).cuda()
    ids0 = torch.from_numpy(np.array(g.input)).cuda()
    vocab = int(ids0.max().item()) + 1

    def ids(r):
        torch.manual_seed(70 + r)
        return (torch.randint(0, vocab, tuple(ids0.shape), device="cuda"),)

    _check_replay(b, ids)
    with torch.no_grad():
        got = b(ids0).float().cpu().numpy()      # replay on the fixture's ids
    want = g.out["tokens"]
    assert float(np.abs(got - want).max()) <= 0.125 * max(1.0, float(np.abs(want).max()) / 4.0)


def test_modules_with_plans_can_be_copied_pickled_and_collected(golden):
    """Plans live outside the module: deepcopy (EMA copies) and pickling (torch.save(model)) keep working after a
    forward, the copy records its own plan, and a dropped module releases its plan (and the workspaces it owns)."""
    import copy
    import gc
    import io
    import weakref

    g = golden("vit_cls")
    m = build_model(g).cuda().bfloat16()
    x = torch.randn(2, *g.input.shape[1:], device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        y = m(x)
        assert torch.equal(m(x), y) and torch.equal(m(x), y) and len(plans.plans_of(m)) == 1
        twin = copy.deepcopy(m)
        assert len(plans.plans_of(twin)) == 0
        assert torch.equal(twin(x), y) and torch.equal(twin(x), y) and torch.equal(twin(x), y)
        buf = io.BytesIO()
        torch.save(m, buf)
        buf.seek(0)
        loaded = torch.load(buf, weights_only=False)
        assert torch.equal(loaded(x), y)
    ref = weakref.ref(m)
    plan_ref = weakref.ref(next(iter(plans.plans_of(m).values())))
    del m
    gc.collect()
    assert ref() is None and plan_ref() is None


def test_one_off_shapes_are_not_recorded():
    """A generation-style loop (the sequence grows by one token per call, text/generator.py) never sees a shape twice:
    no plan is recorded, nothing is kept alive."""
    torch.manual_seed(0)
    dec = pm.Decoder(2, 128).eval().cuda()
    before = dict(plans.STATS)
    with torch.no_grad():
        for L in range(3, 12):
            x = torch.randn(1, L, 128, device="cuda")
            assert torch.equal(dec(x), _per_launch(lambda: dec(x)))
    assert plans.STATS["recorded"] == before["recorded"] and plans.plans_of(dec) == {}


def test_profiling_bypasses_plans(golden):
    m = build_model(golden("vit_cls")).cuda().bfloat16()
    x = torch.randn(2, *golden("vit_cls").input.shape[1:], device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        m(x), m(x), m(x)
        rec = ops.profile(True)
        try:
            m(x)
        finally:
            ops.profile(False)
    assert len(rec) > 0 and all(name.startswith("b200enc_") for name, *_ in rec)


def test_bad_input_still_raises_through_a_plan(golden):
    m = build_model(golden("vit_cls")).cuda()
    c, h, w = golden("vit_cls").input.shape[1:]
    with pytest.raises(ValueError):
        with torch.no_grad():
            m(torch.randn(2, c, h + 16, w, device="cuda"))
