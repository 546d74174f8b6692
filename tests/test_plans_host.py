"""Host-side checks of the launch-plan machinery (plans.py, b200enc_run_ops): struct layout against the C header,
module signatures, argument-free calls into the library. No GPU needed."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

import pytorch_models_b200 as pm
from pytorch_models_b200 import _lib, plans


def test_op_struct_layout_matches_header(tmp_path):
    """ctypes mirrors of b200enc_op / b200enc_linear_args have the size and field offsets gcc gives the header."""
    fields_op = ["kind", "linear", "p", "i", "f"]
    fields_lin = [name for name, _ in _lib.LinearArgs._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200enc.h"', "int main(void) {",
            '  printf("%zu %zu\\n", sizeof(b200enc_op), sizeof(b200enc_linear_args));']
    c_names = dict(rowstats_parts="rowstats_parts")
    for f in fields_op:
        prog.append(f'  printf("%zu\\n", offsetof(b200enc_op, {f}));')
    for f in fields_lin:
        prog.append(f'  printf("%zu\\n", offsetof(b200enc_linear_args, {c_names.get(f, f)}));')
    prog += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    nums = [int(v) for v in out]
    assert nums[0] == ctypes.sizeof(_lib.Op) and nums[1] == ctypes.sizeof(_lib.LinearArgs)
    want = [getattr(_lib.Op, f).offset for f in fields_op] + [getattr(_lib.LinearArgs, f).offset for f in fields_lin]
    assert nums[2:] == want


def test_op_kinds_match_header():
    text = open(os.path.join(ROOT, "include", "b200enc.h")).read()
    for name, kind in _lib.OP_KINDS.items():
        macro = "B200ENC_OP_" + name[len("b200enc_"):].upper()
        assert f"#define {macro} {kind}" in text, macro
        assert name in _lib._SIGNATURES


def test_run_ops_argument_errors():
    """Pure host paths of b200enc_run_ops: an empty plan is a no-op, an unknown kind is an argument error that names
    the failing element (no kernel is launched, so this runs without a GPU)."""
    lib = _lib.load()
    failed = ctypes.c_int(7)
    assert lib.b200enc_run_ops(None, 0, ctypes.byref(failed), None) == 0 and failed.value == -1
    ops = (_lib.Op * 2)()
    ops[0].kind = 99
    rc = lib.b200enc_run_ops(ops, 2, ctypes.byref(failed), None)
    assert rc == -1 and failed.value == 0
    assert b"unknown op kind 99" in lib.b200enc_last_error()
    assert lib.b200enc_run_ops(None, 3, None, None) == -1


def test_signature_sees_every_kind_of_weight_change():
    m = pm.ViT(2, 64, 1, 16, img_size=32).eval()
    sig = plans._Signature(m)
    assert sig.valid()
    with torch.no_grad():
        m.layers[0].sa.q_proj.weight.mul_(2.0)           # in place (what the reference loaders do)
    assert not sig.valid()
    sig = plans._Signature(m)
    m.load_state_dict(m.state_dict())                     # copy_ into every parameter
    assert not sig.valid()
    sig = plans._Signature(m)
    m.resize_pe(64)                                       # replaces the Parameter object
    assert not sig.valid()
    sig = plans._Signature(m)
    m.norm.weight.data = m.norm.weight.data.clone()       # moved storage (what .to(device) does)
    assert not sig.valid()
    sig = plans._Signature(m)
    m.layers[1] = pm.EncoderLayer(64, 1)                  # module surgery
    assert not sig.valid()
    sig = plans._Signature(m)
    m.train()
    assert not sig.valid()
    sig = plans._Signature(m.eval())
    m.register_buffer("extra", torch.zeros(1))
    assert not sig.valid()


def test_cpu_tensors_still_raise_with_plans_enabled():
    m = pm.ViT(1, 64, 1, 16, img_size=32).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 32))
    enc = pm.Encoder(1, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(1, 4, 64))


def test_plan_cache_lives_outside_the_module():
    """deepcopy / pickle of a module must never meet a plan (ctypes arrays, a lock): the cache is a weak map keyed by
    the module, and a signature does not keep its module alive."""
    import copy
    import gc
    import weakref

    m = pm.ViT(1, 64, 1, 16, img_size=32).eval()
    plans._PLANS[m] = {"k": object()}
    sig = plans._Signature(m)
    twin = copy.deepcopy(m)
    assert plans.plans_of(twin) == {} and "_b200_plans" not in m.__dict__
    ref = weakref.ref(m)
    del m
    gc.collect()
    assert ref() is None and not sig.valid()


def test_plans_module_never_imports_the_oracle():
    src = open(os.path.join(ROOT, "pytorch_models_b200", "plans.py")).read()
    assert "oracle" not in src


def test_run_ops_unpacks_arguments_in_declaration_order():
    """Every kind reaches its own entry point, and the positional slots land on the right parameters: the argument
    errors (raised before anything touches the GPU) echo the values that were put into i[]."""
    lib = _lib.load()
    failed = ctypes.c_int(-1)
    ops = (_lib.Op * 1)()
    for name, kind in _lib.OP_KINDS.items():          # zeroed op: each entry point rejects its own null pointers
        ops[0] = _lib.Op()
        ops[0].kind = kind
        assert lib.b200enc_run_ops(ops, 1, ctypes.byref(failed), None) == -1 and failed.value == 0
        assert lib.b200enc_last_error().decode().startswith(name + ":"), (name, lib.b200enc_last_error())
    fake = 0x10000
    op = _lib.Op()
    op.kind = _lib.OP_KINDS["b200enc_attention"]
    for j in range(4):
        op.p[j] = fake
    op.i[:12] = (300, 192, 300, 192, 100, 64, 5, 6, 7, 8, 32, 0)   # ..., B, H, Lq, Lkv, head_dim, flags
    ops[0] = op
    assert lib.b200enc_run_ops(ops, 1, ctypes.byref(failed), None) == -1
    assert b"head_dim=32" in lib.b200enc_last_error()
    op.i[10], op.i[6] = 64, 0
    ops[0] = op
    assert lib.b200enc_run_ops(ops, 1, ctypes.byref(failed), None) == -1
    assert b"B=0 H=6 Lq=7 Lkv=8" in lib.b200enc_last_error()
    op = _lib.Op()
    op.kind = _lib.OP_KINDS["b200enc_patch_embed16"]
    op.linear.x = op.linear.w = op.linear.out = op.linear.residual = fake
    op.i[0], op.i[1] = 30, 224
    ops[0] = op
    assert lib.b200enc_run_ops(ops, 1, ctypes.byref(failed), None) == -1
    assert b"image 30 x 224" in lib.b200enc_last_error()
    two = (_lib.Op * 2)()
    two[0].kind = 99
    two[1] = op
    assert lib.b200enc_run_ops(two, 2, ctypes.byref(failed), None) == -1 and failed.value == 0   # stops at the first failure


def test_launch_plan_packs_a_recording_and_finds_the_patch_slots():
    """LaunchPlan from a hand-made recording (CPU tensors stand in for device buffers; nothing is launched): kinds,
    argument slots, which pointers are patch slots of the input / the output, and what the plan keeps alive."""
    x = torch.zeros(4, 8, 64, dtype=torch.bfloat16)           # the caller's input
    out = torch.zeros(4, 8, 64, dtype=torch.bfloat16)         # the forward's fresh output
    ws = torch.zeros(4 * 8, 192, dtype=torch.bfloat16)        # a workspace the plan must own
    w = torch.zeros(192, 64, dtype=torch.bfloat16)
    gamma = torch.ones(64)
    rec = plans._Recorder()
    la = _lib.LinearArgs(x.data_ptr(), 0, 64, w.data_ptr(), 64, None, None, None, 0, 0.0, x.data_ptr() + 128, 0, 64,
                         ws.data_ptr(), 0, 192, None, 1, 32, 192, 64, 0, 0, 0, None)
    rec.calls.append(("b200enc_linear", (ctypes.byref(la),)))
    rec.calls.append(("b200enc_layernorm", (ws.data_ptr(), 192, gamma.data_ptr(), gamma.data_ptr(), 1e-6, 32, 64,
                                            out.data_ptr() + 64, 64, None)))
    rec.keep.extend([x, w, ws, gamma, out, ws])
    m = pm.Encoder(1, 64).eval()
    plan = plans.LaunchPlan(rec, (x,), out, plans._Signature(m))
    assert plan.n == 2 and [o.kind for o in plan.ops] == [1, 5]
    assert plan.ops[0].linear.N == 192 and plan.ops[0].linear.out == ws.data_ptr()
    ln = plan.ops[1]
    assert ln.p[0] == ws.data_ptr() and ln.p[3] == out.data_ptr() + 64 and ln.p[4] is None
    assert list(ln.i[:4]) == [192, 32, 64, 64] and abs(ln.f[0] - 1e-6) < 1e-12
    in_slots, out_slots = plan.patches
    assert sorted((f, off) for _, f, off in in_slots) == [("residual", 128), ("x", 0)]
    assert [(f, off) for _, f, off in out_slots] == [(3, 64)]
    kept = {t.data_ptr() for t in plan.keep}
    assert kept == {w.data_ptr(), ws.data_ptr(), gamma.data_ptr()}            # not the caller's input / output
    # patching writes through to the op array
    for view, field, off in in_slots:
        setattr(view, field, 0x5000 + off)
    assert plan.ops[0].linear.x == 0x5000 and plan.ops[0].linear.residual == 0x5000 + 128
    with pytest.raises(RuntimeError, match="no recorded launch writes the output"):
        plans.LaunchPlan(rec, (x,), torch.zeros(3), plans._Signature(m))
    with pytest.raises(RuntimeError, match="overlap"):     # decoder(x, memory=x): pointers could not be attributed
        plans.LaunchPlan(rec, (x, x[1:]), out, plans._Signature(m))
