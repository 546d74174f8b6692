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
