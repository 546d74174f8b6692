"""``torch.compile(model, fullgraph=True)`` support (the reference promises and tests it: README.md:7,
tests/image/test_vit.py:14-18, tests/text/test_gpt2.py:22, ...).

The arithmetic of this package runs in libb200enc.so through ctypes — nothing Dynamo can or should trace. A model's
whole forward is therefore ONE opaque custom operator, ``b200enc::module_forward``: under compilation ``forward`` calls
the operator (the module is registered under an integer id at construction; its parameters / buffers are passed as
inputs so that the graph depends on them), the operator's real implementation runs the ordinary eager forward, and its fake implementation
only describes the output's shape and dtype. Eager calls never touch any of this.
"""
from __future__ import annotations

import functools
import weakref

import torch
from torch import Tensor

_MODULES: "weakref.WeakValueDictionary[int, torch.nn.Module]" = weakref.WeakValueDictionary()
_NEXT_ID = [1]


def register(module: torch.nn.Module) -> int:
    """Give ``module`` its own id in the registry the operator looks modules up in. Runs from ``__init__`` of every
    compilable class; call it by hand for a module obtained some other way (``copy.deepcopy``, unpickling) before
    compiling it."""
    mid = _NEXT_ID[0]
    _NEXT_ID[0] += 1
    module.__dict__["_b200_module_id"] = mid
    _MODULES[mid] = module
    return mid


def compilable_module(cls):
    """Class decorator: register every instance right after construction (outside any Dynamo trace)."""
    orig = cls.__init__

    @functools.wraps(orig)
    def init(self, *args, **kwargs):
        orig(self, *args, **kwargs)
        register(self)

    cls.__init__ = init
    return cls


def _lookup(module_id: int) -> torch.nn.Module:
    module = _MODULES.get(module_id)
    if module is None:
        raise RuntimeError(f"b200enc::module_forward: module {module_id} no longer exists")
    return module


@torch.library.custom_op("b200enc::module_forward", mutates_args=())
def module_forward(x: Tensor, extra: Tensor | None, module_id: int, state: list[Tensor]) -> Tensor:
    module = _lookup(module_id)
    first = next(iter(module.parameters()), None)
    if first is not None and (not state or state[0].data_ptr() != first.data_ptr()):
        raise RuntimeError("b200enc::module_forward: the compiled graph belongs to another module instance "
                           "(deep copy?): call pytorch_models_b200.compile.register(module) before torch.compile")
    eager = type(module).forward._eager
    out = eager(module, x) if extra is None else eager(module, x, extra)
    return out.clone() if out.data_ptr() == x.data_ptr() else out  # an operator may not return an alias of its input


@module_forward.register_fake
def _(x, extra, module_id, state):
    module = _lookup(module_id)
    shape, dtype = type(module).forward._out_meta(module, x, extra)
    return x.new_empty(tuple(shape), dtype=dtype)


def compilable(out_meta):
    """Decorator for a model's ``forward(self, x[, extra])``: eager calls go straight through; while Dynamo is
    tracing, the call becomes one ``b200enc::module_forward`` node. ``out_meta(self, x, extra) -> (shape, dtype)``."""

    def deco(fwd):
        @functools.wraps(fwd)
        def wrapper(self, x, extra=None):
            if torch.compiler.is_compiling():
                state = list(self.parameters()) + list(self.buffers())
                return torch.ops.b200enc.module_forward(x, extra, self._b200_module_id, state)
            return fwd(self, x) if extra is None else fwd(self, x, extra)

        wrapper._eager = fwd
        wrapper._out_meta = out_meta
        return wrapper

    return deco


def float_like(x: Tensor) -> torch.dtype:
    """Output dtype rule of the image / audio models: bf16 and fp32 inputs keep their dtype, everything else is fp32."""
    return x.dtype if x.dtype in (torch.bfloat16, torch.float32) else torch.float32
