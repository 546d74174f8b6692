"""Launch plans: a model's forward recorded once and replayed through ONE C-ABI call (``b200enc_run_ops``).

The reference's forward is a Python loop over modules (``nn.Sequential``, transformer.py:133-149); the drop-in modules
walk the same loop and make one ctypes call per kernel, ~29 us of interpreter time each — 1.9 ms for the 65 launches
of a ViT-B/16 forward, against 4.5 ms of GPU time at the 128-image shard of the 8-GPU run. A forward whose body
consists of `ops.*` calls only is therefore recorded the SECOND time it runs for a given (input shapes, stream,
weights) — shapes that occur once (a generation loop over a growing sequence) are not worth a plan: every argument of every launch goes into an array of ``b200enc_op``, the tensors the launches touch
(workspaces, packed weights) are kept alive by the plan, and the pointers that fall inside the inputs / the output are
remembered as patch slots. Later forwards allocate a fresh output, patch those slots and enqueue the whole sequence
with one call. A plan is dropped as soon as a parameter / buffer of the module is replaced, mutated in place
(``_version``), moved (``data_ptr``) or `invalidate_all` is called (``transformer.invalidate_packed`` does).

Not used while per-launch profiling (`ops.profile`) or a CUDA-graph capture is active, or with ``B200ENC_PLANS=0``.
A plan owns its workspaces, so two forwards of one module on the SAME stream reuse them in stream order; forwards on
different streams get different plans. Replays of one plan are serialised by a lock.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref
from collections import OrderedDict
from ctypes import c_float, c_int, c_longlong, c_void_p

import torch
from torch import Tensor, nn

from . import _lib, ops

ENABLED = os.environ.get("B200ENC_PLANS", "1") != "0"
MAX_PLANS_PER_MODULE = 4
STATS = {"recorded": 0, "replayed": 0, "unplannable": 0, "last_unplannable": None}
_EPOCH = 0
# module -> OrderedDict[key, LaunchPlan | None]. Kept OUTSIDE the module's __dict__: plans hold ctypes arrays and a lock,
# which copy.deepcopy / pickle (EMA copies, torch.save(model)) must never meet; a copied module simply records its own.
_PLANS: "weakref.WeakKeyDictionary[nn.Module, OrderedDict]" = weakref.WeakKeyDictionary()
_SEEN: "weakref.WeakKeyDictionary[nn.Module, OrderedDict]" = weakref.WeakKeyDictionary()  # keys met once, not recorded yet
_LINEAR_PTR_FIELDS = ("x", "w", "bias", "colsum", "rowstats", "residual", "out", "stats_out", "acc_scale")
_LINEAR_PTR_TYPE = ctypes.POINTER(_lib.LinearArgs)


def invalidate_all() -> None:
    """Drop every recorded plan of this process (they are rebuilt on the next forward)."""
    global _EPOCH
    _EPOCH += 1


def enable(on: bool) -> bool:
    """Switch plan recording / replay on or off; returns the previous setting."""
    global ENABLED
    prev, ENABLED = ENABLED, bool(on)
    return prev


def _version(t: Tensor) -> int:
    try:
        return t._version
    except RuntimeError:  # inference tensors do not track a version counter
        return -1


class _Signature:
    """Flat snapshot of a module tree: which objects sit in every ``_modules`` / ``_parameters`` / ``_buffers`` slot and
    the (data_ptr, _version) of every tensor. `valid` re-checks it in ~0.1 ms for ViT-B (``module.parameters()`` alone
    takes 0.5 ms), and sees replaced parameters (``resize_pe``), in-place loads, device / dtype moves and surgery."""

    def __init__(self, module: nn.Module) -> None:
        self.training = module.training
        self.dicts: list[tuple[dict, int]] = []
        self.slots: list[tuple[dict, str, object]] = []
        self.tensors: list[Tensor] = []
        seen: set[int] = set()
        for m in module.modules():
            for d in (m._modules, m._parameters, m._buffers):
                self.dicts.append((d, len(d)))
                for k, v in d.items():
                    self.slots.append((d, k, v))
                    if isinstance(v, Tensor) and id(v) not in seen:
                        seen.add(id(v))
                        self.tensors.append(v)
        self.root = weakref.ref(module)  # weak: the plan cache is keyed by the module and must not keep it alive
        self.state = self._state()

    def _state(self) -> list:
        return [(t.data_ptr(), _version(t)) for t in self.tensors]

    def valid(self) -> bool:
        root = self.root()
        if root is None or root.training != self.training:
            return False
        for d, n in self.dicts:
            if len(d) != n:
                return False
        for d, k, v in self.slots:
            if d.get(k) is not v:
                return False
        return self._state() == self.state


class _Recorder:
    def __init__(self) -> None:
        self.thread = threading.get_ident()
        self.calls: list[tuple[str, tuple]] = []
        self.keep: list[Tensor] = []
        self.ok = True


def _span(t: Tensor) -> tuple[int, int]:
    lo = t.data_ptr()
    return lo, lo + max(t.numel() * t.element_size(), 1)


class LaunchPlan:
    """The recorded launches of one forward (see the module docstring)."""

    def __init__(self, rec: _Recorder, inputs: tuple[Tensor, ...], out: Tensor, sig: _Signature) -> None:
        n = len(rec.calls)
        self.n = n
        self.names = [name for name, _ in rec.calls]
        self.ops = (_lib.Op * n)()
        self.failed = c_int(-1)
        self.sig = sig
        self.epoch = _EPOCH
        self.lock = threading.Lock()
        self.out_shape, self.out_dtype, self.device = tuple(out.shape), out.dtype, out.device
        spans = [_span(t) for t in inputs] + [_span(out)]
        for a in range(len(spans)):
            for b in range(a + 1, len(spans)):
                if spans[a][0] < spans[b][1] and spans[b][0] < spans[a][1]:
                    # e.g. decoder(x, memory=x): a pointer inside both could not be attributed to one of them, and a later
                    # call with two distinct tensors would be patched wrongly
                    raise RuntimeError("the caller's tensors overlap in memory")
        # patch slots per input (and, last, for the output): (ctypes view, field name or index, byte offset)
        self.patches: list[list[tuple[object, object, int]]] = [[] for _ in spans]
        self._views: list[object] = []  # keeps the ctypes sub-objects the patch slots write through

        def classify(value, view, field) -> None:
            if not value:
                return
            for which, (lo, hi) in enumerate(spans):
                if lo <= value < hi:
                    self.patches[which].append((view, field, value - lo))
                    return

        for k, (name, args) in enumerate(rec.calls):
            op = self.ops[k]
            self._views.append(op)
            op.kind = _lib.OP_KINDS[name]
            argtypes = _lib._SIGNATURES[name][1][:-1]  # the trailing stream is supplied at replay
            if len(argtypes) != len(args):
                raise RuntimeError(f"{name}: recorded {len(args)} arguments, the signature has {len(argtypes)}")
            n_p = n_i = n_f = 0
            p_view, i_view, f_view = op.p, op.i, op.f
            self._views.append(p_view)
            for at, a in zip(argtypes, args):
                if at is _LINEAR_PTR_TYPE:
                    src = a._obj
                    lin = op.linear
                    self._views.append(lin)
                    ctypes.memmove(ctypes.byref(lin), ctypes.byref(src), ctypes.sizeof(_lib.LinearArgs))
                    for fname in _LINEAR_PTR_FIELDS:
                        classify(getattr(src, fname), lin, fname)
                elif at is c_void_p:
                    p_view[n_p] = a
                    classify(a, p_view, n_p)
                    n_p += 1
                elif at is c_int or at is c_longlong:
                    i_view[n_i] = int(a)
                    n_i += 1
                elif at is c_float:
                    f_view[n_f] = float(a)
                    n_f += 1
                else:
                    raise RuntimeError(f"{name}: argument type {at} cannot be recorded")
        # everything the launches touch stays alive with the plan — except the caller's tensors
        mine = {t.untyped_storage().data_ptr() for t in inputs} | {out.untyped_storage().data_ptr()}
        keep, seen = [], set()
        for t in rec.keep:
            sp = t.untyped_storage().data_ptr()
            if sp not in mine and id(t) not in seen:
                seen.add(id(t))
                keep.append(t)
        self.keep = keep
        if not self.patches[-1]:
            raise RuntimeError("no recorded launch writes the output tensor")

    def replay(self, inputs: tuple[Tensor, ...], out: Tensor, stream: int) -> None:
        with self.lock:
            for t, slots in zip(inputs + (out,), self.patches):
                base = t.data_ptr()
                for view, field, off in slots:
                    if isinstance(field, str):
                        setattr(view, field, base + off)
                    else:
                        view[field] = base + off
            rc = _lib.load().b200enc_run_ops(self.ops, self.n, ctypes.byref(self.failed), stream)
            k = self.failed.value
        ops.LAUNCHES += self.n if rc == 0 else max(k, 0)
        if rc != 0:
            _lib.check(rc, f"b200enc_run_ops: op {k} ({self.names[k] if 0 <= k < self.n else '?'})")


def _plannable_inputs(inputs: tuple[Tensor, ...]) -> bool:
    dev = inputs[0].device
    for t in inputs:
        if not (t.is_cuda and t.device == dev and t.is_contiguous() and t.numel() > 0):
            return False
    return True


def run(module: nn.Module, inputs: tuple[Tensor, ...], fn, extra_key: tuple = ()) -> Tensor:
    """``fn(*inputs)`` — through the module's recorded plan when there is one, else executed (and recorded).

    Contract for ``fn``: it returns ONE freshly allocated contiguous tensor; everything it computes from ``inputs`` is
    computed by `ops.*` launches (PyTorch operations may only build views or touch parameter-derived data, which a
    replay leaves as recorded); its launches depend on the inputs only through their shapes / dtypes and ``extra_key``."""
    if (not ENABLED or ops._PROFILE is not None or ops._RECORD is not None or not _plannable_inputs(inputs)
            or torch.cuda.is_current_stream_capturing()):
        return fn(*inputs)
    dev = inputs[0].device
    stream = torch.cuda.current_stream(dev).cuda_stream
    key = (tuple((tuple(t.shape), t.dtype) for t in inputs), dev, stream, extra_key)
    cache = _PLANS.get(module)
    if cache is None:
        cache = _PLANS[module] = OrderedDict()
    plan = cache.get(key, False)
    if plan is None:  # recorded before and found unplannable
        return fn(*inputs)
    if plan is False:
        # first sight of this key: run it plainly and only remember that it occurred. Shapes that never come back
        # (a generation loop re-running the decoder on a sequence that grows by one token per step, text/generator.py)
        # would otherwise pay for a signature + a plan + its workspaces on every call and evict the plans that matter.
        seen = _SEEN.get(module)
        if seen is None:
            seen = _SEEN[module] = OrderedDict()
        if key not in seen:
            seen[key] = True
            while len(seen) > 64:
                seen.popitem(last=False)
            return fn(*inputs)
        del seen[key]
    if plan is not False:
        if plan.epoch == _EPOCH and plan.sig.valid():
            cache.move_to_end(key)
            out = torch.empty(plan.out_shape, dtype=plan.out_dtype, device=dev)
            if dev.index != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    plan.replay(inputs, out, stream)
            else:
                plan.replay(inputs, out, stream)
            STATS["replayed"] += 1
            return out
        del cache[key]
    sig = _Signature(module)
    rec = _Recorder()
    ops._RECORD = rec
    try:
        out = fn(*inputs)
    finally:
        ops._RECORD = None
    plan = None
    if (rec.ok and rec.calls and isinstance(out, Tensor) and out.is_contiguous() and out.numel() > 0
            and out.untyped_storage().data_ptr() == out.data_ptr()
            and all(out.untyped_storage().data_ptr() != t.untyped_storage().data_ptr() for t in inputs)
            and sig.valid()):
        try:
            plan = LaunchPlan(rec, inputs, out, sig)
        except RuntimeError as e:  # e.g. the output is not written by any launch: stay on the per-call path
            STATS["last_unplannable"] = f"{type(module).__name__}: {e}"
            plan = None
    STATS["recorded" if plan is not None else "unplannable"] += 1
    cache[key] = plan
    while len(cache) > MAX_PLANS_PER_MODULE:
        cache.popitem(last=False)
    return out


def clear(module: nn.Module) -> None:
    """Forget the plans of ``module`` (and free the workspaces they hold)."""
    _PLANS.pop(module, None)
    _SEEN.pop(module, None)


def plans_of(module: nn.Module) -> dict:
    """The recorded plans of ``module`` by key (None = a forward that turned out not to be plannable); for tests."""
    return dict(_PLANS.get(module, {}))
