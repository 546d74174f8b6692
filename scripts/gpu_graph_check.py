#!/usr/bin/env python
"""ViT-B/16 forward: per-launch ctypes calls vs the recorded launch plan (plans.py) vs CUDA-graph replay — GPU time
and CPU enqueue time per forward."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pytorch_models_b200 as pm
from pytorch_models_b200 import plans
from bench import CONFIGS, synthetic_weights_

cfg = CONFIGS["c2"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = cfg["make"](pm).eval(); synthetic_weights_(m, 100); m = m.cuda().bfloat16()
x = torch.randn(B, 3, 224, 224, device="cuda", dtype=torch.bfloat16)

def timeit(fn, iters=20, cpu_iters=8):
    """GPU ms per forward over `iters`; CPU enqueue ms per forward over the first `cpu_iters` only (8 x 65 launches
    stay below the driver's launch-queue depth, so the CPU never waits for the GPU)."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t0 = time.perf_counter()
    for _ in range(cpu_iters): fn()
    t_cpu = time.perf_counter() - t0
    for _ in range(iters - cpu_iters): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, t_cpu / cpu_iters * 1e3

with torch.no_grad():
    plans.enable(False)
    y_calls = m(x)
    gpu_ms, cpu_ms = timeit(lambda: m(x))
    print(f"per-launch ctypes calls: {gpu_ms:.3f} ms/step GPU, CPU enqueue time {cpu_ms:.3f} ms/step")
    plans.enable(True)
    gpu_ms, cpu_ms = timeit(lambda: m(x))
    print(f"launch plan (one b200enc_run_ops call): {gpu_ms:.3f} ms/step GPU, CPU enqueue time {cpu_ms:.3f} ms/step; "
          f"stats {plans.STATS}; output equals per-launch path: {torch.equal(m(x), y_calls)}")
    for iters in (8, 14, 20, 40):  # launches in flight: does a deep launch queue (CPU far ahead) cost GPU time?
        plans.enable(False)
        g0, c0 = timeit(lambda: m(x), iters=iters)
        plans.enable(True)
        g1, c1 = timeit(lambda: m(x), iters=iters)
        print(f"  {iters} forwards in flight: per-launch {g0:.3f} ms GPU / {c0:.3f} ms CPU, plan {g1:.3f} ms GPU / {c1:.3f} ms CPU")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): m(x)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = m(x)
    gpu_ms, cpu_ms = timeit(lambda: g.replay())
    print(f"graph: {gpu_ms:.3f} ms/step GPU, CPU enqueue time {cpu_ms:.3f} ms/step")
    y2 = m(x)
    print("graph output equals eager:", torch.equal(y, y2))
