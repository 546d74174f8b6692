#!/usr/bin/env python
"""Experiment: ViT-B/16 forward, eager launches vs CUDA-graph replay (launch-gap measurement)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pytorch_models_b200 as pm
from bench import CONFIGS, synthetic_weights_

cfg = CONFIGS["c2"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = cfg["make"](pm).eval(); synthetic_weights_(m, 100); m = m.cuda().bfloat16()
x = torch.randn(B, 3, 224, 224, device="cuda", dtype=torch.bfloat16)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters): fn()
    e1.record(); t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, t_cpu / iters * 1e3

with torch.no_grad():
    gpu_ms, cpu_ms = timeit(lambda: m(x))
    print(f"eager: {gpu_ms:.3f} ms/step GPU, CPU enqueue time {cpu_ms:.3f} ms/step")
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
