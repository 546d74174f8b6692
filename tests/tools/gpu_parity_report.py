#!/usr/bin/env python
"""Parity report on the GPU box: for each BASELINE config, error of (a) our sm_100a path and (b) PyTorch's own bf16
forward of the same math (oracle_torch on CUDA bf16 tensors: the 12-kernels-per-layer library path the reference
would take under .cuda().bfloat16()) against the fp32 CPU oracle on identical seeded weights / inputs.
Writes gpurun_out/parity_report.json. Test infrastructure (uses oracle/)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pytorch_models_b200 as pm  # noqa: E402
from bench import CONFIGS, synthetic_weights_  # noqa: E402
from conftest import error_stats  # noqa: E402
from oracle import oracle_torch  # noqa: E402

SAMPLES = {"c2": 8, "c3": 2, "c4": 1, "c5": 1}


@torch.no_grad()
def main() -> None:
    report = {}
    for name in sys.argv[1:] or ["c2", "c3", "c4", "c5"]:
        cfg = CONFIGS[name]
        torch.manual_seed(0)
        m = cfg["make"](pm).eval()
        synthetic_weights_(m, 100)
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        torch.manual_seed(1)
        x = torch.randn(SAMPLES[name], *cfg["shape"])

        def fwd(sd_, x_):
            if cfg["kind"] == "vit":
                return oracle_torch.vit_forward(sd_, x_, cfg["heads"], cfg["pool"])
            return oracle_torch.whisper_encoder_forward(sd_, x_)

        want = fwd(sd, x).numpy()
        ours = m.cuda().bfloat16()(x.cuda().bfloat16()).float().cpu().numpy()
        sd_gpu = {k: v.cuda().bfloat16() if v.is_floating_point() else v.cuda() for k, v in sd.items()}
        lib = fwd(sd_gpu, x.cuda().bfloat16()).float().cpu().numpy()
        o_abs, o_cos = error_stats(ours, want)
        t_abs, t_cos = error_stats(lib, want)
        report[name] = dict(workload=cfg["desc"], samples=SAMPLES[name], out_rms=float(np.sqrt((want ** 2).mean())),
                            ours=dict(max_abs=o_abs, min_cos=o_cos, mean_abs=float(np.abs(ours - want).mean())),
                            torch_bf16=dict(max_abs=t_abs, min_cos=t_cos, mean_abs=float(np.abs(lib - want).mean())))
        print(name, json.dumps(report[name]), flush=True)
        del m, sd_gpu
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
