#!/usr/bin/env python
"""Turn the round-2 ncu captures (scripts/gpu_r2_ncu.sh -> gpurun_out/r2_*.ncu-rep, r2_launches_c2.csv) into the tracked
summaries under profiles/r02/: one markdown table per capture (duration, clocks, tensor / XU / issue utilisation, DRAM
bytes and throughput, L2 hit rate, L2->SM sector efficiency, registers) and profiles/r02/ncu_traffic.json, the
per-launch DRAM traffic that bench.py reports as `roofline.traffic`."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

out_dir = sys.argv[1] if len(sys.argv) > 1 else "profiles/r02"
os.makedirs(out_dir, exist_ok=True)
HBM_PEAK = 6541.1  # MEASURED_PEAKS.json hbm_gbs (GB/s)

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_requests_srcunit_tex_op_read.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def short(name: str) -> str:
    return re.sub(r"\(.*", "", name).replace("void b200::", "").replace("b200::", "").replace("void ", "")[:46]


def to_float(s: str) -> float:
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def raw_page(rep: str):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return None
    return rows[0], rows[1], rows[2:]


def bytes_of(val: str, unit: str) -> float:
    v = to_float(val)
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)


def us_of(val: str, unit: str) -> float:
    v = to_float(val)
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)


traffic = {}
for rep, title in (("r2_c2_head", "C2 (ViT-B/16, batch 1024): patch rows, patch-embedding GEMM, class-token rows, encoder layer 0"),
                   ("r2_c2_ln", "C2: final LayerNorm (class-token rows only)"),
                   ("r2_rows", "row kernels at realistic sizes (tests/tools/gpu_row_kernels.py)"),
                   ("r2_c3_layer", "C3 (ViT-L/16 SigLIP 384, batch 256): one encoder layer"),
                   ("r2_c4_layer", "C4 (DINOv2 L/14 518, batch 128): one encoder layer"),
                   ("r2_c5_layer", "C5 (Whisper large-v3 encoder, batch 64): one encoder layer"),
                   ("r2_attn_l197", "attention kernel alone, L=197, B=1024, H=12 (selftest perf_vitb_b1024)"),
                   ("r2_attn_l1500", "attention kernel alone, L=1500, B=64, H=20 (selftest perf_whisper_b64)"),
                   ("r2_cudnn_sdpa_l1500", "LIBRARY BASELINE: cuDNN SDPA kernel, L=1500, B=64, H=20, bf16 (tests/tools/gpu_sdpa_cudnn_probe.py)")):
    path = f"gpurun_out/{rep}.ncu-rep"
    if not os.path.exists(path):
        print("missing", path)
        continue
    page = raw_page(path)
    if page is None:
        print("empty", path)
        continue
    hdr, units, data = page
    idx = {h: i for i, h in enumerate(hdr)}
    with open(f"{out_dir}/ncu_{rep[3:]}.md", "w") as f:
        f.write(f"# ncu --set full --clock-control none: {title}\n\n")
        f.write("Per-launch values of ONE launch each (ncu replays the kernel; durations are cold-cache, serialised).\n\n")
        f.write("| metric | " + " | ".join(short(d[idx["Kernel Name"]]) for d in data) + " |\n")
        f.write("|---|" + "---:|" * len(data) + "\n")
        for w in WANT:
            if w in idx:
                f.write(f"| {w} [{units[idx[w]]}] | " + " | ".join(d[idx[w]][:14] for d in data) + " |\n")
        # derived: DRAM GB/s and fraction of the measured copy bandwidth
        if "dram__bytes_read.sum" in idx:
            tot = [bytes_of(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) +
                   bytes_of(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for d in data]
            us = [us_of(d[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]]) for d in data]
            f.write("| DRAM read+write [GB] | " + " | ".join(f"{t / 1e9:.4f}" for t in tot) + " |\n")
            f.write("| DRAM GB/s (traffic / duration) | " + " | ".join(f"{t / u / 1e3:.0f}" for t, u in zip(tot, us)) + " |\n")
            f.write(f"| fraction of measured HBM copy peak ({HBM_PEAK:.0f} GB/s) | " +
                    " | ".join(f"{t / u / 1e3 / HBM_PEAK:.2f}" for t, u in zip(tot, us)) + " |\n")
            if rep == "r2_c2_head":
                names = [short(d[idx["Kernel Name"]]) for d in data]
                gemm = [t for n, t in zip(names, tot) if n.startswith("gemm_bf16")]
                att = [t for n, t in zip(names, tot) if n.startswith("attention")]
                layer = gemm[1:5]  # gemm[0] is the patch embedding; then QKV, out_proj, FC1, FC2 of layer 0 (in launch order)
                traffic["c2_b1024"] = {
                    "gemm_mean_bytes_per_launch": sum(layer) / max(len(layer), 1),
                    "gemm_bytes_per_launch": dict(zip(["qkv", "out_proj", "fc1", "fc2"], layer)),
                    "patch_embed_gemm_bytes": gemm[0] if gemm else None,
                    "attention_bytes_per_launch": att[0] if att else None,
                    "algorithmic_bytes": {"qkv": 2 * 201728 * (768 + 2304), "out_proj": 2 * 201728 * (768 + 768 + 768),
                                          "fc1": 2 * 201728 * (768 + 3072), "fc2": 2 * 201728 * (3072 + 768 + 768),
                                          "attention": 8 * 197 * 768 * 1024},
                    "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, "
                              "profiles/r02/ncu_c2_head.md (scripts/gpu_r2_ncu.sh)",
                }
    print(open(f"{out_dir}/ncu_{rep[3:]}.md").read())

if traffic:
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    for v in traffic.values():
        v["captured_at_commit"] = head
    json.dump(traffic, open(f"{out_dir}/ncu_traffic.json", "w"), indent=1)

if os.path.exists("gpurun_out/r2_launches_c2.csv"):
    rows = list(csv.reader(open("gpurun_out/r2_launches_c2.csv")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    kn, mv, mn, mu = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    agg = collections.OrderedDict()
    tot = 0.0
    for r in data:
        if r[mn] != "gpu__time_duration.sum":
            continue
        t = us_of(r[mv], r[mu])
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    with open(f"{out_dir}/ncu_launch_list_c2.md", "w") as f:
        f.write("# ncu launch list: one warmed-up ViT-B/16 forward, batch 1024 (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's step_breakdown_ms.\n\n")
        f.write(f"launches: {sum(a[0] for a in agg.values())}, total {tot / 1000:.2f} ms\n\n| launches | total us | share | kernel |\n|---:|---:|---:|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {n} | {t:.1f} | {100 * t / tot:.2f}% | `{k}` |\n")
    print(open(f"{out_dir}/ncu_launch_list_c2.md").read())
