#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/launches.csv, gpurun_out/prof_layer.ncu-rep) into the tracked summaries under profiles/."""
import collections
import csv
import re
import subprocess
import sys

out_dir = sys.argv[1] if len(sys.argv) > 1 else "profiles/r01"

rows = list(csv.reader(open("gpurun_out/launches.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
kn, mv, mn, mu = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
agg = collections.OrderedDict()
tot = 0.0
for r in data:
    if r[mn] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void b200::", "").replace("b200::", "")
    t = float(r[mv].replace(",", "")) / (1000.0 if r[mu] == "ns" else 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
with open(f"{out_dir}/ncu_launch_list_summary.md", "w") as f:
    f.write("# ncu launch list: one warmed-up ViT-B/16 forward, batch 1024 (gpu__time_duration.sum, --clock-control none)\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's step_breakdown, not absolutes.\n\n")
    f.write(f"launches: {sum(a[0] for a in agg.values())}, total {tot / 1000:.2f} ms\n\n| launches | total us | share | kernel |\n|---:|---:|---:|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {n} | {t:.1f} | {100 * t / tot:.2f}% | `{k[:100]}` |\n")
print(open(f"{out_dir}/ncu_launch_list_summary.md").read())

raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_layer.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = {h: i for i, h in enumerate(hdr)}
with open(f"{out_dir}/ncu_full_layer_summary.md", "w") as f:
    f.write("# ncu --set full, one encoder layer of ViT-B/16 at batch 1024 (--clock-control none)\n\n")
    f.write("| metric | " + " | ".join(re.sub(r"\(.*", "", d[idx["Kernel Name"]]).replace("void b200::", "").replace("b200::", "")[:40] for d in data) + " |\n")
    f.write("|---|" + "---:|" * len(data) + "\n")
    for w in want[1:]:
        if w in idx:
            f.write(f"| {w} [{units[idx[w]]}] | " + " | ".join(d[idx[w]][:14] for d in data) + " |\n")
print(open(f"{out_dir}/ncu_full_layer_summary.md").read())
