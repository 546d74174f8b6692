#!/usr/bin/env python
"""SASS opcode census of libb200enc.so per kernel: how many tcgen05 / TMEM / TMA instructions each kernel carries
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, MUFU.EX2, FFMA2 = packed fp32). Legacy tensor paths (HMMA = mma.sync) must be absent.
usage: python scripts/sass_census.py [path/to/libb200enc.so] > profiles/r02/sass_census.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "pytorch_models_b200/libb200enc.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2", "HMMA", "LDGSTS"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                per[cur][o] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            per[cur]["UTCHMMA.2CTA"] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except OSError:
        return n


print(f"# SASS opcode census of `{lib}` (cuobjdump -sass, sm_100a)\n")
print("| kernel | instrs | " + " | ".join(OPS) + " |")
print("|---|---:|" + "---:|" * len(OPS))
tot = collections.Counter()
for k, c in per.items():
    name = re.sub(r"^void ", "", demangle(k))
    name = re.sub(r"\(.*", "", name).replace("b200::", "")
    print(f"| `{name[:70]}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    tot.update(c)
print(f"| **all {len(per)} kernels** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
