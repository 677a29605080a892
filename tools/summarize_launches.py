#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/summarize_launches.py file.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(float)
cnt = collections.Counter()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
    name = row["Kernel Name"]
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    if "gemm_tcgen05_kernel" in short:
        short = short[:60]
    else:
        short = re.sub(r"<.*", "", short)[:60]
    tot[short] += v
    cnt[short] += 1
T = sum(tot.values())
print(f"launches {sum(cnt.values())}  total {T:.2f} ms (cold-cache, serialised: compare shares)")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:40]:
    print(f"{v:9.3f} ms {100 * v / T:5.1f}%  n={cnt[k]:5d}  {k}")
