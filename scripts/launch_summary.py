#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and share.
usage: python scripts/launch_summary.py gpurun_out/<tag>_launches.csv [steps]   (steps = bench steps incl. warm-up captured)"""
import csv, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else None
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("<unnamed>::", "")
    ns = float(r[-1])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us total (cold-cache, serialised under ncu: compare SHARES with bench.py's kernel_ms_per_step)")
print(f"{'kernel':48s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:48s} {n:8d} {ns / 1e3:10.1f} {ns / 1e3 / n:9.2f} {100 * ns / tot:6.1f}%")
