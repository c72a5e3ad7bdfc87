"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: one iteration of update() =
the launches between the last two pose_kernel launches.  usage: python tools/launch_summary.py launches.csv [out.md] [title]"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_us = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)
        rows.append((r["Kernel Name"].split("(")[0], v_us))
pose = [i for i, (k, _) in enumerate(rows) if k == "pose_kernel"]
if len(pose) >= 2:
    rows = rows[pose[-2]:pose[-1]]
agg = OrderedDict()
for k, v in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(v for _, v in rows)
out = []
title = sys.argv[3] if len(sys.argv) > 3 else "ncu launch list, one update() iteration"
out.append(f"# {title}\n")
out.append(f"launches per iteration: {len(rows)}, sum of kernel times: {total / 1e3:.3f} ms (cold-cache, serialised: compare SHARES)\n")
out.append("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {n} | {v / 1e3:.3f} | {v / n:.1f} | {100 * v / total:.1f}% |")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
