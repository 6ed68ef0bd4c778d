#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares of ONE step
(the launches between the last two pack_input kernels).  Usage: summarize_launches.py launches.csv [out.md]"""
import collections
import csv
import sys

src = sys.argv[1]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
idx = [i for i, n in enumerate(names) if "pack_input" in n]
s, e = idx[-2], idx[-1]
agg = collections.OrderedDict()
tot = 0.0
for r in rows[s:e]:
    n = r["Kernel Name"].split("(")[0][:70]
    t = float(r["Metric Value"]) / 1000.0
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
out = ["| kernel | launches/step | us/step (ncu, cold cache, serialised) | share |", "|---|---|---|---|"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| `%s` | %d | %.1f | %.1f%% |" % (n, c, t, 100 * t / tot))
out.append("| **total** | %d | %.1f | 100%% |" % (e - s, tot))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
