#!/usr/bin/env python
"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:  python tools/launch_summary.py launches.csv "command" """
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t, n = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv or "gpu__time_duration" not in ",".join(r):
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    t[r[ik]] += v
    n[r[ik]] += 1
tot = sum(t.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print(f"total launches {sum(n.values())}, total device time {tot:.1f} ms (cold-cache, serialised: compare shares)")
for k in sorted(t, key=lambda k: -t[k]):
    print(f"{n[k]:6d} launches {t[k]:10.3f} ms {100 * t[k] / tot:6.2f}%  {k[:150]}")
