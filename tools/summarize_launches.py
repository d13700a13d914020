#!/usr/bin/env python3
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel launches / total ms / share.
usage: summarize_launches.py raw.csv 'header comment' > summary.csv"""
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = r[ki]
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"\(.*$", "", short)
    tag = " [G2]" if ("G2" in name) else ""
    short = re.sub(r"<.*$", "", short) + tag
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
for c in sys.argv[2:]:
    print("# " + c)
print("kernel,launches,total_ms,share")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.3f,%.4f" % (k, n, ms, ms / tot))
