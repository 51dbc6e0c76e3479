#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / total / mean."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = collections.OrderedDict()
for r in rows[hdr + 2 + skip:]:
    if len(r) > vi:
        d.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):4d} sum={sum(v) / 1e6:9.3f} ms  mean={sum(v) / len(v) / 1e3:9.1f} us")
