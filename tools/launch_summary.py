#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel, launches and mean duration of the last few."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    agg.setdefault(re.sub(r"\(.*", "", r[ki])[:70], []).append(v)
tail = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for k, v in agg.items():
    print(f"{k:70s} n={len(v):5d} last{tail}_mean_us={sum(v[-tail:]) / len(v[-tail:]):10.1f} max_us={max(v):10.1f}")
