#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py <launches.csv> <out.csv>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
unit = rows[start][ui]
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("reef::", "")
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
out = ["kernel,launches,total_%s,share" % unit]
for k, v in tot.most_common():
    out.append("%s,%d,%.1f,%.4f" % (k, cnt[k], v, v / T))
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out[:24]))
