#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: tools/launch_shares.py launches.csv [skip_first_n]"""
import csv
import collections
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1 + skip:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1e3 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1e3)
    name = r[ki].split("(")[0]
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':40s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for k, v in tot.most_common():
    print(f"{k[:40]:40s} {cnt[k]:8d} {v:10.1f} {v / cnt[k]:8.1f} {100 * v / total:5.1f}%")
print(f"{'TOTAL':40s} {sum(cnt.values()):8d} {total:10.1f}")
