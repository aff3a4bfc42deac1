#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/launch_summary.py gpurun_out/launches.csv "command line" > profiles/launches_xxx.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r[ik]).replace("rvs::", "").replace("(anonymous namespace)::", "").strip()
    tot[name] += us
    cnt[name] += 1
all_us = sum(tot.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("(per-launch times are cold-cache and serialised: compare SHARES)\n")
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
for k, v in tot.most_common():
    print(f"{k[:70]:70s} {cnt[k]:8d} {v:12.1f} {100 * v / all_us:6.1f}%")
