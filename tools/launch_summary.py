#!/usr/bin/env python3
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
usage: python tools/launch_summary.py gpurun_out/launches_raw.csv [out.csv]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v * 1e3 if r[iu] in ("s", "second") else v
    name = re.sub(r"\(.*", "", r[ik])[:110]
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
lines = ["kernel,launches,total_ms,share"] + [f'"{k}",{cnt[k]},{v:.3f},{v / total:.4f}' for k, v in tot.most_common()]
txt = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
print(txt)
