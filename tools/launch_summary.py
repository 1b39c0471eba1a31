"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) into per-kernel totals
and the launch sequence of one update:  python tools/launch_summary.py gpurun_out/X.csv [first [count]]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
idx = {h: i for i, h in enumerate(hdr)}
tot, cnt, seq = collections.OrderedDict(), collections.Counter(), []
for r in data:
    if len(r) < len(hdr):
        continue
    v, unit = float(r[idx["Metric Value"]]), r[idx["Metric Unit"]]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    short = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
    tot[short] = tot.get(short, 0) + v
    cnt[short] += 1
    seq.append((short, round(v, 1), r[idx["Grid Size"]]))
T = sum(tot.values())
print(f"total {T:.1f} us, {len(seq)} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:9.1f} us {100 * v / T:5.1f}% x{cnt[k]:3d}  {k}")
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for s in seq[first:first + count]:
    print(f"  {s[1]:7.1f} us  {s[2]:14s} {s[0]}")
