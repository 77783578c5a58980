#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instruction mix by opcode and the hottest SASS instructions."""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci, cs, csrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
cth = h.index("Avg. Threads Executed")
data = []
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[ci]:
        continue
    data.append((r[0], r[csrc], int(r[ci] or 0), int(r[cs] or 0), r[cth]))
tot_i = sum(d[2] for d in data)
tot_s = sum(d[3] for d in data)
print(f"SASS instructions: {len(data)}  executed (warp-level): {tot_i:,}  samples: {tot_s:,}")
mix = collections.Counter()
smix = collections.Counter()
for a, src, n, s, th in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    mix[op] += n
    smix[op] += s
print("opcode mix (executed %, samples %):")
for op, n in mix.most_common(22):
    print(f"  {op:10s} {100*n/tot_i:5.1f}%  {100*smix[op]/max(tot_s,1):5.1f}%")
print(f"hottest {top} instructions by samples:")
for i in sorted(range(len(data)), key=lambda i: -data[i][3])[:top]:
    a, src, n, s, th = data[i]
    print(f"  [{i:4d}] {100*s/max(tot_s,1):5.2f}% smp  {100*n/tot_i:5.2f}% exe thr={th:>5s}  {src[:100]}")
