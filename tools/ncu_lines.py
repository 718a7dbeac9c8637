#!/usr/bin/env python
"""Per-source-line instruction shares of one kernel from an .ncu-rep captured with --import-source on:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME | python tools/ncu_lines.py [top_n]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file, hdr = "", None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        ti = hdr.index("Thread Instructions Executed")
        si = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":  # source-line rows have '-' as address
        continue
    try:
        n, t, s = int(r[ii]), int(r[ti]), int(r[si])
    except ValueError:
        continue
    k = (cur_file, int(r[0]))
    a = agg.setdefault(k, [0, 0, 0, r[1].strip()[:120]])
    a[0] += n
    a[1] += t
    a[2] += s
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print(f"total warp instructions {tot}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / tot:5.1f}% inst {100 * a[2] / tots:5.1f}% samples  lanes={a[1] / max(a[0], 1):4.1f}  {f}:{ln}  {a[3]}")
