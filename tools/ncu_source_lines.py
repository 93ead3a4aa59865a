#!/usr/bin/env python
"""Aggregate the warp-stall samples of an ncu source-page export by CUDA source line.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_source_lines.py src.csv [top]

The export holds one table per source file (view cuda,sass: every SASS row is preceded by its source line)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
cur_file = None
cur_line = '?'
hdr = None
per_line = defaultdict(lambda: defaultdict(int))
text = {}
total = 0
kernel_idx = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        n = int(d["# Samples"] or 0)
    except ValueError:
        continue
    if r[0].strip():
        cur_line = r[0].strip()
    if not d.get("Address"):
        if r[1].strip():
            text[(cur_file, r[0].strip())] = r[1].strip()[:110]
        continue
    key = (cur_file, cur_line)
    if True:                  # a SASS row: its samples count under the source line it follows
        per_line[key]["n"] += n
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    per_line[key][k] += int(d[k] or 0)
                except ValueError:
                    pass
        total += n
print(f"total samples {total}")
for key, v in sorted(per_line.items(), key=lambda kv: -kv[1]["n"])[:top]:
    st = sorted(((k[6:], c) for k, c in v.items() if k != "n" and c), key=lambda x: -x[1])[:3]
    print(f"{100.0 * v['n'] / max(total, 1):5.1f}%  {key[0]}:{key[1]:>5}  {text.get(key, '')}\n        {st}")
