#!/usr/bin/env python
"""Executed warp-instructions per SASS region and opcode class (from the ncu source page)."""
import csv, io, subprocess, sys, collections, re
rep, kre = sys.argv[1], sys.argv[2]
nreg = int(sys.argv[3]) if len(sys.argv) > 3 else 20
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[idx["# Samples"]].isdigit()]
seen, first = set(), []
for r in data:
    if r[idx["Address"]] in seen: break
    seen.add(r[idx["Address"]]); first.append(r)
data = first
tot = sum(int(r[idx["Instructions Executed"]]) for r in data)
print("total warp instructions", tot)
byop = collections.Counter()
for r in data:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[idx["Source"]])
    byop[m.group(1) if m else "?"] += int(r[idx["Instructions Executed"]])
print("by opcode:", [(k, round(100 * v / tot, 1)) for k, v in byop.most_common(22)])
step = max(len(data) // nreg, 1)
for lo in range(0, len(data), step):
    seg = data[lo:lo + step]
    ex = sum(int(r[idx["Instructions Executed"]]) for r in seg)
    sm = sum(int(r[idx["# Samples"]]) for r in seg)
    ops = collections.Counter()
    for r in seg:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[idx["Source"]])
        ops[m.group(1) if m else "?"] += int(r[idx["Instructions Executed"]])
    print(f"[{lo:5d},{lo+len(seg):5d}) exec {100*ex/tot:5.1f}%  samples {sm:6d}  {[(k, round(100*v/max(ex,1))) for k, v in ops.most_common(5)]}")
