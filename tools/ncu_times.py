#!/usr/bin/env python
"""usage: tools/ncu_times.py launches.csv -- mean duration / DRAM bytes of k_face_flux, k_element_rk stage 1 and stages 2-4 from
an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of bench.py"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[iid]), r[ik]), {})[r[im]] = float(r[iv].replace(",", ""))
grp = collections.defaultdict(list)
ne = 0
for (i, k), v in d.items():
    if "k_face_flux" in k:
        grp["face"].append(v)
    elif "k_element_rk" in k or "k_fused" in k:
        grp["elem_stage1" if ne % 4 == 0 else "elem_stage2-4"].append(v)
        ne += 1
out = []
for g, vs in grp.items():
    t = sum(v["gpu__time_duration.sum"] for v in vs) / len(vs) / 1e3
    if "dram__bytes_read.sum" in vs[0]:
        b = sum(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in vs) / len(vs) / 1e6
        out.append("%s %.1f us %.0f MB (%d launches)" % (g, t, b, len(vs)))
    else:
        out.append("%s %.1f us (%d launches)" % (g, t, len(vs)))
print(sys.argv[1].split("/")[-1], " | ".join(out))
