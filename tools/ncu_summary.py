#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/.
usage: ncu_summary.py <rep> [kernel-regex] > profiles/xxx.txt   (runs on the CPU box: ncu -i)"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "k_residual"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name"), " id", d.get("ID"))
    for k in KEYS:
        if k in d:
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    st = {k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(v)
          for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")}
    print("  stalls per issue:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr) and r[idx["# Samples"]].isdigit()]
    # first kernel instance only
    seen, first = set(), []
    for r in data:
        if r[idx["Address"]] in seen:
            break
        seen.add(r[idx["Address"]]); first.append(r)
    data = first
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"== source page: {len(data)} SASS instructions, {tot} samples; top stall sites")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:25]:
        why = {k: int(r[idx[k]]) for k in stalls if int(r[idx[k]]) > 0}
        why = sorted(why.items(), key=lambda kv: -kv[1])[:3]
        print(f"  {100*int(r[idx['# Samples']])/tot:5.1f}%  {r[idx['Source']].strip()[:60]:60s} {why}")
    step = max(len(data) // 24, 1)
    print("== samples by SASS region")
    for lo in range(0, len(data), step):
        seg = data[lo:lo + step]
        s = sum(int(r[idx["# Samples"]]) for r in seg)
        agg = collections.Counter()
        for r in seg:
            for k in stalls:
                agg[k] += int(r[idx[k]])
        print(f"  [{lo:5d},{lo+len(seg):5d}) {100*s/tot:5.1f}%  {agg.most_common(3)}  first: {seg[0][idx['Source']].strip()[:40]}")
