#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics + the source lines that execute the most warp instructions / collect the
most stall samples (needs -lineinfo and --import-source on).  usage: ncu_top_lines.py REPORT.ncu-rep [N]"""
import csv, io, subprocess, sys, collections, json

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__grid_size", "launch__block_size"]
head = {h: (v + " " + u).strip() for h, u, v in zip(hdr, units, vals) if h in keys}
stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
          for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
print(json.dumps(head, indent=1))
print("stalls per issue:", {k: round(v, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur = None
agg = collections.defaultdict(lambda: [0, 0, ""])
for row in csv.reader(io.StringIO(src)):
    if len(row) == 2 and row[0] == "File Path":
        cur = row[1].split("/")[-1]
        continue
    if len(row) > 8 and row[0].isdigit() and cur:
        try:
            inst = int(row[7]); samp = int(row[6])
        except ValueError:
            continue
        a = agg[(cur, int(row[0]))]
        a[0] += inst; a[1] += samp; a[2] = row[1].strip()[:110]
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), a in agg.items():
    byfile[f][0] += a[0]; byfile[f][1] += a[1]
print("by file (inst%, stall-sample%):", {f: (round(100 * v[0] / tot, 1), round(100 * v[1] / tots, 1)) for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0])})
print("top lines by warp instructions executed:")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / tots, f, l, a[2]))
print("top lines by stall samples:")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * a[1] / tots, 100 * a[0] / tot, f, l, a[2]))
