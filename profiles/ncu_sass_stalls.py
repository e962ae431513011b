#!/usr/bin/env python3
"""Per-SASS-instruction stall breakdown of an .ncu-rep (needs --set full / source counters).
usage: ncu_sass_stalls.py REPORT.ncu-rep [TOPN]
Prints: totals per stall reason, then the TOPN instructions by no_inst samples with 3 instructions of context."""
import csv, io, subprocess, sys
rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
data = []
for r in rows:
    if len(r) > 10 and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(int(r[ix[h]] or 0) for r in data) for h in reasons}
allsamp = sum(int(r[ix["# Samples"]] or 0) for r in data)
inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("SASS instructions:", len(data), " executed warp-inst:", inst, " samples:", allsamp)
print("stall totals:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
key = "stall_no_inst"
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix[key]] or 0))[:N]
for i in order:
    print("---- no_inst=%s samples=%s exec=%s" % (data[i][ix[key]], data[i][ix["# Samples"]], data[i][ix["Instructions Executed"]]))
    for j in range(max(0, i - 3), min(len(data), i + 2)):
        r = data[j]
        print("   %s %-60s exec=%-10s samp=%-6s noinst=%-6s wait=%-5s lsb=%-5s br=%s" % ("=>" if j == i else "  ", r[1].strip()[:60], r[ix["Instructions Executed"]],
              r[ix["# Samples"]], r[ix[key]], r[ix["stall_wait"]], r[ix["stall_long_sb"]], r[ix["stall_branch_resolving"]]))
