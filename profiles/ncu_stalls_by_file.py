import csv, io, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and (r[0] in ("Line", "#") or r[1] == "Source"):
        hdr = r; continue
    if hdr and cur and len(r) == len(hdr) and r[0].isdigit():
        ix = {h: i for i, h in enumerate(hdr)}
        for k in ("# Samples", "Instructions Executed", "stall_no_inst", "stall_long_sb", "stall_wait", "stall_branch_resolving", "stall_short_sb", "stall_selected"):
            if k in ix:
                try: agg[cur][k] += int(r[ix[k]] or 0)
                except ValueError: pass
tot = collections.Counter()
for f, c in agg.items(): tot.update(c)
print("total", dict(tot))
for f, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
    if c["# Samples"] < 0.005 * tot["# Samples"]: continue
    print("%-28s inst %5.1f%%  samples %5.1f%%  | of its samples: no_inst %4.1f%% long_sb %4.1f%% wait %4.1f%% branch %4.1f%% selected %4.1f%%" % (
        f, 100*c["Instructions Executed"]/tot["Instructions Executed"], 100*c["# Samples"]/tot["# Samples"],
        100*c["stall_no_inst"]/max(1,c["# Samples"]), 100*c["stall_long_sb"]/max(1,c["# Samples"]), 100*c["stall_wait"]/max(1,c["# Samples"]),
        100*c["stall_branch_resolving"]/max(1,c["# Samples"]), 100*c["stall_selected"]/max(1,c["# Samples"])))
