import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = [r for r in rows if len(r) > 10 and r[0] == "Address"][0]
ix = {h: i for i, h in enumerate(hdr)}
data = [(int(r[0], 16), int(r[ix["Instructions Executed"]] or 0), int(r[ix["stall_no_inst"]] or 0)) for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
tot = sum(d[1] for d in data)
base = min(d[0] for d in data)
# 128-byte lines touched weighted by exec
lines = {}
for a, e, n in data:
    l = (a - base) // 128
    lines[l] = lines.get(l, 0) + e
ls = sorted(lines.values(), reverse=True)
acc = 0
for frac in (0.5, 0.8, 0.9, 0.95, 0.99):
    acc = 0
    for i, v in enumerate(ls):
        acc += v
        if acc >= frac * tot:
            print("%.0f%% of executed instructions come from %d lines = %.1f KB" % (100 * frac, i + 1, (i + 1) * 128 / 1024)); break
print("total lines with any execution:", sum(1 for v in ls if v > 0), "of", len(ls), "=", sum(1 for v in ls if v > 0) * 128 / 1024, "KB")
