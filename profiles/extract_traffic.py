#!/usr/bin/env python3
"""Pull dram__bytes_read.sum + dram__bytes_write.sum (`ncu --set full`) out of .ncu-rep files and record them in
profiles/traffic.json, which bench.py quotes as roofline.traffic when it runs the same workload.  A report holds the launches of
ONE step (two for the single-end two-pass launch, four for the staged paired launch); their bytes are summed.
usage: extract_traffic.py KEY REPORT.ncu-rep "description of the captured command" [KEY REPORT DESC ...]"""
import csv, hashlib, io, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "traffic.json")


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * mult[unit]


sys.path.insert(0, HERE)
from kernel_stamp import csrc_sha16, sass_sha16_for       # the stamps bench.py compares (profiles/kernel_stamp.py)


def main():
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    args = sys.argv[1:]
    for k in range(0, len(args), 3):
        key, rep, desc = args[k], args[k + 1], args[k + 2]
        # REPORT: a .ncu-rep, or the CSV `ncu -i REPORT --page raw --csv` printed (profiles/refresh_traffic.sh leaves those in gpurun_out/)
        raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        rd = wr = inst = 0.0
        per_launch = []
        for vals in rows[2:]:
            if len(vals) != len(hdr):
                continue
            r = to_bytes(vals[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            w = to_bytes(vals[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            if r != r or w != w:          # a launch too short for the DRAM counters (e.g. the empty retry pass)
                r = w = 0.0
            rd += r; wr += w
            try:
                v = float(vals[col["smsp__inst_executed.sum"]].replace(",", ""))
                if v == v:
                    inst += v
            except Exception:
                pass
            per_launch.append({"kernel": vals[col["Kernel Name"]][:40], "dram_bytes": int(r + w),
                               "duration_under_ncu": vals[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]})
        data[key] = {"dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr), "launches": per_launch,
                     "warp_instructions_per_step": int(inst), "csrc_sha16": csrc_sha16(), "sass_sha16": sass_sha16_for(key), "captured_with": desc, "report": os.path.basename(rep)}
        print(key, data[key])
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
