#!/usr/bin/env python3
"""Pull dram__bytes_read.sum + dram__bytes_write.sum (one launch, `ncu --set full`) out of .ncu-rep files and record them in
profiles/traffic.json, which bench.py quotes as roofline.traffic when it runs the same workload.
usage: extract_traffic.py KEY REPORT.ncu-rep "description of the captured command" [KEY REPORT DESC ...]"""
import csv, io, json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "traffic.json")


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * mult[unit]


def main():
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    args = sys.argv[1:]
    for k in range(0, len(args), 3):
        key, rep, desc = args[k], args[k + 1], args[k + 2]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        col = {h: i for i, h in enumerate(hdr)}
        rd = to_bytes(vals[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(vals[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        dur = vals[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]
        data[key] = {"dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr), "kernel": vals[col["Kernel Name"]][:60],
                     "duration_under_ncu": dur, "captured_with": desc, "report": os.path.basename(rep)}
        print(key, data[key])
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
