#!/usr/bin/env python3
"""Per-kernel SASS hashes of snap_b200/csrc/libsnapgpu.so (cuobjdump -sass), written to / compared with a JSON file.
usage: compare_sass.py --write FILE | --check FILE
The last full GPU test run of the round (profiles/r02_final_pytest_gpu.log) was of the build recorded in profiles/r02_final_kernel_sass.json; later commits
changed host code only, and `--check` shows that every kernel of the current build is bit-identical to the tested one."""
import hashlib, json, os, subprocess, sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "snap_b200", "csrc", "libsnapgpu.so")


def funcs(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d, name = {}, None
    for line in out.split("\n"):
        s = line.strip()
        if s.startswith("Function :"):
            name = s.split(":", 1)[1].strip(); d[name] = hashlib.sha256()
        elif name and s.startswith("/*"):
            d[name].update(s.encode())
    return {k: v.hexdigest()[:16] for k, v in d.items()}


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    now = funcs(LIB)
    if mode == "--write":
        json.dump(now, open(path, "w"), indent=1, sort_keys=True)
        print("wrote", len(now), "kernels")
    else:
        ref = json.load(open(path))
        diff = sorted(k for k in set(ref) | set(now) if ref.get(k) != now.get(k))
        print("%d kernels, %d differ%s" % (len(now), len(diff), ": " + ", ".join(d[:50] for d in diff) if diff else ""))
        sys.exit(1 if diff else 0)
