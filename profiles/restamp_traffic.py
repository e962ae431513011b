#!/usr/bin/env python3
"""After an edit of snap_b200/csrc that leaves the alignment / lookup kernels' SASS as captured (a host-side function, another kernel), bring the
`csrc_sha16` stamps of profiles/traffic.json up to the current sources, entry by entry, ONLY where the entry's `sass_sha16` equals the stamp of the built
library (profiles/kernel_stamp.py).  bench.py accepts either stamp anyway; this merely spares it the `cuobjdump -sass` pass (9 s) on every run.
Entries whose kernels changed are left alone (bench.py then prints `traffic: null` and says why): they need profiles/refresh_traffic.sh on the GPU box."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from kernel_stamp import csrc_sha16, sass_sha16_for

path = os.path.join(HERE, "traffic.json")
t = json.load(open(path))
now = csrc_sha16()
for key, e in sorted(t.items()):
    sass = sass_sha16_for(key)
    if e.get("csrc_sha16") == now:
        print(key, "already at", now)
    elif sass and e.get("sass_sha16") == sass:
        e.setdefault("csrc_sha16_at_capture", e.get("csrc_sha16"))
        e["csrc_sha16"] = now
        e["restamped"] = "kernel SASS unchanged since the capture (sass_sha16 %s)" % sass
        print(key, "restamped to", now)
    else:
        print(key, "NOT restamped: SASS differs (captured %s, built %s)" % (e.get("sass_sha16"), sass))
json.dump(t, open(path, "w"), indent=1, sort_keys=True)
