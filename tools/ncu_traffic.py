#!/usr/bin/env python
"""Per-launch DRAM traffic of each kernel in an .ncu-rep (dram__bytes_read.sum + dram__bytes_write.sum,
averaged over the captured launches) -> JSON fragment for profiles/r01_traffic.json.
usage: python tools/ncu_traffic.py gpurun_out/x.ncu-rep workload >> merged by hand / tools/run_profiles.sh"""
import csv
import io
import json
import subprocess
import sys

rep, workload = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
acc = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = col[key]
        tot += float(r[i].replace(",", "")) * scale[units[i]]
    if tot != tot:                      # ncu sometimes fails to collect a launch (nan): skip it
        continue
    a = acc.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += tot
    a[2] += float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[col["gpu__time_duration.sum"]], 1e-6)
out = {workload: {k: v[1] / v[0] for k, v in acc.items()},
       workload + "_detail": {k: {"launches_captured": v[0], "dram_bytes_per_launch": v[1] / v[0], "ms_per_launch_under_ncu": v[2] / v[0]} for k, v in acc.items()}}
print(json.dumps(out, indent=1))
