#!/usr/bin/env python
"""Count the rays the UNMODIFIED reference casts on the CPU-runnable workloads, with the counting
build oracle/_ref/ref_bake_count (reference sources compiled with -finstrument-functions; see
oracle/bake_driver.cpp).  Writes tests/golden/ref_counts.json, which bench.py's reference arm uses to
turn the reference's wall time into rays/s and which tests/test_gpu_parity.py compares with the GPU
counters.  Run in the build container:  python tools/count_reference_rays.py
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import scenes  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "ref_bake_count")
OUT = os.path.join(ROOT, "tests", "golden", "ref_counts.json")

if __name__ == "__main__":
    res = {}
    names = sys.argv[1:] or ["basic", "mesh1", "rad1", "mesh2", "config3_sibling", "config4_sibling"]
    with tempfile.TemporaryDirectory() as td:
        for name in names:
            sc = scenes.NAMED[name]() if name in scenes.NAMED else scenes.workload(name)
            sp = sc.write(os.path.join(td, name + ".scn"))
            r = subprocess.run([EXE, sp, os.path.join(td, "o.bin"), "--threads", "1", "--quiet"], capture_output=True, text=True, check=True)
            c = json.loads(r.stdout.strip().splitlines()[-1])
            c["rays"] = c["distance_queries"] + c["visibility_segments"] + c["ao_segments"] + c["correction_rays"]
            c["triangles"] = sc.triangle_count()
            res[name] = c
            print(name, c)
    if os.path.exists(OUT) and sys.argv[1:]:
        old = json.load(open(OUT)); old.update(res); res = old
    json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
