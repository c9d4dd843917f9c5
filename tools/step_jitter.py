#!/usr/bin/env python
"""Step-to-step jitter of resident bakes: N bakes of a workload on one resident scene, per step the CUDA-event span, the host
wall and the host-side stage times (ltrx_GetStats); then the same again with LTR_TRACE=1 (the library's phase laps on stderr).
    python tools/step_jitter.py [workload] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
keys = ("t_samples", "t_direct", "t_radiosity", "t_ao", "t_finalize")
gk = ("gpu_ms_samples", "gpu_ms_direct", "gpu_ms_radiosity", "gpu_ms_ao", "gpu_ms_finalize")
sc = scenes.workload(name)
with api.BakeHandle(sc) as h:
    h.prepare()
    for _ in range(3):
        h.bake_resident()
    for phase in ("plain", "trace"):
        if phase == "trace":
            os.environ["LTR_TRACE"] = "1"
        for i in range(steps):
            t = time.perf_counter()
            if phase == "trace":
                sys.stderr.write("---- step %d ----\n" % i)
                sys.stderr.flush()
            ms = h.bake_resident()
            wall = (time.perf_counter() - t) * 1e3
            st = h.stats()
            print("%s step %2d: span %7.1f ms  wall %7.1f | host " % (phase, i, ms, wall) + " ".join("%s %6.1f" % (k[2:], st[k] * 1e3) for k in keys)
                  + " | gpu " + " ".join("%s %6.1f" % (k[7:], st[k]) for k in gk) + " | batches %d" % st["n_rad_batches"], flush=True)
