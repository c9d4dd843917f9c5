#!/usr/bin/env python
"""sample_fn batching at size (ref lighter.cpp:690-715): end-to-end bakes of a workload with and without the
library's native example callback (ltrx_SampleFnChecker, one call per mesh lumel, serial, reference order).
    python tools/sample_fn_timing.py [config4] [bakes]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config4"
bakes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
res = {}
for kind in (0, 2):
    sc = scenes.workload(name)
    sc.cfg["sample_fn_kind"] = kind
    walls, st, hsh = [], None, ""
    for it in range(bakes + 1):
        with api.BakeHandle(sc) as h:
            w = h.run()
            if it:
                walls.append(w)
                st = h.stats()
                hsh = h.output_hash()
    res["with_callback" if kind else "no_callback"] = dict(
        wall_s=sum(walls) / len(walls), lumels=st["n_lumels_total"], t_sample_fn_s=st["t_sample_fn"], t_direct_s=st["t_direct"], t_radiosity_s=st["t_radiosity"],
        gpu_ms_span=st["gpu_ms_span"], d2h_bytes=st["d2h_bytes"], h2d_bytes=st["h2d_bytes"], lightmap_fnv1a64=hsh)
a, b = res["no_callback"], res["with_callback"]
res["summary"] = dict(workload=name, callbacks=b["lumels"], callback_thread_s=b["t_sample_fn_s"],
                      ns_per_callback_incl_request=b["t_sample_fn_s"] / max(b["lumels"], 1) * 1e9,
                      added_wall_s=b["wall_s"] - a["wall_s"],
                      note="the callback thread runs beside direct light + link generation; added_wall_s is what is NOT hidden (plus the 24 B/lumel material upload)")
print(json.dumps(res))
