#!/usr/bin/env python
"""Bake named workloads on the GPU and print stage timings / counters; optionally compare with the
live reference (`--ref`, only sensible for the small siblings).  Run under gpurun."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, parity, scenes  # noqa: E402

args = sys.argv[1:]
with_ref = "--ref" in args
repeat = 2 if "--twice" in args else 1
names = [a for a in args if not a.startswith("--")]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
report = {}
for name in names:
    t0 = time.time()
    sc = scenes.workload(name)
    t_gen = time.time() - t0
    r = dict(tris=sc.triangle_count(), instances=len(sc.instances), lights=len(sc.lights), t_scene_gen=t_gen)
    for it in range(repeat):
        out = api.bake(sc, debug=with_ref)
        st = out["stats"]
        r[f"wall_s_{it}"] = out["wall_s"]
        r[f"stats_{it}"] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items() if v}
        rays = st["n_distance_queries"] + st["n_ao_segments"] + st["n_rad_segments"] + st["n_correction_rays"]
        r[f"rays_{it}"] = rays
        r[f"rays_per_s_wall_{it}"] = rays / out["wall_s"]
    r["lightmap_means"] = [float(lm["rgb"].mean()) for lm in out["lightmaps"][:4]]
    if with_ref:
        ref = parity.run_reference(sc, threads=1, internals=True)
        r["ref_wall_1thr"] = ref["wall_s"]
        r["lumels_equal"] = [a["n"] for a in out["instances"]] == [b["n"] for b in ref["instances"]]
        r["pos_bitexact"] = [float((a["pos"].view(np.uint32) == b["pos"].view(np.uint32)).all(axis=1).mean()) for a, b in zip(out["instances"], ref["instances"]) if a["n"] and a["n"] == b["n"]]
        r["texels"] = [parity.texel_parity(a["rgb"], b["rgb"]) for a, b in zip(out["lightmaps"], ref["lightmaps"])]
        if ref["links"] is not None:
            r["links"] = (len(ref["links"]["other"]), int(len(out["links"]["other"])))
    report[name] = r
    print(name, json.dumps(r, indent=1, default=str), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_probe.json"), "w") as f:
        json.dump(report, f, indent=1, default=str)
