#!/usr/bin/env python
"""First-contact GPU script: bake the reference's own scenarios on the GPU, run the unmodified
reference (oracle/_ref) on the host beside it, and print stage-level parity.  Run under gpurun."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, parity, scenes  # noqa: E402

names = sys.argv[1:] or ["basic", "mesh1", "rad1", "mesh2", "hugeoverlap"]
report = {}
for name in names:
    sc = scenes.NAMED[name]()
    t0 = time.time()
    ref = parity.run_reference(sc, threads=1, internals=True)
    t_ref = time.time() - t0
    ours = api.bake(sc, debug=True)
    r = dict(ref_wall=ref["wall_s"], ours_wall=ours["wall_s"], stats={k: v for k, v in ours["stats"].items() if v}, stages=ours["stages"])
    r["lumels"] = []
    for i, (a, b) in enumerate(zip(ours["instances"], ref["instances"])):
        d = dict(n_ours=a["n"], n_ref=b["n"])
        if a["n"] == b["n"] and a["n"]:
            d["loc_equal"] = bool(np.array_equal(a["loc"], b["loc"]))
            d["pos_bitexact"] = float((a["pos"].view(np.uint32) == b["pos"].view(np.uint32)).all(axis=1).mean())
            d["pos_maxabs"] = float(np.abs(a["pos"] - b["pos"]).max())
            d["nrm_bitexact"] = float((a["nrm"].view(np.uint32) == b["nrm"].view(np.uint32)).all(axis=1).mean())
            d["rad_bitexact"] = float((a["radinfo"].view(np.uint32) == b["radinfo"].view(np.uint32)).all(axis=1).mean())
            d["rgb_bitexact"] = float((a["rgb"].view(np.uint32) == b["rgb"].view(np.uint32)).all(axis=1).mean())
            d["rgb_maxabs"] = float(np.abs(a["rgb"] - b["rgb"]).max())
        r["lumels"].append(d)
    r["lightmaps"] = []
    for a, b in zip(ours["lightmaps"], ref["lightmaps"]):
        d = dict(uid=(a["uid"], b["uid"]), size=(a["width"], a["height"], b["width"], b["height"]))
        if a["rgb"].shape == b["rgb"].shape:
            d.update(parity.texel_parity(a["rgb"], b["rgb"]))
            if a["normals"] is not None and b["normals"] is not None:
                d["normals_maxabs"] = float(np.abs(a["normals"] - b["normals"]).max())
        r["lightmaps"].append(d)
    if ref["links"] is not None and len(ref["links"]["other"]):
        r["links_ref"] = int(len(ref["links"]["other"]))
        r["links_ours_directed"] = int(len(ours["links"]["other"]))
    report[name] = r
    print(name, json.dumps(r, indent=1, default=str))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_first.json"), "w") as f:
    json.dump(report, f, indent=1, default=str)
