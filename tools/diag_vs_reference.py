#!/usr/bin/env python
"""Stage-by-stage comparison of a GPU bake with the live reference (oracle/_ref) on one workload: lumel positions,
per-lumel colours, radiosity link set, final texels.  Diagnostic for parity work:  python tools/diag_vs_reference.py NAME"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, parity, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config4_sibling"
sc = scenes.NAMED[name]() if name in scenes.NAMED else scenes.workload(name)
ref = parity.run_reference(sc, threads=1, internals=True)
out = api.bake(sc, debug=True)
print(name, "env", {k: v for k, v in os.environ.items() if k.startswith("LTR_")})
for k, (a, b) in enumerate(zip(out["instances"], ref["instances"])):
    if a["n"] != b["n"]:
        print("inst", k, "lumel count", a["n"], b["n"]); continue
    if not a["n"]:
        continue
    dp = (a["pos"].view(np.uint32) != b["pos"].view(np.uint32)).any(axis=1)
    dc = np.abs(a["rgb"] - b["rgb"]).max(axis=1)
    print(f"inst {k}: {a['n']} lumels, positions differing {int(dp.sum())}, colour max abs diff {dc.max():.3g}, >1e-5: {int((dc > 1e-5).sum())}, >1e-3: {int((dc > 1e-3).sum())}")
    if dp.any():
        i = np.nonzero(dp)[0][:5]
        for j in i:
            print("   lumel", j, "gpu", a["pos"][j], "ref", b["pos"][j], "nrm", b["nrm"][j])
lk = out["links"]
if lk["rows"]:
    rows = np.repeat(np.arange(lk["rows"], dtype=np.uint32), np.diff(lk["row_offset"]).astype(np.int64))
    fwd = rows < lk["other"]
    lm, oth, fac = ref["links"]["map"], ref["links"]["other"], ref["links"]["factor"]
    ri = np.repeat(np.arange(len(lm), dtype=np.uint32), lm[:, 1])
    key = lambda x, y: x.astype(np.uint64) << np.uint64(32) | y.astype(np.uint64)
    kg, kr = key(rows[fwd], lk["other"][fwd]), key(ri, oth)
    sg, sr = set(kg.tolist()), set(kr.tolist())
    print("links gpu", len(kg), "ref", len(kr), "only gpu", len(sg - sr), "only ref", len(sr - sg))
    for k in sorted(sg ^ sr)[:8]:
        print("   pair", k >> 32, k & 0xffffffff, "gpu" if k in sg else "ref")
for a, b in zip(out["lightmaps"], ref["lightmaps"]):
    p = parity.texel_parity(a["rgb"], b["rgb"])
    print("lightmap", a["uid"], {k: v for k, v in p.items()})
st = out["stats"]
print({k: st[k] for k in ("n_marches", "n_distance_queries", "n_ao_segments", "n_rad_segments", "n_rad_links", "n_correction_rays")})
