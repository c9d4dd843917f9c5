#!/usr/bin/env python
"""Which shadow factors differ between two builds of the library on a full-size workload, and what does brute force say?
  python tools/march_variant_diff.py dump  config4 /tmp/fv_a.npz        (run once per build)
  python tools/march_variant_diff.py check config4 /tmp/fv_a.npz /tmp/fv_b.npz
The second form recomputes every differing (lumel, light) factor with the oracle's brute-force march over the triangles within
2.05 of the march segment (what the reference's own pruning can see at most) and over ALL triangles."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from lighter_b200 import api, scenes  # noqa: E402

mode, name = sys.argv[1], sys.argv[2]
sc = scenes.workload(name)
if mode == "dump":
    with api.BakeHandle(sc, debug=True) as h:
        h.run()
        insts = [h.lumels(i) for i in range(len(sc.instances) + 1)]
        fv = np.stack([h.shadow_factors(l) for l in range(len(sc.lights))])
    pos = np.concatenate([i["pos"] for i in insts if i["n"]]); nrm = np.concatenate([i["nrm"] for i in insts if i["n"]])
    np.savez(sys.argv[3], fv=fv, pos=pos, nrm=nrm)
    print("dumped", fv.shape)
else:
    from conftest import scene_tris
    from oracle import Oracle
    orc = Oracle()
    a, b = np.load(sys.argv[3]), np.load(sys.argv[4])
    fa, fb = a["fv"], b["fv"]
    diff = np.argwhere(fa.view(np.uint32) != fb.view(np.uint32))
    print(f"{name}: {len(diff)} of {fa.size} factors differ between the two builds")
    tris = scene_tris(sc)
    t3 = tris.reshape(-1, 3, 3); tlo, thi = t3.min(1), t3.max(1)
    fp = C.POINTER(C.c_float)
    agree = {"a_near": 0, "b_near": 0, "a_all": 0, "b_all": 0, "n": 0}
    for l, g in diff[:40]:
        lt = sc.lights[l]
        P, N = a["pos"][g].astype(np.float32), a["nrm"][g].astype(np.float32)
        pl = orc.pack_light(lt)
        frm = P + N * np.float32(0.005)
        to = (P + pl[4:7] * np.float32(lt.range)) if lt.type == 3 else np.asarray(lt.position, np.float32)
        lo, hi = np.minimum(frm, to) - np.float32(2.05), np.maximum(frm, to) + np.float32(2.05)
        near = np.ascontiguousarray(tris[((thi >= lo) & (tlo <= hi)).all(1)])
        res = {}
        for tag, tt in (("near", near), ("all", tris)):
            rgb, fv = np.zeros(3, np.float32), C.c_float(np.nan)
            orc.L.o_direct_lumel(tt.ctypes.data_as(fp), len(tt), pl.ctypes.data_as(fp), np.ascontiguousarray(P).ctypes.data_as(fp),
                                 np.ascontiguousarray(N).ctypes.data_as(fp), rgb.ctypes.data_as(fp), C.byref(fv))
            res[tag] = np.float32(fv.value)
            agree["a_" + tag] += int(res[tag].view(np.uint32) == fa[l, g].view(np.uint32))
            agree["b_" + tag] += int(res[tag].view(np.uint32) == fb[l, g].view(np.uint32))
        agree["n"] += 1
        print(f"  light {l} lumel {g}: a {fa[l, g]:.9g}  b {fb[l, g]:.9g}  brute(near) {res['near']:.9g}  brute(all) {res['all']:.9g}")
    print("agreement with brute force over", agree["n"], "differing factors:", agree)
