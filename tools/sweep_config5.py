#!/usr/bin/env python
"""BASELINE configs[4]: light-count sweep 1-1024 lights x 1-64 soft-shadow samples on the 500k-triangle,
1024^2-texel scene (one GPU).  Two units, never mixed:

  march mode   (reference semantics; shadow_sample_count is ignored, finding 1): marches/s and distance
               queries/s versus the light count;
  sampled mode (extension, ltrx_SetShadowMode): shadow rays/s = lumel x light x sample any-hit rays per
               second of the ray kernel, versus lights x samples.

  python tools/sweep_config5.py [--lights 1,4,16,64,256,1024] [--samples 1,4,16,64] > gpurun_out/r01_sweep_config5.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lights", default="1,4,16,64,256,1024")
ap.add_argument("--samples", default="1,4,16,64")
ap.add_argument("--tris", type=int, default=500_000)
args = ap.parse_args()
rows = []
for L in [int(x) for x in args.lights.split(",")]:
    sc = scenes.scene_config5(n_lights=L, samples=1, target_tris=args.tris)
    with api.BakeHandle(sc) as h:
        h.prepare()
        h.bake_resident()
        ms = h.bake_resident()
        st = h.stats()
    rows.append(dict(mode="march", lights=L, samples=None, lumels=st["n_lumels_total"], triangles=st["n_triangles"], step_ms=ms,
                     kernel_ms=st["gpu_ms_march"], marches=st["n_marches"], distance_queries=st["n_distance_queries"],
                     marches_per_s=st["n_marches"] / (st["gpu_ms_march"] * 1e-3) if st["gpu_ms_march"] else 0.0,
                     distance_queries_per_s=st["n_distance_queries"] / (st["gpu_ms_march"] * 1e-3) if st["gpu_ms_march"] else 0.0))
    print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    for S in [int(x) for x in args.samples.split(",")]:
        for lt in sc.lights:
            lt.shadow_sample_count = S
        with api.BakeHandle(sc, shadow_mode=1) as h:
            h.prepare()
            h.bake_resident()
            ms = h.bake_resident()
            st = h.stats()
        rows.append(dict(mode="sampled", lights=L, samples=S, lumels=st["n_lumels_total"], triangles=st["n_triangles"], step_ms=ms,
                         kernel_ms=st["gpu_ms_march"], shadow_rays=st["n_shadow_rays"],
                         shadow_rays_per_s=st["n_shadow_rays"] / (st["gpu_ms_march"] * 1e-3) if st["gpu_ms_march"] else 0.0,
                         node_visits_per_ray=st["n_ray_node_visits"] / max(st["n_shadow_rays"], 1)))
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
print(json.dumps(dict(workload="config5 sweep: 500k-tri closed interior, 1024^2 texels, point/spot lights of range 14 on a grid", rows=rows), indent=1))
