#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the UNMODIFIED reference
(oracle/_ref, compiled from /root/reference by `make -C oracle ref`).  Run in the build container:

    python tools/make_golden.py

prims.npz  seeded random inputs + the reference's own outputs for every primitive on the path
bakes.npz  the reference's scenario bakes (1 thread => deterministic AO): float lightmaps, per-lumel
           arrays and radiosity links for the small scenes, 8-bit lightmaps for mesh2
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from lighter_b200 import parity, scenes  # noqa: E402
from oracle import RefPrims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def random_tris(rng, n, extent=4.0, size=1.0):
    c = rng.uniform(-extent, extent, (n, 1, 3))
    return (c + rng.uniform(-size, size, (n, 3, 3))).astype(np.float32).reshape(n, 9)


def make_prims():
    R = RefPrims()
    rng = np.random.default_rng(20261017)
    out = {}
    # point/triangle distance: generic + degenerate-ish + points on/near the triangle
    n = 4096
    tris = random_tris(rng, n)
    pts = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    near = (tris.reshape(n, 3, 3) * rng.dirichlet((1, 1, 1), n)[:, :, None].astype(np.float32)).sum(1) + rng.normal(0, 0.01, (n, 3)).astype(np.float32)
    pts[::2] = near[::2]
    out["ptd_pts"], out["ptd_tris"] = pts, tris
    out["ptd_out"] = R.point_tri_distance(pts, tris)
    # segment/triangle
    a = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    b = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    through = (tris.reshape(n, 3, 3) * rng.dirichlet((1, 1, 1), n)[:, :, None].astype(np.float32)).sum(1)
    b[::2] = (a[::2] + (through[::2] - a[::2]) * rng.uniform(0.5, 2.0, (n // 2, 1))).astype(np.float32)
    out["seg_a"], out["seg_b"], out["seg_tris"] = a, b, tris
    out["seg_out"] = R.seg_tri(a, b, tris)
    # spiral directions
    m = 256
    nrm = rng.normal(0, 1, (m, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    nrm[0] = (0, 0, 1); nrm[1] = (1, 0, 0); nrm[2] = (0, -1, 0)
    ro = rng.uniform(0, 1, m).astype(np.float32)
    out["spiral_nrm"], out["spiral_randoff"] = nrm, ro
    out["spiral_out17"] = R.spiral_dirs(nrm, ro, 17)
    # scene queries against a triangle soup (reference tree == brute force, SURVEY finding 3)
    soup = random_tris(rng, 300, extent=3.0, size=0.8)
    qa = rng.uniform(-4, 4, (2048, 3)).astype(np.float32)
    qb = rng.uniform(-4, 4, (2048, 3)).astype(np.float32)
    q = R.tree_queries(soup, qa, qb)
    out["soup"], out["q_a"], out["q_b"] = soup, qa, qb
    out["q_dist"], out["q_anyhit"], out["q_closest"], out["q_closest_tri"] = q["dist"], q["anyhit"], q["closest"], q["closest_tri"]
    # rasteriser: a few triangles into one 32x32 image, both margins, in order
    w = h = 32
    imgs = (np.zeros((h, w, 3), np.float32), np.zeros((h, w, 3), np.float32), np.zeros((h, w, 4), np.float32))
    rp = rng.uniform(-2, 34, (6, 6)).astype(np.float32)
    rva = rng.uniform(-1, 1, (6, 9)).astype(np.float32)
    rvb = rng.normal(0, 1, (6, 9)).astype(np.float32)
    rvc = rng.uniform(0, 1, (6, 12)).astype(np.float32)
    for margin in (np.float32(0.5) + np.float32(0.001), np.float32(0.0)):
        for k in range(6):
            R.raster_tri(w, h, margin, rp[k], rva[k], rvb[k], rvc[k], imgs)
    out["raster_p"], out["raster_va"], out["raster_vb"], out["raster_vc"] = rp, rva, rvb, rvc
    out["raster_img1"], out["raster_img2"], out["raster_img3"] = imgs
    # blur / downsample
    img = rng.uniform(0, 1, (24, 40, 3)).astype(np.float32)
    out["post_img"] = img
    out["post_blur05"] = R.blur(img, 0.5)
    out["post_blur22"] = R.blur(img, 2.2)
    out["post_ds2x"] = R.downsample2x(img)
    np.savez_compressed(os.path.join(GOLD, "prims.npz"), **out)
    print("prims.npz:", {k: v.shape for k, v in out.items()})


def make_bakes():
    out = {}
    for name in ("basic", "hugeoverlap", "mesh1", "rad1", "mesh2"):
        sc = scenes.NAMED[name]()
        ref = parity.run_reference(sc, threads=1, internals=True)
        out[f"{name}_lumel_counts"] = np.array([i["n"] for i in ref["instances"]], np.int64)
        for lm in ref["lightmaps"]:
            if name == "mesh2":
                out[f"{name}_lm{lm['uid']}_q8"] = parity.quantize8(lm["rgb"]).astype(np.uint8)
                out[f"{name}_lm{lm['uid']}_nq8"] = parity.quantize8(lm["normals"] * 0.5 + 0.5).astype(np.uint8)
            else:
                out[f"{name}_lm{lm['uid']}_rgb"] = lm["rgb"]
        if name in ("basic", "mesh1", "rad1"):
            for i, inst in enumerate(ref["instances"]):
                if not inst["n"]:
                    continue
                for k in ("pos", "nrm", "loc", "radinfo", "rgb"):
                    out[f"{name}_inst{i}_{k}"] = inst[k]
        if name == "rad1":
            lk = ref["links"]
            rows = np.repeat(np.arange(len(lk["map"]), dtype=np.uint32), lk["map"][:, 1])
            out["rad1_link_i"], out["rad1_link_j"], out["rad1_link_f"] = rows, lk["other"], lk["factor"]
        print(name, "lumels", out[f"{name}_lumel_counts"].tolist(), "ref wall %.3fs" % ref["wall_s"])
    np.savez_compressed(os.path.join(GOLD, "bakes.npz"), **out)


if __name__ == "__main__":
    make_prims()
    make_bakes()
    for f in ("prims.npz", "bakes.npz"):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")
