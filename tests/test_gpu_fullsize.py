"""Parity at FULL size (BASELINE configs[3]: 1.0 M triangles, 7.76 M lumels), where the reference itself cannot finish
(its link loop alone is ~1e6 s): a seeded sample of lumels of the full bake is recomputed by the oracle's brute-force
restatement and compared bit for bit.

What makes brute force affordable is only WHICH triangles / lumels are handed to the oracle, and each restriction is exact:
  * a shadow march only ever asks for distances capped at 2 (MAX_PENUMBRA_SIZE, lighter.cpp:150-188), so triangles whose box is
    farther than 2 from the march segment cannot change any step;
  * a radiosity pair needs factor >= 0.001, i.e. a distance <= sqrt(1 / (0.001 pi)) = 17.84 (lighter.cpp:746-749), and its
    visibility test only sees triangles the segment can touch.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal, scene_tris
from lighter_b200 import api, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,min_tris,min_lumels,n_picks", [("config4", 1_000_000, 7_000_000, 160), ("config3", 240_000, 1_000_000, 24), ("config5", 490_000, 500_000, 12)])
def test_full_size_sampled_against_oracle(oracle, name, min_tris, min_lumels, n_picks):
    """config4: 1.0 M triangles, 256 instances, 9 lights, radiosity.  config3: closed interior, 32 lights (march dominated).
    config5: ONE merged 500 k-triangle instance, 64 lights.  Sampled lumels x every light: shadow factors; config4: link rows."""
    sc = scenes.workload(name)
    tris = scene_tris(sc)
    assert len(tris) > min_tris
    t3 = tris.reshape(-1, 3, 3)
    tlo, thi = t3.min(1), t3.max(1)
    with api.BakeHandle(sc, debug=True) as h:
        h.run()
        st = h.stats()
        insts = [h.lumels(i) for i in range(len(sc.instances) + 1)]
        lk = h.links()
        fvis = [h.shadow_factors(l) for l in range(len(sc.lights))]
    pos = np.concatenate([i["pos"] for i in insts if i["n"]]).astype(np.float32)
    nrm = np.concatenate([i["nrm"] for i in insts if i["n"]]).astype(np.float32)
    n = len(pos)
    assert n == st["n_lumels_total"] > min_lumels
    rng = np.random.default_rng(11)
    picks = np.sort(rng.choice(n, n_picks, replace=False))

    # ---- shadow-march factors of the sampled lumels, every light (lighter.cpp:485-600, 190-207) ------------------
    fp = C.POINTER(C.c_float)
    compared = 0
    for g in picks:
        P, N = pos[g], nrm[g]
        frm = P + N * np.float32(0.005)
        for l, lt in enumerate(sc.lights):
            pl = oracle.pack_light(lt)
            to = (P + pl[4:7] * np.float32(lt.range)) if lt.type == 3 else np.asarray(lt.position, np.float32)
            lo, hi = np.minimum(frm, to) - np.float32(2.05), np.maximum(frm, to) + np.float32(2.05)
            near = np.ascontiguousarray(tris[((thi >= lo) & (tlo <= hi)).all(1)])
            rgb, fv = np.zeros(3, np.float32), C.c_float(np.nan)
            lit = oracle.L.o_direct_lumel(near.ctypes.data_as(fp), len(near), pl.ctypes.data_as(fp), np.ascontiguousarray(P).ctypes.data_as(fp),
                                          np.ascontiguousarray(N).ctypes.data_as(fp), rgb.ctypes.data_as(fp), C.byref(fv))
            if not lit or (lt.type == 3 and not (rgb > 0).any() and float(np.dot(pl[4:7], N)) <= 0):
                continue                      # the reference early-outs / the GPU skips marches whose result is multiplied by 0
            assert np.float32(fv.value).view(np.uint32) == fvis[l][g].view(np.uint32), (int(g), l, fv.value, float(fvis[l][g]))
            compared += 1
    assert compared >= (120 if name == "config4" else 60)

    # ---- link rows of the sampled lumels: partners and factors (lighter.cpp:728-760) --------------------------
    if not sc.cfg["bounce_count"]:
        return
    assert lk["rows"] == n and len(lk["other"]) >= st["n_rad_links"]
    ro, other, fac = lk["row_offset"], lk["other"], lk["factor"]
    total_links = total_segments = 0
    for g in picks:
        near = np.nonzero((np.abs(pos - pos[g]) <= np.float32(18.0)).all(1))[0]
        near = near[near != g]
        lo, hi = pos[g] - np.float32(18.1), pos[g] + np.float32(18.1)
        tsel = np.ascontiguousarray(tris[((thi >= lo) & (tlo <= hi)).all(1)])
        linked, f, sg = oracle.rad_row(tsel, pos[g], nrm[g], pos[near], nrm[near], int(np.searchsorted(near, g)))
        row, rf = other[ro[g]:ro[g + 1]], fac[ro[g]:ro[g + 1]]
        assert np.array_equal(row, near[linked].astype(np.uint32)), (int(g), len(row), int(linked.sum()))
        assert bits_equal(rf, f[linked]), int(g)
        total_links += int(linked.sum()); total_segments += sg
    assert total_segments > 40_000 and total_links > 200
