"""The scene BVH built on the device (csrc/gpu_bvh.cu) -- replaces the host build of Job_ColInfo_Inner /
AABBTree::SetAABBs (lighter.cpp:349-384, lighter_math.cpp:674-781) for every hot query.

Two kinds of evidence: (1) the device builder runs the same binned SAH as the host builder (csrc/bvh.cpp), so on
non-degenerate input the two trees must be EQUAL record by record; (2) on any input -- degenerate included -- every
query on the device-built tree equals brute force over the triangles (the oracle), which is all the bake needs
(SURVEY finding 3: the topology is free)."""
import numpy as np
import pytest

from conftest import bits_equal, scene_tris
from lighter_b200 import api, scenes

pytestmark = pytest.mark.gpu


def _soup(n, seed, spread=6.0, size=0.5):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-spread, spread, (n, 1, 3)) + rng.uniform(-size, size, (n, 3, 3))).astype(np.float32).reshape(n, 9)


@pytest.mark.parametrize("n,leaf", [(3, 2), (33, 2), (34, 2), (100, 1), (5000, 2), (5000, 4), (70000, 2), (200000, 7)])
def test_device_tree_equals_host_tree_on_random_soups(n, leaf):
    r = api.test_device_bvh(_soup(n, 1000 + n + leaf), leaf)
    assert (r["mismatch_nodes"], r["mismatch_nodes4"], r["mismatch_leaves"]) == (0, 0, 0), r
    assert r["height"] == r["host_height"] and r["n_nodes"] > 0


@pytest.mark.parametrize("name", ["mesh2", "config3_sibling", "config4", "config5"])
def test_device_tree_equals_host_tree_on_the_workloads(name):
    sc = scenes.NAMED[name]() if name in scenes.NAMED else scenes.workload(name)
    tris = scene_tris(sc)
    r = api.test_device_bvh(tris, 2)
    print(name, len(tris), r)
    assert (r["mismatch_nodes"], r["mismatch_nodes4"], r["mismatch_leaves"]) == (0, 0, 0), r
    assert r["height"] == r["host_height"] <= 40
    if len(tris) > 500_000:
        assert r["build_ms"] < 30.0, r          # the host build this replaces took 70-96 ms on 16 cores


def _check_queries(tris, oracle, seed):
    rng = np.random.default_rng(seed)
    n = 400
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    a = rng.uniform(lo - 0.5, hi + 0.5, (n, 3)).astype(np.float32)
    b = (a + rng.normal(0, 0.3 * float(np.max(hi - lo)) + 0.1, (n, 3))).astype(np.float32)
    q = api.test_scene_queries(tris, a, b)
    assert bits_equal(q["dist"], oracle.scene_distance(tris, a))
    assert np.array_equal(q["anyhit"], oracle.anyhit_raw(tris, a, b))
    c, tid = oracle.closest_raw(tris, a, b)
    assert bits_equal(q["closest"], c) and np.array_equal(q["closest_tri"], tid)


def test_degenerate_geometry_builds_a_bounded_tree_and_answers_like_brute_force(oracle):
    """Coincident triangles (no SAH split exists: position splits), a chain of slivers with exponentially growing gaps
    (every SAH split peels one triangle off: the height limit must kick in)."""
    one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    same = np.repeat(one, 3000, axis=0)
    r = api.test_device_bvh(same, 2)
    assert r["height"] <= 40 and r["n_nodes"] >= 1499
    _check_queries(same, oracle, 1)
    k = np.arange(60, dtype=np.float64)
    x = (1.5 ** k).astype(np.float32)
    chain = np.stack([x, 0 * x, 0 * x, x + 1e-3, 0 * x, 0 * x + 1, x, 0 * x + 1, 0 * x], 1).astype(np.float32)
    chain = np.concatenate([chain, np.repeat(one, 200, axis=0)])
    r = api.test_device_bvh(chain, 1)
    assert r["height"] <= 40, r
    _check_queries(chain, oracle, 2)


def test_position_split_fallback_is_still_exact(oracle, monkeypatch):
    """LTR_BVH_SAH_DEPTH=3: below three SAH levels every node splits by position -- a poor tree, the same answers."""
    monkeypatch.setenv("LTR_BVH_SAH_DEPTH", "3")
    tris = _soup(4000, 5)
    r = api.test_device_bvh(tris, 2)
    assert r["height"] <= 3 + 12
    _check_queries(tris, oracle, 4)


def test_bake_is_identical_with_host_and_device_built_trees(monkeypatch):
    """A/B switch LTR_BVH_HOST=1: same lightmaps bit for bit (the trees are equal; even if they were not, results may not change)."""
    sc = scenes.workload("config4_sibling")
    a = api.bake(sc)
    monkeypatch.setenv("LTR_BVH_HOST", "1")
    b = api.bake(sc)
    for x, y in zip(a["lightmaps"], b["lightmaps"]):
        assert bits_equal(x["rgb"], y["rgb"])
    assert a["stats"]["n_bvh_nodes"] == b["stats"]["n_bvh_nodes"] > 0
