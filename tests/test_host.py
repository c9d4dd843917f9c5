"""CPU tests of the host-side logic: shard partitioning (single process and a world_size-2 gloo
job), scene-file round trip through the C++ reader, the reference-order tree builder against the
reference's own tree dump, and the flat BVH builder's structural invariants."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from lighter_b200 import api, parity, scenes


def test_shard_ranges_tile_the_lumel_array():
    for n in (0, 1, 7, 8, 9, 1000, 16_777_216, 12_345_677):
        for world in (1, 2, 3, 4, 8):
            edges = [api.shard_range(n, r, world) for r in range(world)]
            chunk = (n + world - 1) // world
            assert edges[0][0] == 0 and edges[-1][1] == n
            for r, (b, e) in enumerate(edges):
                assert b == min(r * chunk, n) and e == min(b + chunk, n)       # equal-size chunks: in-place all-gather layout
                if r:
                    assert b == edges[r - 1][1]


def test_world_size_2_gloo_sharding(tmp_path):
    """Two CPU processes over gloo agree on a lumel count, take their ltrx_ShardRange slices, and
    all-gather padded per-rank payloads the way the radiance exchange lays them out."""
    script = tmp_path / "w2.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        from lighter_b200 import api
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n = 1001                                   # deliberately not divisible by the world size
        t = torch.tensor([n if rank == 0 else -1])
        dist.broadcast(t, 0)
        n = int(t.item())
        b, e = api.shard_range(n, rank, world)
        chunk = (n + world - 1) // world
        payload = torch.zeros(chunk, dtype=torch.float32)
        payload[: e - b] = torch.arange(b, e, dtype=torch.float32)       # "radiance" of my lumels
        gathered = [torch.zeros(chunk) for _ in range(world)]
        dist.all_gather(gathered, payload)
        full = torch.cat(gathered)[:n]
        assert torch.equal(full, torch.arange(n, dtype=torch.float32)), "gathered lumel array is not contiguous"
        # a 128-byte id travels from rank 0 to everyone, as the NCCL unique id does
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            ident = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(ident, 0)
        assert ident[127].item() == 127
        open(os.path.join({str(tmp_path)!r}, f"rank{{rank}}.txt"), "w").write(f"ok {{b}} {{e}}")
        dist.destroy_process_group()
    """))
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # per-rank files: the two ranks' stdout lines can interleave
    assert (tmp_path / "rank0.txt").read_text() == "ok 0 501" and (tmp_path / "rank1.txt").read_text() == "ok 501 1001"


def test_scene_file_roundtrip_through_reference_driver(tmp_path):
    """scenes.Scene.write -> oracle/scene_io.h reader -> reference bake: the basic scenario must give
    the reference's lightmap sizes 8, 1, 16 and three lightmaps with uids 1..3."""
    if not parity.have_reference():
        pytest.skip("oracle/_ref not built")
    out = parity.run_reference(scenes.scene_basic(), threads=1)
    assert [(lm["uid"], lm["width"], lm["height"]) for lm in out["lightmaps"]] == [(1, 8, 8), (2, 1, 1), (3, 16, 16)]
    assert [i["n"] for i in out["instances"]] == [0, 64, 1, 196]


def test_mesh_import_counts():
    assert [len(scenes.load_mesh(n).idx) // 3 for n in ("test-mesh", "test-set2-mesh1", "test-set2-mesh2")] == [214, 246, 2048]
    sc = scenes.scene_mesh2()
    assert sc.triangle_count() == 246 + 2048


@pytest.mark.parametrize("threads,fork_min", [(1, 0), (4, 8), (3, 300)])
def test_reference_order_tree_equals_reference_dump(refprims, monkeypatch, threads, fork_min):
    """threads > 1: the halves of big nodes are built by different threads and spliced (reftree.cpp); the bytes must still be
    the reference's.  Inputs include a 200 x 150 grid whose triangle centres tie on every axis (the order among ties is
    whatever std::sort leaves -- the reference's and ours must leave the same)."""
    import ctypes as C
    if threads > 1:
        monkeypatch.setenv("LTR_REFTREE_THREADS", str(threads))
        monkeypatch.setenv("LTR_REFTREE_FORK_MIN", str(fork_min))
    rng = np.random.default_rng(3)
    for n in (0, 1, 4, 5, 33, 700, 6000, 60000):
        tris = (rng.uniform(-5, 5, (n, 1, 3)) + rng.uniform(-0.6, 0.6, (n, 3, 3))).astype(np.float32).reshape(n, 9)
        if n >= 33:                                   # big triangles that must stay in inner nodes
            tris[::11] = (tris[::11].reshape(-1, 3, 3) * np.float32(6)).reshape(-1, 9)
        if n == 60000:                                # regular grid: many equal keys
            gx, gy = np.meshgrid(np.arange(200, dtype=np.float32), np.arange(150, dtype=np.float32), indexing="ij")
            o = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size, np.float32)], 1) * np.float32(0.25)
            e = np.float32(0.25)
            a = np.concatenate([o, o + [e, 0, 0], o + [e, e, 0]], 1); b = np.concatenate([o, o + [e, e, 0], o + [0, e, 0]], 1)
            tris = np.concatenate([a, b]).astype(np.float32)
        nodes, items = api.test_reftree(tris)
        h = refprims.L.refp_tritree_create(tris.ctypes.data_as(C.POINTER(C.c_float)), n)
        assert refprims.L.refp_tritree_tri_count(h) == n
        rn = np.zeros((refprims.L.refp_tritree_node_count(h), 8), np.uint32)
        ri = np.zeros(max(refprims.L.refp_tritree_item_count(h), 1), np.int32)
        refprims.L.refp_tritree_dump(h, rn.ctypes.data, ri.ctypes.data)
        refprims.L.refp_tritree_destroy(h)
        assert np.array_equal(rn, nodes), n
        assert np.array_equal(ri[:len(items)], items), n


def test_flat_bvh_invariants():
    rng = np.random.default_rng(5)
    for n, leaf in ((0, 4), (1, 4), (2, 1), (100, 2), (5000, 4), (5000, 7)):
        tris = (rng.uniform(-5, 5, (n, 1, 3)) + rng.uniform(-0.3, 0.3, (n, 3, 3))).astype(np.float32).reshape(n, 9)
        b = api.test_bvh(tris, leaf)
        assert b["ok"], (n, leaf)
        assert sorted(b["order"].tolist()) == list(range(n))
        if n:
            assert np.allclose(b["bounds"][:3], tris.reshape(-1, 3).min(0)) and np.allclose(b["bounds"][3:], tris.reshape(-1, 3).max(0))
    # 4000 coincident triangles cannot be separated spatially: the builder must still terminate with bounded depth
    tris = np.tile(np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32), (4000, 1))
    b = api.test_bvh(tris, 4)
    assert b["ok"] and b["depth"] <= 20


def _pair_links(Pr, Nr, Pc, Nc):
    """The reference's pair criterion in its float32 operation order (lighter.cpp:735-750): dotA = Ni.d, dotB = Nj.(-d) with
    d = Pj - Pi, both > 0.001, factor dotA*dotB/(len^4*pi) >= 0.001.  Returns the (rows, cols) matrix of passing pairs."""
    f = np.float32
    d = (Pc[None, :, :] - Pr[:, None, :]).astype(f)
    dot = lambda a, b: ((a[..., 0] * b[..., 0]).astype(f) + (a[..., 1] * b[..., 1]).astype(f)).astype(f) + (a[..., 2] * b[..., 2]).astype(f)
    dA = dot(np.broadcast_to(Nr[:, None, :], d.shape), d).astype(f)
    dB = dot(np.broadcast_to(Nc[None, :, :], d.shape), (-d).astype(f)).astype(f)
    l2 = dot(d, d).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        fac = ((dA * dB).astype(f) / ((l2 * l2).astype(f) * f(3.14159274101257324)).astype(f)).astype(f)
    return (dA > f(0.001)) & (dB > f(0.001)) & ~(fac < f(0.001))


def test_radiosity_culling_never_rejects_a_linking_pair(oracle):
    """csrc/rad_cull.h (shared by the pair-sweep kernel and this host hook): the tile x tile interval test and the per-row
    group test may only skip blocks in which NO pair passes the reference's pair criterion, and the FMA pre-filter may only drop
    pairs that fail it.  Blocks are drawn around the
    thresholds: grazing normals, distances around the 17.84-unit cut-off, facing / averted / coplanar patches, large
    coordinates.  The numpy criterion used as truth is itself anchored to the oracle's candidate count."""
    rng = np.random.default_rng(23)

    def unit(v):
        return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)

    # anchor: candidates of a random point set, numpy criterion vs the oracle's own count (no triangles: nothing is blocked)
    P = rng.uniform(-3, 3, (160, 3)).astype(np.float32)
    N = unit(rng.normal(size=(160, 3)))
    m = _pair_links(P, N, P, N)
    n_np = int(np.triu(m, 1).sum())
    _, _, _, pairs, segs = oracle.rad_links(np.zeros((0, 9), np.float32), P, N)
    assert pairs == 160 * 159 // 2 and segs == n_np and n_np > 100

    linking_blocks = culled_blocks = n_fast = n_truth = 0
    for trial in range(1500):
        nr, nc = int(rng.choice([1, 8, 32])), int(rng.choice([4, 8, 128]))
        off = rng.choice([0.0, 50.0, 390.0]) * rng.uniform(-1, 1, 3)
        cr = off + rng.uniform(-1, 1, 3)
        kind = trial % 5
        dist = [rng.uniform(0.05, 2.0), rng.uniform(2.0, 12.0), rng.uniform(17.0, 18.5), rng.uniform(0.3, 6.0), rng.uniform(18.0, 30.0)][kind]
        dirv = unit(rng.normal(size=3))
        cc = cr + dirv * dist
        Pr = (cr + rng.normal(size=(nr, 3)) * rng.choice([0.02, 0.2, 0.8])).astype(np.float32)
        Pc = (cc + rng.normal(size=(nc, 3)) * rng.choice([0.02, 0.2, 0.8])).astype(np.float32)
        if kind == 3:                                # nearly coplanar patches with grazing normals (floor-to-floor pairs)
            n0 = unit(np.cross(dirv, rng.normal(size=3)))
            Nr = unit(n0 + rng.normal(size=(nr, 3)) * 0.02 + dirv * rng.uniform(-0.01, 0.02))
            Nc = unit(n0 + rng.normal(size=(nc, 3)) * 0.02 - dirv * rng.uniform(-0.01, 0.02))
        else:                                        # facing each other, with a random spread
            spread = rng.choice([0.0, 0.3, 1.5])
            Nr = unit(dirv + rng.normal(size=(nr, 3)) * spread)
            Nc = unit(-dirv + rng.normal(size=(nc, 3)) * spread)
        truth = _pair_links(Pr, Nr, Pc, Nc)
        block_ok, row_ok, fast = api.test_rad_cull(Pr, Nr, Pc, Nc)
        assert (fast | ~truth).all(), (trial, kind)            # the lock-step FMA pre-filter keeps every linking pair
        n_fast += int(fast.sum()); n_truth += int(truth.sum())
        if truth.any():
            linking_blocks += 1
            assert block_ok, (trial, kind)
        else:
            culled_blocks += not block_ok
        assert (row_ok | ~truth.any(1)).all(), (trial, kind)
    assert linking_blocks > 300 and culled_blocks > 100, (linking_blocks, culled_blocks)
    assert n_truth > 10000 and n_fast < 1.5 * n_truth, (n_fast, n_truth)      # ... and is not vacuous: few extra survivors


def _host_prepare_scenes():
    from lighter_b200 import scenes
    yield "basic", scenes.NAMED["basic"]()
    yield "mesh2", scenes.NAMED["mesh2"]()
    yield "rad1", scenes.NAMED["rad1"]()
    yield "hugeoverlap", scenes.NAMED["hugeoverlap"]()
    yield "config4_sibling", scenes.workload("config4_sibling")
    sc = scenes.workload("config4_sibling")
    for k, inst in enumerate(sc.instances):
        if k % 2:
            inst.shadow = 0                          # instances that do not cast shadows: the scene BVH gets a compacted copy
    yield "sibling_noshadow_inst", sc
    sc = scenes.NAMED["mesh2"]()
    for m in sc.meshes:
        for j, part in enumerate(m.parts):
            if j % 2 == 0:
                part.shadow = 0                      # parts that do not cast shadows never enter the triangle lists
    yield "mesh2_noshadow_parts", sc


@pytest.mark.parametrize("poison", [False, True])
def test_host_prepass_arrays_are_pinned(poison, monkeypatch):
    """The host pre-pass (world transform, triangle lists, reference-order trees, raster list, scene BVH, instance table) runs
    without a device; FNV-1a fingerprints of every array it would upload are pinned (tests/golden/host_prepare.json, generated
    by this same hook before the pre-pass was reorganised around one shared triangle array), so host-side optimisation
    cannot silently change what the GPU stages see.  poison=True fills every default-initialised big array with a byte pattern
    first (fresh pages are zero and would hide an element that nobody writes)."""
    if poison:
        monkeypatch.setenv("LTR_POISON_BIGALLOC", "1")
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "host_prepare.json")))
    names = ["wpos", "wnrm", "tex", "rtris", "rnodes", "ritems", "rtree_tris", "bvh.nodes", "bvh.nodes4", "bvh.order", "bvh_tris", "inst+light_inst"]
    for name, sc in _host_prepare_scenes():
        with api.BakeHandle(sc) as h:
            got = h.host_prepare_fingerprints()
        diff = [names[i] for i in range(12) if hex(got[i]) != gold[name][i]]
        assert not diff, (name, diff)


def test_flat_bvh_does_not_depend_on_thread_count_or_page_size(monkeypatch):
    """The builder's cooperative top, its dynamically scheduled sub-trees, the SSE2 passes and the huge-page work arrays must
    not change the tree: the triangle order (which encodes every partition) is the same for 1, 3 and 8 threads, with and
    without huge pages, on a scene large enough (> 65536 triangles) to take the cooperative path."""
    rng = np.random.default_rng(17)
    n = 150_000
    tris = (rng.uniform(-40, 40, (n, 1, 3)) * np.array([1, 1, 0.05]) + rng.uniform(-0.3, 0.3, (n, 3, 3))).astype(np.float32).reshape(n, 9)
    ref = None
    for threads, nohuge in (("1", False), ("3", False), ("8", False), ("8", True)):
        monkeypatch.setenv("LTR_BVH_THREADS", threads)
        if nohuge:
            monkeypatch.setenv("LTR_NO_HUGEPAGES", "1")
        b = api.test_bvh(tris, 2)
        assert b["ok"], threads
        key = (b["n_nodes"], b["depth"], b["order"].tobytes())
        if ref is None:
            ref = key
        assert key == ref, (threads, nohuge)


def test_bvh_entry_set_keeps_every_triangle_in_range_reachable():
    """csrc/bvh_entry.h: for bundles of segments (compact, scene-wide, degenerate, empty) the entry set must keep every
    triangle whose box overlaps the bundle box reachable, and the any-hit walk from the entry set must agree with the walk
    from the root on every segment -- while reading fewer nodes on compact bundles."""
    rng = np.random.default_rng(11)
    for n, leaf in ((0, 2), (1, 2), (3, 2), (400, 2), (20000, 2), (20000, 7)):
        tris = (rng.uniform(-20, 20, (n, 1, 3)) * np.array([1, 1, 0.1]) + rng.uniform(-0.4, 0.4, (n, 3, 3))).astype(np.float32).reshape(n, 9)
        segs, off = [], [0]
        for b in range(40):
            k = int(rng.integers(1, 200))
            if b % 4 == 3:                           # scene-wide bundle: the entry set has to stay near the root
                a, e = rng.uniform(-22, 22, (k, 3)), rng.uniform(-22, 22, (k, 3))
            else:                                    # compact bundle
                c = rng.uniform(-18, 18, 3) * np.array([1, 1, 0.1])
                a, e = c + rng.uniform(-1, 1, (k, 3)), c + rng.uniform(-3, 3, (k, 3))
            if b == 5:
                e = a.copy()                         # zero-length segments
            if b == 6:
                a[:, 0] = e[:, 0] = 1.5              # axis-parallel: zero direction components
            segs.append(np.c_[a, e]); off.append(off[-1] + k)
        off.append(off[-1])                          # an empty bundle
        r = api.test_bvh_entry(tris, np.concatenate(segs), np.array(off, np.uint32), leaf)
        if n == 0:
            continue
        assert r["ok"], (n, leaf)
        assert r["mismatches"] == 0, (n, leaf, r["mismatches"])
        assert r["entries"].max() <= 8 and r["entries"][-1] == 0
        if n >= 20000:
            assert r["visits_entry"] < 0.8 * r["visits_root"], r


def test_bvh_entry_set_v2_shaft_and_leaf_entries_change_no_ray():
    """csrc/bvh_entry.h version 2 (rad_visibility_kernel): bundles whose segments run from one box to another.  With the shaft
    planes, leaf entries and up to 16 entries, every segment must get the SAME answer as on a walk from the root and -- the
    sharper check -- test exactly the same number of triangles: a box dropped by the shaft test that a ray would have entered
    shows up as a difference even if it held no blocker."""
    rng = np.random.default_rng(12)
    for n, leaf in ((1, 2), (3, 2), (400, 2), (30000, 2), (30000, 7)):
        tris = (rng.uniform(-20, 20, (n, 1, 3)) * np.array([1, 1, 0.15]) + rng.uniform(-0.5, 0.5, (n, 3, 3))).astype(np.float32).reshape(n, 9)
        segs, off = [], [0]
        for b in range(60):
            k = int(rng.integers(1, 300))
            r0 = rng.uniform(-18, 18, 3) * np.array([1, 1, 0.15])
            c0 = r0 + rng.uniform(-1, 1, 3) * np.array([12, 12, 2]) * (b % 3 != 0)      # every third bundle: R and C overlap
            a = r0 + rng.uniform(-0.4, 0.4, (k, 3))
            e = c0 + rng.uniform(-1.5, 1.5, (k, 3))
            if b == 5:
                e = a.copy()                         # zero-length segments
            if b == 6:
                a[:, 0] = e[:, 0] = 1.5              # axis-parallel: zero direction components, flat boxes
            if b == 7:
                a[:] = a[0]                          # R is a point
            if b == 8:                               # grazing: segments in the plane z = const of flat triangles
                a[:, 2] = e[:, 2] = 0.0
            segs.append(np.c_[a, e]); off.append(off[-1] + k)
        off.append(off[-1])                          # an empty bundle
        segs = np.concatenate(segs)
        for me, shaft in ((8, True), (8, False), (5, True), (3, True)):
            r = api.test_bvh_entry2(tris, segs, np.array(off, np.uint32), leaf, me, shaft)
            assert r["ok"], (n, leaf, me, shaft)
            assert r["mismatches"] == 0 and r["test_diffs"] == 0, (n, leaf, me, shaft, r["mismatches"], r["test_diffs"])
            assert r["entries"].max() <= me and r["entries"][-1] == 0
        for batch, shaft in ((32, True), (32, False), (7, True), (1, True)):       # packet form: leaves listed per batch of segments
            r = api.test_bvh_entry2(tris, segs, np.array(off, np.uint32), leaf, 8, shaft, batch)
            assert r["ok"] and r["mismatches"] == 0 and r["test_diffs"] == 0, (n, leaf, batch, shaft, r["mismatches"], r["test_diffs"])
        if n >= 30000 and leaf == 2:
            r1 = api.test_bvh_entry2(tris, segs, np.array(off, np.uint32), leaf, 8, True)
            r0 = api.test_bvh_entry2(tris, segs, np.array(off, np.uint32), leaf, 8, False)
            assert r1["visits_entry"] < r0["visits_entry"] < r0["visits_root"], (r0, r1)


def test_rand_replay_equals_libc_stream_and_leaves_libc_in_step():
    """The AO pass consumes one libc rand() per lumel (lighter.cpp:819).  The library's fast replay must give
    exactly rand()/RAND_MAX, and the libc generator must afterwards stand where that many rand() calls leave
    it (a caller's own rand() use, and the next bake, continue the same stream)."""
    RAND_MAX = 2147483647
    for seed, n in ((1, 5), (1, 1000), (20261017, 100_003), (7, 0)):
        api.srand(seed)
        want = np.array([api.libc_rand() for _ in range(n)], np.float64)
        tail_want = [api.libc_rand() for _ in range(40)]
        api.srand(seed)
        got, fast = api.test_rand_fill(n)
        tail_got = [api.libc_rand() for _ in range(40)]
        assert fast, "glibc table-level path expected in this image"
        assert np.array_equal(got, (want.astype(np.float32) / np.float32(RAND_MAX)))
        assert tail_got == tail_want
