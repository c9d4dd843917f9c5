"""Offline (CPU, no GPU) estimate of what the bundle entry sets (csrc/bvh_entry.h) save on a workload: a synthetic lumel
cloud is laid over a window of the scene's triangles, Morton-sorted and cut into 128-lumel tiles as the radiosity sweep
does; for a sample of row warps the linking candidates are generated in sweep order, cut into 1024-candidate chunks, and
ltrx_test_bvh_entry walks every segment from the root and from its chunk's entry set on the REAL scene BVH.

    python tools/entry_estimate.py [config4] [window] [row_warps] [chunk]

What this model got right and wrong (round 2, profiles/r02_ab_runs.md): its CHUNK-level predictions held on the real bake
(entry sets, version-2 shaft planes: node reads per ray within 10 % of the GPU's counters).  Its BATCH-level prediction did not:
it promised 7 leaves in the shaft of a batch of 32 consecutive candidates, the GPU's counters say 35 (the packet form of the
visibility kernel was built on this number and retired).  Known differences from the real bake: a 60-unit window of the scene,
lumels scattered at random instead of on texel grids, candidates generated row-major instead of in the sweep's (column group,
row, column) order.
"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lighter_b200 import api, scenes


def world_tris(sc):
    out = []
    for inst in sc.instances:
        for p in sc.meshes[inst.mesh].parts:
            v = np.c_[p.pos, np.ones(len(p.pos), np.float32)] @ inst.matrix
            out.append(v[:, :3][p.idx.reshape(-1, 3)].astype(np.float32))
    return np.concatenate(out).reshape(-1, 9)


def morton(p, lo, ext):
    q = np.clip(((p - lo) / ext * 1023).astype(np.int64), 0, 1023)
    def spread(x):
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "config4"
    window = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    n_rw = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
    sc = scenes.workload(name)
    tris = world_tris(sc)
    print("triangles", len(tris))
    t3 = tris.reshape(-1, 3, 3)
    ctr = t3.mean(1)
    mid = (t3.reshape(-1, 3).min(0) + t3.reshape(-1, 3).max(0)) / 2
    sel = (np.abs(ctr[:, 0] - mid[0]) < window / 2) & (np.abs(ctr[:, 1] - mid[1]) < window / 2)
    w = t3[sel]
    n = np.cross(w[:, 1] - w[:, 0], w[:, 2] - w[:, 0])
    area = np.linalg.norm(n, axis=1) / 2
    n = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    rng = np.random.default_rng(1)
    density = 105.0                                   # lumels per unit^2: 256^2 texels over a 25 x 25 tile
    cnt = rng.poisson(area * density)
    tid = np.repeat(np.arange(len(w)), cnt)
    u, v = rng.random(len(tid)), rng.random(len(tid))
    f = u + v > 1
    u[f], v[f] = 1 - u[f], 1 - v[f]
    P = (w[tid, 0] + (w[tid, 1] - w[tid, 0]) * u[:, None] + (w[tid, 2] - w[tid, 0]) * v[:, None]).astype(np.float32)
    N = n[tid].astype(np.float32)
    print("lumels in window", len(P))
    lo, hi = P.min(0), P.max(0)
    order = np.argsort(morton(P, lo, (hi - lo).max()), kind="stable")
    P, N = P[order], N[order]
    segs, offs = [], [0]
    margin = 18.0
    inner = np.where((np.abs(P[:, 0] - mid[0]) < window / 2 - margin) & (np.abs(P[:, 1] - mid[1]) < window / 2 - margin))[0] // 32
    rws = rng.choice(np.unique(inner), size=min(n_rw, len(np.unique(inner))), replace=False)
    total_c = 0
    for rw in rws:
        rows = np.arange(rw * 32, min(rw * 32 + 32, len(P)))
        cand = []
        near = np.where(np.linalg.norm(P - P[rows].mean(0), axis=1) < 19.0)[0]
        near = near[near > rows[-1]]                  # a pair is swept once: partner later on the curve
        for t0 in np.unique(near // 128):             # column tiles in Morton order
            cols = near[near // 128 == t0]
            d = P[cols][None, :, :] - P[rows][:, None, :]
            da = np.einsum("rk,rck->rc", N[rows], d)
            db = -np.einsum("ck,rck->rc", N[cols], d)
            l2 = (d * d).sum(-1)
            ok = (da > 1e-3) & (db > 1e-3) & (da * db / (l2 * l2 * np.pi + 1e-30) >= 1e-3)
            r, c = np.nonzero(ok)
            if len(r):
                A, B = P[rows][r], P[cols][c]
                dn = (B - A) / np.linalg.norm(B - A, axis=1, keepdims=True)
                cand.append(np.c_[A + dn * 1e-3, B - dn * 1e-3])
        if not cand:
            continue
        cand = np.concatenate(cand).astype(np.float32)
        total_c += len(cand)
        for k in range(0, len(cand), chunk):
            segs.append(cand[k:k + chunk])
            offs.append(offs[-1] + len(segs[-1]))
    segs = np.concatenate(segs)
    print("row warps", len(rws), "candidates", total_c, "chunks", len(offs) - 1)
    t = time.time()
    r = api.test_bvh_entry(tris, segs, np.array(offs, np.uint32), leaf_max=2)
    print("ok", r["ok"], "mismatches", r["mismatches"], "time %.1fs" % (time.time() - t))
    ns = len(segs)
    print("4-wide node reads per segment: root %.2f  entry %.2f (+ %.2f entry boxes)" % (r["visits_root"] / ns, r["visits_entry"] / ns, r["entry_tests"] / ns))
    print("entries per chunk: mean %.2f  hist %s" % (r["entries"].mean(), np.bincount(r["entries"], minlength=9)))
    # SIMT waiting: a warp traces 32 consecutive segments of a chunk in lock step, so a batch costs what its slowest lane costs.
    # Cost model per segment (instruction slots, from the ncu hot-line shares): 110 per node read, 70 per triangle test.
    nodes, tests = api.test_bvh_entry_cost(tris, segs, np.array(offs, np.uint32), leaf_max=2)
    cost = 110.0 * nodes + 70.0 * tests
    lock = refill = useful = 0.0
    for b in range(len(offs) - 1):
        c = cost[offs[b]:offs[b + 1]]
        useful += c.sum()
        pad = (-len(c)) % 32
        cb = np.concatenate([c, np.zeros(pad)]).reshape(-1, 32)
        lock += cb.max(1).sum() * 32                      # every batch waits for its slowest lane
        lanes = np.zeros(32)                              # lane refill: each lane takes the chunk's next segment when it is done
        for x in c:
            lanes[lanes.argmin()] += x
        refill += lanes.max() * 32
    print("walk cost per segment (model): %.0f slots; lock-step batches use %.0f %% of their lane slots, lane refill within a chunk %.0f %%"
          % (useful / ns, 100 * useful / lock, 100 * useful / refill))
    print("=> upper bound of the gain from lane refill on the traversal part: %.0f %%" % (100 * (1 - refill / lock)))
    # upper bound of ANY reordering of a chunk's segments before they are cut into lock-step batches: sorted by their true cost
    tot = 0.0
    for b in range(len(offs) - 1):
        c = np.sort(cost[offs[b]:offs[b + 1]])
        pad = (-len(c)) % 32
        tot += np.concatenate([c, np.zeros(pad)]).reshape(-1, 32).max(1).sum() * 32
    print("   lock-step batches of a chunk sorted by TRUE cost (bound of any reordering): %.0f %% of the lane slots used" % (100 * useful / tot))
    # by segment length (a key a kernel could compute)
    tot = 0.0
    seglen = np.linalg.norm(segs[:, 3:6] - segs[:, 0:3], axis=1)
    for b in range(len(offs) - 1):
        o = np.argsort(seglen[offs[b]:offs[b + 1]], kind="stable")
        c = cost[offs[b]:offs[b + 1]][o]
        pad = (-len(c)) % 32
        tot += np.concatenate([c, np.zeros(pad)]).reshape(-1, 32).max(1).sum() * 32
    print("   lock-step batches of a chunk sorted by segment LENGTH: %.0f %% of the lane slots used" % (100 * useful / tot))
    # refill inside groups of G segments only (a per-warp ring of G prepared rays, drained before the next G are prepared)
    for G in (64, 128, 256, 512):
        tot = 0.0
        for b in range(len(offs) - 1):
            c = cost[offs[b]:offs[b + 1]]
            for g0 in range(0, len(c), G):
                lanes = np.zeros(32)
                for x in c[g0:g0 + G]:
                    lanes[lanes.argmin()] += x
                tot += lanes.max() * 32
        print("   refill within groups of %4d segments: %.0f %% of the lane slots used" % (G, 100 * useful / tot))


if __name__ == "__main__":
    main()
