#!/usr/bin/env python
"""bench.py -- the bake-path benchmark (contract in the task statement; metric of BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one full bake (lumel generation -> direct-light march -> radiosity -> AO -> finalize)
of the workload.  Default workload: BASELINE.json configs[3], the configuration the headline metric
is quoted on -- synthetic 1M-triangle scene, 256 x 256^2 = 4096^2 lightmap texels, 8 local lights +
1 directional, AO (17 segments/lumel) and 3 radiosity bounces -- sharded over the N GPUs (strong
scaling: the scene is fixed, lumels are split).

  value   rays/s = (distance queries + AO/radiosity/correction segments actually traced, summed
          over ranks) / device time of K steps with the scene already resident in HBM
          (ltrx_Prepare once, then ltrx_BakeResident per step; CUDA events on the bake stream,
          max over ranks).
  e2e     the same metric through the reference-facing C API with HOST buffers: ltr_Start ->
          ltr_GetStatus()==0 polling loop (SURVEY 8d wall-time definition: host pre-pass, BVH build,
          H2D, all GPU stages and the D2H of the lightmaps are inside the timed region).
  --impl reference   times the UNMODIFIED reference (oracle/_ref/ref_bake, compiled from
          /root/reference) on the host's cores, all threads, on a bounded same-generator sibling of
          the workload; ray counts come from the reference's own counting build
          (tests/golden/ref_counts.json).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD_DESC = {
    "config4": "BASELINE configs[3]: synthetic 1M-tri open terrain+pillars, 256 instances x 256^2 (=4096^2) lightmap texels, "
               "8 point/spot lights + 1 directional (range 60), AO 17 samples, 3 radiosity bounces",
    "config4_quarter": "quarter-size sibling of configs[3]: 250k tris, 64 instances x 256^2, same lights/AO/bounces",
    "config3": "BASELINE configs[2]: synthetic 250k-tri closed interior, 64 instances x 256^2 (=2048^2) texels, 32 point/spot lights",
    "config4_sibling": "CPU-sized sibling of configs[3]: same generator, 4 instances x 64^2, 15.6k tris, 2 lights + 1 directional, AO, 3 bounces",
    "config3_sibling": "CPU-sized sibling of configs[2]: same generator, 4 instances x 64^2, 15.6k tris, 4 lights",
    "mesh1": "BASELINE configs[0]: bin/test-mesh.data, 3 lights, AO",
    "mesh2": "BASELINE configs[1]: test-set2 two-mesh scene, 4 lights, AO, normal map",
}
CPU_SIBLING = {"config4": "config4_sibling", "config4_quarter": "config4_sibling", "config3": "config3_sibling"}


def rays_of(st: dict) -> int:
    return int(st["n_distance_queries"] + st["n_ao_segments"] + st["n_rad_segments"] + st["n_correction_rays"])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.idx, self.samples, self.stop_flag = gpu_index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.samples.append([x.strip() for x in r.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self) -> dict:
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def reference_arm(args) -> dict:
    """Time the unmodified reference on the host cores (rank 0 only)."""
    from lighter_b200 import parity, scenes
    name = CPU_SIBLING.get(args.workload, args.workload)
    counts = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_counts.json")))
    if name not in counts:
        raise SystemExit(f"no committed reference ray count for workload {name}")
    sc = scenes.workload(name)
    walls = []
    for it in range(args.warmup + args.steps):
        out = parity.run_reference(sc, threads=0, internals=False)
        if it >= args.warmup:
            walls.append(out["wall_s"])
    cores = out["threads"]
    rays = counts[name]["rays"]
    val = rays * len(walls) / sum(walls)
    sample = f"{name}: {WORKLOAD_DESC.get(name, name)}; {rays} rays per bake (reference's own counting build), full bake per step"
    return {
        "metric": "rays_per_s", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "reference_sample": sample},
        "bake_wall_s": sum(walls) / len(walls),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("LTR_BENCH_WORKLOAD", "config4"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args)), flush=True)
        return

    import numpy as np
    import torch
    from lighter_b200 import api, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- lighter_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (api.C.c_char * 128)()
            if not api.lib().ltrx_NcclUniqueId(buf):
                raise SystemExit("ltrx_NcclUniqueId failed")
            idt = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)                         # torch.distributed is plumbing: rendezvous + barriers + max-reduce
        nccl_id = bytes(idt.cpu().tolist())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def reduce_max(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sc = scenes.workload(args.workload)
    shard = (rank, world, nccl_id) if world > 1 else None

    # ---- device-resident arm: scene + BVH uploaded once, K timed bakes ----------------------------
    h = api.BakeHandle(sc, device=local_rank, shard=shard)
    h.prepare()
    for _ in range(args.warmup):
        h.bake_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    step_ms, stats = [], None
    for _ in range(args.steps):
        ms = h.bake_resident()                         # CUDA events on the bake stream, first stage -> last stage
        step_ms.append(reduce_max(ms))
        stats = h.stats()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=3)
    rays_step = reduce_sum(rays_of(stats))
    launches_step = stats["kernel_launches"]
    total_ms = sum(step_ms)
    value = rays_step * args.steps / (total_ms * 1e-3)

    # roofline of the dominant kernel (DESIGN.md "Roofline"): traversal kernels are charged the bytes
    # their BVH walk touches: 64 B per node visit + 160 B per point/triangle test.
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    stage_ms = {"march": stats["gpu_ms_march"], "radiosity_pairs": stats["gpu_ms_rad_pairs"], "radiosity_visibility": stats["gpu_ms_rad_vis"],
                "ao": stats["gpu_ms_ao"], "lumels": stats["gpu_ms_samples"], "finalize": stats["gpu_ms_finalize"]}
    # Per-launch DRAM traffic measured once with `ncu --set full` (profiles/r01_traffic.json; None when not captured)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {})

    def roof(kernel, nbytes, ms, launches, note, extra=None):
        """achieved = algorithmic bytes per launch / average launch duration (CUDA events on the bake stream)."""
        if not ms or ms <= 0:
            return None
        launches = max(int(launches or 1), 1)
        ach = nbytes / (ms * 1e-3) / 1e9
        r = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
             "traffic": traffic.get(kernel), "peak_source": peak_src, "launches_per_step": launches,
             "algorithmic_bytes_per_launch": nbytes / launches, "kernel_ms_per_launch": ms / launches,
             "algorithmic_bytes_per_step": nbytes, "kernel_ms_per_step": ms, "note": note}
        if extra:
            r.update(extra)
        return r

    # SURVEY 8d unit costs: 64 B per BVH node visit, 160 B per point/triangle test (PreparedTri), 64 B per
    # segment/triangle test (RayTri), 36 B per march (lumel in, factor out), 32 B per segment.
    batches = max(int(stats.get("n_rad_batches", 1)), 1)
    rad_share = stats["n_rad_segments"] / max(stats["n_rad_segments"] + stats["n_ao_segments"], 1)
    roofs = {
        "rad_visibility_kernel": roof(
            "rad_visibility_kernel",
            (stats["n_ray_node_visits"] * 64.0 + stats["n_ray_tri_tests"] * 64.0 + stats.get("n_ray_entry_tests", 0) * 28.0) * rad_share
            + stats["n_rad_segments"] * (12.0 + 32.0),
            stats["gpu_ms_rad_vis"], batches,
            "any-hit segment traversal of the radiosity candidates, started at the entry set of each 1024-candidate chunk: 64 B/node visit + "
            "64 B/triangle test + 28 B/entry box (the counters are shared with the AO pass and apportioned by segment count) + 12 B candidate + "
            "32 B endpoints per segment; the bytes are served by L1/L2 (the scene is cache resident, ncu: DRAM traffic is a few % of them), so "
            "this is CACHE bandwidth and the fraction of the HBM peak can exceed 1; ncu: issue slots 81 % busy, L1 data-pipe wavefronts 76 % of "
            "peak, 24 of 32 lanes active -- the kernel is co-limited by instruction issue and L1 wavefronts, not by HBM",
            {"segments_per_s": stats["n_rad_segments"] / (stats["gpu_ms_rad_vis"] * 1e-3) if stats["gpu_ms_rad_vis"] else None}),
        "rad_candidates_kernel": roof(
            "rad_candidates_kernel",
            stats["n_rad_tile_loads"] * 5120.0 + stats["n_rad_segments"] * 12.0,
            stats["gpu_ms_rad_pairs"], batches,
            "pair sweep: 5 KiB per column tile staged by TMA bulk copy (positions, normals, group bounds) + 12 B per candidate written; "
            "FP32-issue bound (22 instructions per lumel pair per lane), the HBM fraction is small by construction",
            {"pair_tests_per_s": stats["n_rad_pairs"] / (stats["gpu_ms_rad_pairs"] * 1e-3) if stats["gpu_ms_rad_pairs"] else None}),
        "direct_march_kernel": roof(
            "direct_march_kernel",
            stats["n_node_visits"] * 64.0 + stats["n_tri_tests"] * 160.0 + stats["n_marches"] * 36.0,
            stats["gpu_ms_march"], 1,
            "distance-query traversal: 64 B/node + 160 B/point-triangle test + 36 B/march; L1/L2 served"),
    }
    roofs = {k: v for k, v in roofs.items() if v}
    dominant = max(roofs, key=lambda k: roofs[k]["kernel_ms_per_step"]) if roofs else None
    roofline = roofs.get(dominant)

    # ---- end-to-end arm: public C API, host buffers in, host lightmaps out ---------------------------
    e2e_walls, e2e_stats = [], None
    for it in range(args.e2e_steps + 1):
        hh = api.BakeHandle(sc, device=local_rank, shard=shard)       # scene set-up (ltr_MeshAddPart...) is outside the timed region
        barrier()
        w = hh.run()
        wmax = reduce_max(w)
        if it > 0:                                                     # first one warms allocator / NCCL communicator
            e2e_walls.append(wmax)
            e2e_stats = hh.stats()
        hh.close()
    e2e_rays = reduce_sum(rays_of(e2e_stats))
    e2e_val = e2e_rays * len(e2e_walls) / sum(e2e_walls)

    line = {
        "metric": "rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "name": args.workload, "triangles": sc.triangle_count(),
                   "lumels": int(stats["n_lumels_total"]), "lights": len(sc.lights), "parallelism": f"lumel-shard x{world}",
                   "l2_policy": "inputs (>=126 MB of lumel/link arrays per step) exceed L2; every step regenerates all buffers"},
        "bake_wall_s": sum(e2e_walls) / len(e2e_walls),
        "bake_device_s": total_ms / args.steps * 1e-3,
        "rays_per_step": rays_step,
        "stage_ms": stage_ms,
        "counters": {k: int(stats[k]) for k in ("n_marches", "n_distance_queries", "n_ao_segments", "n_rad_pairs", "n_rad_segments", "n_rad_links",
                                                 "n_node_visits", "n_tri_tests", "n_ray_node_visits", "n_ray_tri_tests", "n_ray_entry_tests", "n_rad_tile_loads", "n_rad_batches")},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": int(e2e_stats["h2d_bytes"]), "d2h_bytes_per_step": int(e2e_stats["d2h_bytes"]),
                "bake_wall_s": sum(e2e_walls) / len(e2e_walls), "steps": len(e2e_walls),
                "host_s": {k: e2e_stats[k] for k in ("t_prexform", "t_accel", "t_upload", "t_samples", "t_direct", "t_radiosity", "t_ao", "t_finalize", "t_readback")}},
        "gpu_launches": int(launches_step * args.steps),
        "roofline": roofline,
        "roofline_other": {k: v for k, v in roofs.items() if k != dominant},
    }
    h.close()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ra = argparse.Namespace(**vars(args))
            ra.steps, ra.warmup = 2, 0
            line["cpu_baseline"] = reference_arm(ra)["cpu_baseline"]
        except Exception as e:                              # noqa: BLE001
            line["cpu_baseline"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
