#!/usr/bin/env python
"""bench.py -- the bake-path benchmark (contract in the task statement; metric of BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one full bake (lumel generation -> direct-light march -> radiosity -> AO -> finalize)
of the workload.  Default workload: BASELINE.json configs[3], the configuration the headline metric
is quoted on -- synthetic 1M-triangle scene, 256 x 256^2 = 4096^2 lightmap texels, 8 local lights +
1 directional, AO (17 segments/lumel) and 3 radiosity bounces -- sharded over the N GPUs (strong
scaling: the scene is fixed, lumels are split).

  value     rays/s = (distance queries + AO/radiosity/correction segments actually traced, summed
            over ranks) / device time of K steps with the scene already resident in HBM
            (ltrx_Prepare once, then ltrx_BakeResident per step; CUDA events on the bake stream,
            max over ranks).  The units are ALSO reported separately (`rates`): marches/s, distance
            queries/s, any-hit segments/s, AO segments/s, lumel pairs/s (never mixed).
  e2e       the same metric through the reference-facing C API with HOST buffers: ltr_Start ->
            ltr_GetStatus()==0 polling loop (SURVEY 8d wall-time definition: host pre-pass, BVH build,
            H2D, all GPU stages and the D2H of the lightmaps are inside the timed region).
  parity    FNV-1a-64 of every lightmap of the last end-to-end bake (rank 0), checked against the hash
            committed for the workload in tests/golden/bench_hashes.json -- recorded from a single-GPU
            bake, so an N-GPU line proves sharded == solo.  A mismatch fails the run.
  roofline  dominant kernel: `achieved` = DRAM bytes it really moves (ncu dram__bytes per launch,
            profiles/r02_traffic.json) / its CUDA-event time, against the measured HBM peak; the BVH
            bytes its walk touches are served by L1/L2 and reported as `cache` (GB/s), with the limiter
            ncu names.  Traversal is instruction-issue / L1 bound, not HBM bound.
  cpu_baseline / same_config (N=1, rank 0)
            the UNMODIFIED reference (oracle/_ref/ref_bake, compiled from /root/reference) on the
            host's cores: per-stage per-unit rates on a CPU-sized same-generator sibling, a labelled
            EXTRAPOLATION of its wall time on the full workload, and wall / wall on workloads both
            sides bake completely (the sibling, BASELINE configs[1]).
  --impl reference   prints the reference arm alone: the sibling named in config.workload, rays from
            the reference's own counting build (tests/golden/ref_counts.json).
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
    "config5": "BASELINE configs[4] (sweep member): one merged 500k-tri instance, 1024^2 texels, 64 point/spot lights of range 14",
    "config4_sibling": "CPU-sized sibling of configs[3]: same generator, 4 instances x 64^2, 15.6k tris, 2 lights + 1 directional, AO, 3 bounces",
    "config3_sibling": "CPU-sized sibling of configs[2]: same generator, 4 instances x 64^2, 15.6k tris, 4 lights",
    "mesh1": "BASELINE configs[0]: bin/test-mesh.data, 3 lights, AO",
    "mesh2": "BASELINE configs[1]: test-set2 two-mesh scene, 4 lights, AO, normal map",
}
CPU_SIBLING = {"config4": "config4_sibling", "config4_quarter": "config4_sibling", "config3": "config3_sibling", "config5": "config3_sibling"}
SAME_CONFIG = {"config4": ["config4_sibling", "mesh2"], "config3": ["config3_sibling", "mesh2"], "config5": ["config3_sibling"], "mesh1": ["mesh1"], "mesh2": ["mesh2"]}


def rays_of(st: dict) -> int:
    return int(st["n_distance_queries"] + st["n_ao_segments"] + st["n_rad_segments"] + st["n_correction_rays"])


def load_json(*parts):
    p = os.path.join(ROOT, *parts)
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler(threading.Thread):
    """SM clock and clock-event (throttle) reasons of THIS rank's GPU during the timed region (B200_PROFILING.md recipe).
    Read through NVML in-process (pynvml): a query costs microseconds and takes no driver-wide lock.  Spawning `nvidia-smi`
    per sample -- what this class did before -- takes seconds on an 8-GPU box, overlapped the end-to-end arm and slowed every
    CUDA call of all eight ranks (r02: 0.235 s per end-to-end bake with eight nvidia-smi loops running, 0.112 s without);
    it is kept only as the fallback when pynvml is missing."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.idx, self.samples, self.stop_flag, self.source = gpu_index, [], False, "nvml"
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception:                                   # noqa: BLE001
            self.nv, self.source = None, "nvidia-smi"

    def sample(self):
        if self.nv is not None:
            mhz = int(self.nv.nvmlDeviceGetClockInfo(self.dev, self.nv.NVML_CLOCK_SM))
            mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
            self.samples.append((mhz, mask))
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        r = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=10)
        if r.returncode == 0 and r.stdout.strip():
            f = [x.strip() for x in r.stdout.strip().split(",")]
            mask = 0
            for (name, bit), v in zip(self.REASONS, f[2:6]):
                if v.lower().startswith("active"):
                    mask |= bit
            if f[0].isdigit():
                self.samples.append((int(f[0]), mask))
            if f[1].isdigit():
                self.max_mhz = int(f[1])

    def run(self):
        while not self.stop_flag:
            try:
                self.sample()
            except Exception:                               # noqa: BLE001
                pass
            time.sleep(0.05 if self.nv is not None else 0.2)        # ~20 samples per second of timed region (every NVML query takes a driver lock; the bake no longer calls into the driver's memory manager while it is timed)

    def stop(self):
        self.stop_flag = True
        self.join()                                         # nothing of the sampler may run on into the end-to-end arm

    def summary(self) -> dict:
        sm = sorted(s[0] for s in self.samples)
        mask = 0
        for s in self.samples:
            mask |= s[1]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for name, bit in self.REASONS if mask & bit), "samples": len(self.samples), "source": self.source}


# ---------------------------------------------------------------------------------------------------------
# the reference on the host cores
# ---------------------------------------------------------------------------------------------------------
def reference_walls(name: str, steps: int, warmup: int):
    from lighter_b200 import parity, scenes
    sc = scenes.NAMED[name]() if name in scenes.NAMED else scenes.workload(name)
    walls, out = [], None
    for it in range(warmup + steps):
        out = parity.run_reference(sc, threads=0, internals=False)
        if it >= warmup:
            walls.append(out["wall_s"])
    return walls, out["threads"], sc


def reference_arm(args) -> dict:
    """Time the unmodified reference on the host cores (rank 0 only) on the CPU-sized sibling of the workload."""
    name = CPU_SIBLING.get(args.workload, args.workload)
    counts = load_json("tests", "golden", "ref_counts.json")
    if name not in counts:
        raise SystemExit(f"no committed reference ray count for workload {name}")
    walls, cores, _ = reference_walls(name, args.steps, args.warmup)
    rays = counts[name]["rays"]
    val = rays * len(walls) / sum(walls)
    sample = f"{name}: {WORKLOAD_DESC.get(name, name)}; {rays} rays per bake (reference's own counting build), full bake per step"
    workload = WORKLOAD_DESC.get(name, name)
    if name != args.workload:
        workload = (f"{workload} -- a BOUNDED SAMPLE standing in for {args.workload} ({WORKLOAD_DESC.get(args.workload, args.workload)}), which the "
                    "reference cannot finish (serial O(N^2) radiosity, O(n^2) BSP build): this line's rays/s is NOT a same-work figure for the full "
                    "workload; same-scene wall ratios are in the b200 line's `same_config`")
    return {
        "metric": "rays_per_s", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload, "name": name, "sample_of": args.workload},
        "bake_wall_s": sum(walls) / len(walls),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def cpu_unit_rates(sibling: str, full_stats: dict, full_scene) -> dict:
    """SURVEY 8d 'CPU baseline timing': per-stage per-unit rates of the reference on the sibling (one timed bake, stage
    transitions from its own status polling loop), and the wall time they imply for the full workload -- an EXTRAPOLATION."""
    from lighter_b200 import parity, scenes
    counts = load_json("tests", "golden", "ref_counts.json")[sibling]
    sc = scenes.workload(sibling)
    r = parity.run_reference_timed(sc, threads=0)
    st, cores = r["stage_s"], r["threads"]
    n_inst = len(sc.instances)
    # the counting build gives exact unit counts of the sibling (tests/golden/ref_counts.json)
    t_direct, t_rad, t_ao = st.get("rendering lightmaps", 0.0), st.get("calculating radiosity", 0.0) + st.get("bouncing light", 0.0), st.get("rendering ambient occlusion", 0.0)
    t_samples, t_struct = st.get("generating samples", 0.0), st.get("generating data structures", 0.0)
    n_lumels_sib = counts["correction_rays"]                      # one correction ray per lumel on these scenes (>= 1 in general)
    pairs_sib = n_lumels_sib * (n_lumels_sib - 1) // 2              # the reference tests EVERY pair i < j (lighter.cpp:728-759)
    rates = {
        "cores": cores, "sample": sibling, "stage_s": {k: round(v, 4) for k, v in st.items()},
        "us_per_march": 1e6 * t_direct / max(counts["marches"], 1),
        "us_per_distance_query": 1e6 * t_direct / max(counts["distance_queries"], 1),
        "lumel_pairs_per_s": pairs_sib / t_rad if t_rad > 0 else None,
        "us_per_visibility_segment": 1e6 * t_rad / max(counts["visibility_segments"], 1) if counts["visibility_segments"] else None,
        "us_per_ao_segment": 1e6 * t_ao / max(counts["ao_segments"], 1) if counts["ao_segments"] else None,
        "lumels_per_s": n_lumels_sib / t_samples if t_samples > 0 else None,
        "data_structures_s_per_instance": t_struct * min(cores, n_inst) / n_inst,
        "note": "wall-clock per unit on all host threads; direct light and AO are parallel-fors over lumels, radiosity link generation is serial",
    }
    n = int(full_stats["n_lumels_total"])
    ni_full = len(full_scene.instances)
    parts = {
        "data structures (BSP + trees, per instance, one instance per thread)": rates["data_structures_s_per_instance"] * ni_full / max(min(cores, ni_full), 1),
        "generating samples": n / rates["lumels_per_s"] if rates["lumels_per_s"] else 0.0,
        "direct light (>= the GPU's distance-query count: the reference also marches back-facing directional pairs)":
            full_stats["n_distance_queries"] * rates["us_per_distance_query"] * 1e-6,
        "radiosity link generation, serial, ALL pairs i<j": (n * (n - 1) / 2) / rates["lumel_pairs_per_s"] if (rates["lumel_pairs_per_s"] and full_stats["n_rad_segments"]) else 0.0,
        "ambient occlusion": full_stats["n_ao_segments"] * (rates["us_per_ao_segment"] or 0.0) * 1e-6,
    }
    rates["extrapolated_full_workload"] = {
        "label": "EXTRAPOLATED from the sibling's per-unit rates, not measured: the reference cannot finish this workload",
        "parts_s": {k: float(f"{v:.4g}") for k, v in parts.items()}, "cpu_extrapolated_wall_s": float(f"{sum(parts.values()):.4g}"),
    }
    return rates


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("LTR_BENCH_WORKLOAD", "config4"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--record-hash", action="store_true", help="print the hash without checking it (to (re)generate tests/golden/bench_hashes.json)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args)), flush=True)
        return

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

    def reduce(x: float, op) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    reduce_max = lambda x: reduce(x, dist.ReduceOp.MAX if dist else None)
    reduce_sum = lambda x: reduce(x, dist.ReduceOp.SUM if dist else None)

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
    sampler.stop()
    keys = ("n_marches", "n_distance_queries", "n_ao_segments", "n_rad_pairs", "n_rad_segments", "n_rad_links", "n_node_visits", "n_tri_tests",
            "n_ray_node_visits", "n_ray_tri_tests", "n_ray_entry_tests", "n_rad_tile_loads", "n_correction_rays")
    job = {k: reduce_sum(float(stats[k])) for k in keys}          # whole-job unit counts (sum over ranks)
    rays_step = job["n_distance_queries"] + job["n_ao_segments"] + job["n_rad_segments"] + job["n_correction_rays"]
    launches_step = stats["kernel_launches"]
    total_ms = sum(step_ms)
    value = rays_step * args.steps / (total_ms * 1e-3)
    stage_ms = {"march": stats["gpu_ms_march"], "radiosity_pairs": stats["gpu_ms_rad_pairs"], "radiosity_visibility": stats["gpu_ms_rad_vis"],
                "ao": stats["gpu_ms_ao"], "lumels": stats["gpu_ms_samples"], "finalize": stats["gpu_ms_finalize"]}
    stage_ms = {k: reduce_max(v) for k, v in stage_ms.items()}    # slowest rank per stage

    def per_s(count, ms):
        return count / (ms * 1e-3) if ms and ms > 0 else None
    rates = {   # each unit against the kernel(s) that trace it, never mixed
        "marches_per_s": per_s(job["n_marches"], stage_ms["march"]),
        "distance_queries_per_s": per_s(job["n_distance_queries"], stage_ms["march"]),
        "visibility_segments_per_s": per_s(job["n_rad_segments"], stage_ms["radiosity_visibility"]),
        "lumel_pairs_per_s": per_s(job["n_rad_pairs"], stage_ms["radiosity_pairs"]),
        "ao_segments_per_s": per_s(job["n_ao_segments"], stage_ms["ao"]),
        "lumels_per_s": per_s(float(stats["n_lumels_total"]), stage_ms["lumels"]),
        "note": "whole-job unit count / slowest rank's CUDA-event time of the kernels that do that unit",
    }

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peaks = load_json("MEASURED_PEAKS.json")
    peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)") if peaks else (6650.0, "fallback (B200_PROFILING.md)")
    prof = load_json("profiles", "r02_traffic.json") or load_json("profiles", "r01_traffic.json")
    traffic = prof.get(args.workload, {})
    ncu = prof.get("ncu", {}).get(args.workload, {})
    batches = max(int(stats.get("n_rad_batches", 1)), 1)
    rad_share = stats["n_rad_segments"] / max(stats["n_rad_segments"] + stats["n_ao_segments"], 1)

    def roof(kernel, ms, launches, compulsory_bytes, cache_bytes, limiter, note, extra=None):
        """achieved = DRAM bytes the kernel moves per launch (ncu, one --set full capture scaled by duration) / its average
        CUDA-event launch time; cache = the BVH / triangle bytes its walk touches (L1/L2-served) over the same time."""
        if not ms or ms <= 0:
            return None
        launches = max(int(launches or 1), 1)
        t = traffic.get(kernel)
        ach = (t / (ms / launches * 1e-3) / 1e9) if t else (compulsory_bytes / (ms * 1e-3) / 1e9)
        r = {"bound": limiter, "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
             "traffic": t, "achieved_source": "ncu dram__bytes_read+write per average launch" if t else "compulsory algorithmic bytes (no ncu capture for this workload)",
             "peak_source": peak_src, "launches_per_step": launches, "kernel_ms_per_launch": ms / launches, "kernel_ms_per_step": ms,
             "algorithmic_hbm_bytes_per_launch": compulsory_bytes / launches,
             "cache": {"bytes_per_step": cache_bytes, "gbs": cache_bytes / (ms * 1e-3) / 1e9,
                       "what": "BVH node + triangle + entry-box bytes the walk reads (64 B/node, 64 or 160 B/triangle record, 28 B/entry box): served by L1/L2, "
                               "the scene is cache resident; NOT an HBM figure"},
             "ncu": ncu.get(kernel), "note": note}
        if extra:
            r.update(extra)
        return r

    roofs = {
        "rad_visibility_kernel": roof(
            "rad_visibility_kernel", stats["gpu_ms_rad_vis"], batches,
            stats["n_rad_segments"] * 12.0 + stats["n_rad_links"] * 12.0,
            (stats["n_ray_node_visits"] * 64.0 + stats["n_ray_tri_tests"] * 64.0 + stats.get("n_ray_entry_tests", 0) * 28.0) * rad_share + stats["n_rad_segments"] * 32.0,
            "issue+l1",
            "any-hit walk of the radiosity candidates from each chunk's entry set.  HBM traffic = the 12 B candidate records in + 12 B per link out; "
            "everything the walk itself reads comes from L1/L2.  Limiter per ncu: instruction issue and L1 wavefronts under SIMT divergence.",
            {"segments_per_s": per_s(stats["n_rad_segments"], stats["gpu_ms_rad_vis"])}),
        "rad_candidates_kernel": roof(
            "rad_candidates_kernel", stats["gpu_ms_rad_pairs"], batches,
            stats["n_rad_tile_loads"] * 5120.0 + stats["n_rad_segments"] * 12.0, stats["n_rad_tile_loads"] * 5120.0,
            "fp32-issue",
            "pair sweep: 5 KiB per column tile staged by TMA bulk copy + 12 B per candidate written; FP32-issue bound (22 instructions per lumel pair per lane)",
            {"pair_tests_per_s": per_s(stats["n_rad_pairs"], stats["gpu_ms_rad_pairs"])}),
        "direct_march_kernel": roof(
            "direct_march_kernel", stats["gpu_ms_march"], 1,
            stats["n_marches"] * 36.0, stats["n_node_visits"] * 64.0 + stats["n_tri_tests"] * 160.0,
            "latency+divergence",
            "sphere-traced shadow march, each step a nearest-distance query: 36 B per march of HBM (lumel in, factor out), the walk is L1/L2 served",
            {"distance_queries_per_s": per_s(stats["n_distance_queries"], stats["gpu_ms_march"])}),
    }
    roofs = {k: v for k, v in roofs.items() if v}
    dominant = max(roofs, key=lambda k: roofs[k]["kernel_ms_per_step"]) if roofs else None
    roofline = roofs.get(dominant)

    # ---- end-to-end arm: public C API, host buffers in, host lightmaps out ---------------------------
    e2e_walls, e2e_stats, out_hash = [], None, ""
    for it in range(args.e2e_steps + 1):
        hh = api.BakeHandle(sc, device=local_rank, shard=shard, output_root_only=True)   # scene set-up (ltr_MeshAddPart...) is outside the timed region; the lightmaps land on rank 0
        barrier()
        w = hh.run()
        wmax = reduce_max(w)
        if it > 0:                                                     # first one warms allocator / NCCL communicator
            e2e_walls.append(wmax)
            e2e_stats = hh.stats()
        if it == args.e2e_steps and rank == 0:
            out_hash = hh.output_hash()
        hh.close()
    e2e_rays = reduce_sum(rays_of(e2e_stats))
    e2e_val = e2e_rays * len(e2e_walls) / sum(e2e_walls)
    host_s = {k: reduce_max(e2e_stats[k]) for k in ("t_prexform", "t_accel", "t_upload", "t_samples", "t_direct", "t_radiosity", "t_ao", "t_finalize", "t_readback")}

    expected = load_json("tests", "golden", "bench_hashes.json").get(args.workload)
    parity_block = {"lightmap_fnv1a64": out_hash, "expected_single_gpu": expected, "match": (out_hash == expected) if expected else None,
                    "what": "FNV-1a-64 of all output lightmaps (+ probes) of the last end-to-end bake on rank 0; the expected value was recorded from a "
                            "single-GPU bake and the reference-parity tests (tests/test_gpu_parity.py) pin single-GPU bakes against the reference"}

    line = {
        "metric": "rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "step_ms": [round(x, 3) for x in step_ms], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "name": args.workload, "triangles": sc.triangle_count(),
                   "lumels": int(stats["n_lumels_total"]), "lights": len(sc.lights), "parallelism": f"lumel-shard x{world}",
                   "l2_policy": "inputs (>=126 MB of lumel/link arrays per step) exceed L2; every step regenerates all buffers"},
        "bake_wall_s": sum(e2e_walls) / len(e2e_walls),
        "bake_device_s": total_ms / args.steps * 1e-3,
        "rays_per_step": rays_step,
        "rates": rates,
        "stage_ms": stage_ms,
        "counters": {k: int(v) for k, v in job.items()},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": int(e2e_stats["h2d_bytes"]), "d2h_bytes_per_step": int(e2e_stats["d2h_bytes"]),
                "bake_wall_s": sum(e2e_walls) / len(e2e_walls), "steps": len(e2e_walls), "host_s": host_s,
                "result": "every lightmap on the host of rank 0 (ltrx_SetOutputRoot); d2h_bytes_per_step is rank 0's"},
        "parity": parity_block,
        "gpu_launches": int(launches_step * args.steps),
        "roofline": roofline,
        "roofline_other": {k: v for k, v in roofs.items() if k != dominant},
    }
    h.close()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ra = argparse.Namespace(**vars(args))
            ra.steps, ra.warmup = 2, 0
            cb = reference_arm(ra)["cpu_baseline"]
            sib = CPU_SIBLING.get(args.workload)
            if sib:
                cb["per_unit"] = cpu_unit_rates(sib, stats, sc)
                cb["cpu_extrapolated_wall_s"] = cb["per_unit"]["extrapolated_full_workload"]["cpu_extrapolated_wall_s"]
            line["cpu_baseline"] = cb
            # same-work comparison: both sides bake the SAME scene completely, wall / wall through ltr_Start -> ltr_GetStatus
            same = []
            for name in SAME_CONFIG.get(args.workload, []):
                walls, cores, ssc = reference_walls(name, 2, 0)
                gw = []
                for it in range(3):
                    with api.BakeHandle(ssc, device=local_rank) as g:
                        w = g.run()
                    if it:
                        gw.append(w)
                same.append({"workload": name, "same_config": True, "cpu_wall_s": sum(walls) / len(walls), "cpu_cores": cores,
                             "gpu_wall_s": sum(gw) / len(gw), "speedup_wall": (sum(walls) / len(walls)) / (sum(gw) / len(gw))})
            line["same_config"] = same
        except Exception as e:                              # noqa: BLE001
            line["cpu_baseline"] = {"error": str(e)}
    fail = reduce_max(1.0 if (rank == 0 and expected and out_hash != expected and not args.record_hash) else 0.0)
    if rank == 0:
        if fail:
            print(json.dumps(line), file=sys.stderr, flush=True)
            print(f"bench.py: PARITY FAILURE: lightmap hash {out_hash} != {expected} recorded for {args.workload} on one GPU", file=sys.stderr, flush=True)
        else:
            print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if fail:
        sys.exit(3)


if __name__ == "__main__":
    main()
