"""Parity metrics and reference-runner helpers shared by tests/, bench.py and tools/.

The 8-bit texel rule is the reference test driver's TGA quantisation (lighter_test.cpp:111-164):
`(unsigned char) min(v * 255, 255)` -- truncation, not rounding.  BASELINE.json's bar: mean
absolute error <= 1/255 and |diff| <= 2/255 on >= 99.5 % of texels.
"""
from __future__ import annotations

import os
import subprocess
import tempfile

import numpy as np

from . import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BAKE = os.path.join(ROOT, "oracle", "_ref", "ref_bake")
REF_PRIMS = os.path.join(ROOT, "oracle", "_ref", "libref_prims.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")


def quantize8(img: np.ndarray) -> np.ndarray:
    v = np.minimum(np.asarray(img, np.float32) * np.float32(255.0), np.float32(255.0))
    v = np.where(np.isfinite(v), v, 0)
    return np.clip(np.trunc(v), 0, 255).astype(np.int32)


def texel_parity(ours: np.ndarray, ref: np.ndarray) -> dict:
    """Per-texel (all channels) 8-bit comparison."""
    a, b = quantize8(ours), quantize8(ref)
    d = np.abs(a - b)
    per_texel = d.reshape(-1, d.shape[-1]).max(axis=1) if d.size else np.zeros(0, np.int32)
    return dict(
        mae=float(d.mean()) if d.size else 0.0,
        within2=float((per_texel <= 2).mean()) if per_texel.size else 1.0,
        max=int(d.max()) if d.size else 0,
        float_max_abs=float(np.abs(np.asarray(ours, np.float32) - np.asarray(ref, np.float32)).max()) if d.size else 0.0,
        bit_exact=bool(np.array_equal(np.asarray(ours, np.float32).view(np.uint32), np.asarray(ref, np.float32).view(np.uint32))),
    )


def meets_bar(p: dict) -> bool:
    return p["mae"] <= 1.0 and p["within2"] >= 0.995


def have_reference() -> bool:
    return os.path.exists(REF_BAKE)


def run_reference(scene: scenes.Scene, threads: int = 1, internals: bool = True, repeat: int = 1, timeout: float = 3600) -> dict:
    """Run the UNMODIFIED reference (oracle/_ref/ref_bake, compiled from /root/reference by
    oracle/Makefile) on the scene, on host cores.  threads=1 makes its AO deterministic."""
    if not have_reference():
        raise RuntimeError("oracle/_ref/ref_bake is missing: run `make -C oracle ref` where /root/reference exists")
    with tempfile.TemporaryDirectory() as td:
        sp, op = os.path.join(td, "scene.bin"), os.path.join(td, "out.bin")
        scene.write(sp)
        cmd = [REF_BAKE, sp, op, "--quiet", "--repeat", str(repeat)]
        if threads > 0:
            cmd += ["--threads", str(threads)]
        if internals:
            cmd += ["--internals"]
        subprocess.run(cmd, check=True, timeout=timeout)
        return scenes.read_output(op)


REF_STAGES = ("transforming spatial data", "generating data structures", "generating samples", "rendering lightmaps", "calculating radiosity",
              "bouncing light", "committing radiosity", "rendering ambient occlusion", "exporting lightmaps")


def run_reference_timed(scene: scenes.Scene, threads: int = 0, timeout: float = 3600) -> dict:
    """One bake of the UNMODIFIED reference with its stage transitions time-stamped by the driver's 200 us polling loop
    (oracle/bake_driver.cpp prints `[ t s] stage` on stderr): wall seconds per reference stage (lighter.cpp:1052-1138)."""
    if not have_reference():
        raise RuntimeError("oracle/_ref/ref_bake is missing: run `make -C oracle ref` where /root/reference exists")
    import re
    with tempfile.TemporaryDirectory() as td:
        sp, op = os.path.join(td, "scene.bin"), os.path.join(td, "out.bin")
        scene.write(sp)
        cmd = [REF_BAKE, sp, op]
        if threads > 0:
            cmd += ["--threads", str(threads)]
        r = subprocess.run(cmd, check=True, timeout=timeout, capture_output=True, text=True)
        out = scenes.read_output(op)
    marks = [(float(t), name.strip()) for t, name in re.findall(r"\[\s*([0-9.]+)s\]\s+(.*)", r.stderr)]
    stage_s = {}
    for k, (t, name) in enumerate(marks):
        end = marks[k + 1][0] if k + 1 < len(marks) else out["wall_s"]
        stage_s[name] = stage_s.get(name, 0.0) + max(end - t, 0.0)
    return dict(wall_s=out["wall_s"], threads=out["threads"], stage_s=stage_s, lightmaps=out["lightmaps"])


def fnv1a64(data: bytes) -> int:
    h = 0xcbf29ce484222325
    for x in data:
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h
