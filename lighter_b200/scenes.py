"""Scene descriptions for the bake path: the reference's own test scenarios and the synthetic
BASELINE.json configs, in a form that can be (a) fed to liblighter_b200.so through the ltr_* C ABI
(lighter_b200.api) and (b) written as an LTRSCN01 file for the C++ drivers (oracle/bake_driver.cpp,
which runs the same scene through the unmodified reference).

Scenario parameters restate the reference's test driver: basic (lighter_test.cpp:183-257),
hugeoverlap (:259-306), mesh1 (:308-368), mesh2 (:370-468), rad1 (:482-544).  The mesh loader
restates lighter_test.cpp:28-108 (OBJ-like text, v flipped to 1-v, vertices de-duplicated on exact
(P,N,T) equality in first-seen order).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field

import numpy as np

LT_POINT, LT_SPOT, LT_DIRECT = 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
MESH_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden", "meshes")


# --------------------------------------------------------------------------------------------
# containers
# --------------------------------------------------------------------------------------------
@dataclass
class Part:
    pos: np.ndarray          # (V,3) f32
    nrm: np.ndarray          # (V,3) f32
    uv1: np.ndarray          # (V,2) f32
    uv2: np.ndarray          # (V,2) f32  lightmap UVs
    idx: np.ndarray          # (I,)  u32
    shadow: int = 1


@dataclass
class Mesh:
    ident: str
    parts: list = field(default_factory=list)


@dataclass
class Instance:
    mesh: int
    matrix: np.ndarray       # (4,4) f32, row-vector convention (v * M)
    importance: float = 1.0
    shadow: int = 1
    ident: str = ""
    force_size: tuple = (0, 0)


@dataclass
class Light:
    type: int
    position: tuple = (0.0, 0.0, 0.0)
    direction: tuple = (0.0, 0.0, 0.0)
    up_direction: tuple = (0.0, 0.0, 0.0)
    color_rgb: tuple = (1.0, 1.0, 1.0)
    range: float = 1.0
    power: float = 1.0
    light_radius: float = 0.1
    shadow_sample_count: int = 1
    spot_angle_out: float = 0.0
    spot_angle_in: float = 0.0
    spot_curve: float = 0.0

    def pack(self) -> bytes:
        return struct.pack("<I12f3fi3f", self.type, *self.position, *self.direction, *self.up_direction,
                           *self.color_rgb, self.range, self.power, self.light_radius,
                           self.shadow_sample_count, self.spot_angle_out, self.spot_angle_in, self.spot_curve)


def default_config() -> dict:
    """Defaults of ltr_GetConfig(cfg, NULL) (lighter.cpp:1179-1207)."""
    return dict(
        max_lightmap_size=1024, default_width=64, default_height=64, global_size_factor=4.0,
        max_correct_dist=0.1, max_correct_angle=60.0, clear_color=(0.0, 0.0, 0.0),
        ambient_color=(0.0, 0.0, 0.0), bounce_count=0, sample_fn_kind=0, ao_distance=0.0,
        ao_multiplier=1.2, ao_falloff=1.0, ao_effect=0.0, ao_divergence=0.0, ao_color=(0.0, 0.0, 0.0),
        ao_num_samples=17, blur_size=0.5, ds2x=0, generate_normalmap_data=0, size_fn_kind=0,
    )


@dataclass
class Scene:
    name: str = "scene"
    cfg: dict = field(default_factory=default_config)
    meshes: list = field(default_factory=list)
    instances: list = field(default_factory=list)
    lights: list = field(default_factory=list)
    probes: list = field(default_factory=list)   # (id, pos3, nrm3)

    # -- counts --------------------------------------------------------------------------------
    def triangle_count(self) -> int:
        per_mesh = [sum(len(p.idx) // 3 for p in m.parts) for m in self.meshes]
        return sum(per_mesh[i.mesh] for i in self.instances)

    # -- LTRSCN01 writer (layout documented in oracle/scene_io.h) -------------------------------
    def write(self, path: str) -> str:
        c = self.cfg
        with open(path, "wb") as f:
            f.write(b"LTRSCN01")
            f.write(struct.pack("<3I3f3f3fii5f3fifiii",
                                c["max_lightmap_size"], c["default_width"], c["default_height"],
                                c["global_size_factor"], c["max_correct_dist"], c["max_correct_angle"],
                                *c["clear_color"], *c["ambient_color"], c["bounce_count"], c["sample_fn_kind"],
                                c["ao_distance"], c["ao_multiplier"], c["ao_falloff"], c["ao_effect"],
                                c["ao_divergence"], *c["ao_color"], c["ao_num_samples"], c["blur_size"],
                                c["ds2x"], c["generate_normalmap_data"], c["size_fn_kind"]))
            f.write(struct.pack("<I", len(self.meshes)))
            for m in self.meshes:
                b = m.ident.encode()
                f.write(struct.pack("<I", len(b)) + b + struct.pack("<I", len(m.parts)))
                for p in m.parts:
                    f.write(struct.pack("<IIi", len(p.pos), len(p.idx), p.shadow))
                    for a, dt in ((p.pos, "<f4"), (p.nrm, "<f4"), (p.uv1, "<f4"), (p.uv2, "<f4"), (p.idx, "<u4")):
                        f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
            f.write(struct.pack("<I", len(self.instances)))
            for i in self.instances:
                b = i.ident.encode()
                f.write(struct.pack("<I", i.mesh))
                f.write(np.ascontiguousarray(i.matrix, dtype="<f4").tobytes())
                f.write(struct.pack("<fiI", i.importance, i.shadow, len(b)) + b)
                f.write(struct.pack("<II", *i.force_size))
            f.write(struct.pack("<I", len(self.lights)))
            for lt in self.lights:
                f.write(lt.pack())
            f.write(struct.pack("<I", len(self.probes)))
            for pid, pos, nrm in self.probes:
                f.write(struct.pack("<I3f3f", pid, *pos, *nrm))
        return path


# --------------------------------------------------------------------------------------------
# LTROUT01 reader (layout documented in oracle/bake_driver.cpp)
# --------------------------------------------------------------------------------------------
def read_output(path: str) -> dict:
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"LTROUT01", "not an LTROUT01 file"
    off = 8
    wall, threads, nlm = struct.unpack_from("<dII", buf, off)
    off += 16
    out = dict(wall_s=wall, threads=threads, lightmaps=[], probes=None, instances=None, links=None)
    for _ in range(nlm):
        uid, w, h, has_n = struct.unpack_from("<4I", buf, off)
        off += 16
        rgb = np.frombuffer(buf, "<f4", w * h * 3, off).reshape(h, w, 3).copy()
        off += w * h * 12
        nrm = None
        if has_n:
            nrm = np.frombuffer(buf, "<f4", w * h * 4, off).reshape(h, w, 4).copy()
            off += w * h * 16
        out["lightmaps"].append(dict(uid=uid, width=w, height=h, rgb=rgb, normals=nrm))
    (npr,) = struct.unpack_from("<I", buf, off)
    off += 4
    out["probes"] = np.frombuffer(buf, "<f4", npr * 3, off).reshape(npr, 3).copy()
    off += npr * 12
    (has_int,) = struct.unpack_from("<I", buf, off)
    off += 4
    if has_int:
        (ni,) = struct.unpack_from("<I", buf, off)
        off += 4
        insts = []
        for _ in range(ni):
            w, h, n = struct.unpack_from("<3I", buf, off)
            off += 12
            d = dict(width=w, height=h, n=n)
            for key, dt, k in (("pos", "<f4", 3), ("nrm", "<f4", 3), ("loc", "<u4", 1), ("radinfo", "<f4", 4), ("rgb", "<f4", 3)):
                a = np.frombuffer(buf, dt, n * k, off).copy()
                off += n * k * 4
                d[key] = a.reshape(n, k) if k > 1 else a
            insts.append(d)
        out["instances"] = insts
        (rows,) = struct.unpack_from("<I", buf, off)
        off += 4
        linkmap = np.frombuffer(buf, "<u4", rows * 2, off).reshape(rows, 2).copy()
        off += rows * 8
        (nl,) = struct.unpack_from("<I", buf, off)
        off += 4
        rec = np.frombuffer(buf, np.dtype([("other", "<u4"), ("factor", "<f4")]), nl, off).copy()
        out["links"] = dict(map=linkmap, other=rec["other"], factor=rec["factor"])
    return out


# --------------------------------------------------------------------------------------------
# mesh loading
# --------------------------------------------------------------------------------------------
def parse_data_mesh(path: str) -> Part:
    """OBJ-like text loader, semantics of lighter_test.cpp:48-107 (v flipped, exact-match dedup)."""
    plist, nlist, tlist = [], [], []
    pos, nrm, uv, idx = [], [], [], []
    seen = {}
    one = np.float32(1.0)
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                plist.append(tuple(np.float32(x) for x in t[1:4]))
            elif t[0] == "vt":
                u, v = np.float32(t[1]), np.float32(t[2])
                tlist.append((u, np.float32(one - v)))
            elif t[0] == "vn":
                nlist.append(tuple(np.float32(x) for x in t[1:4]))
            elif t[0] == "f":
                for corner in t[1:4]:
                    a, b, c = (int(x) - 1 for x in corner.split("/"))
                    key = tuple(float(x) for x in (*plist[a], *nlist[c], *tlist[b]))
                    k = seen.get(key)
                    if k is None:
                        k = len(pos)
                        seen[key] = k
                        pos.append(plist[a]); nrm.append(nlist[c]); uv.append(tlist[b])
                    idx.append(k)
    uv = np.array(uv, np.float32)
    return Part(np.array(pos, np.float32), np.array(nrm, np.float32), uv, uv.copy(), np.array(idx, np.uint32), 1)


def load_mesh(name: str) -> Part:
    """Load one of the bundled test meshes from tests/golden/meshes/<name>.npz (converted from the
    reference's bin/<name>.data by tools/import_reference_meshes.py)."""
    z = np.load(os.path.join(MESH_DIR, name + ".npz"))
    return Part(z["pos"], z["nrm"], z["uv"], z["uv"].copy(), z["idx"], 1)


IDENTITY = np.eye(4, dtype=np.float32)


def _libm_f(fn: str, x: float) -> float:
    """Call the host libm's float function (glibc cosf/sinf/...) so scene constants computed in the
    reference's C driver are reproduced bit for bit."""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    f = getattr(libm, fn)
    f.restype = ctypes.c_float
    f.argtypes = [ctypes.c_float]
    return float(f(ctypes.c_float(x)))


# --------------------------------------------------------------------------------------------
# the reference's own scenarios
# --------------------------------------------------------------------------------------------
def _quad_part(index_repeats: int = 1) -> Part:
    pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    nrm = np.array([[0, 0, 1]] * 4, np.float32)
    uv = np.array([[0.1, 0.1], [0.9, 0.1], [0.9, 0.9], [0.1, 0.9]], np.float32)
    idx = np.array([0, 1, 2, 2, 3, 0] * index_repeats, np.uint32)
    return Part(pos, nrm, uv, uv.copy(), idx, 1)


def scene_basic() -> Scene:
    s = Scene("basic")
    s.meshes.append(Mesh("mesh1", [_quad_part()]))
    c = np.float32(_libm_f("cosf", 0.5)) * np.float32(0.2)     # float cos(0.5f) * 0.2f as in the C driver
    sn = np.float32(_libm_f("sinf", 0.5)) * np.float32(0.2)
    m2 = np.array([[c, sn, 0, 0], [-sn, c, 0, 0], [0, 0, 0.2, 0], [0, 0, 0.5, 1]], np.float32)
    m3 = np.array([[1.9, 0, 0, 0], [0, 1.9, 0, 0], [0, 0, 1.9, 0], [0, 0, -0.5, 1]], np.float32)
    for m in (IDENTITY, m2, m3):
        s.instances.append(Instance(0, m.copy(), 1.0, 1, ""))
    s.lights.append(Light(LT_POINT, (-0.4, -0.4, 1.0), color_rgb=(0.9, 0.1, 0.05), range=4.0, power=1.0,
                          light_radius=0.1, shadow_sample_count=9))
    return s


def scene_hugeoverlap() -> Scene:
    s = Scene("hugeoverlap")
    s.meshes.append(Mesh("mesh1", [_quad_part(16)]))
    for _ in range(10):
        s.instances.append(Instance(0, IDENTITY.copy(), 1.0, 1, ""))
    s.lights.append(Light(LT_POINT, (-0.4, -0.4, 1.0), color_rgb=(0.9, 0.1, 0.05), range=4.0, power=1.0,
                          light_radius=0.1, shadow_sample_count=9))
    return s


def _mesh12_lights(radius: float) -> list:
    return [
        Light(LT_POINT, (-2.18, -4.04, 1.40), color_rgb=(0.9, 0.7, 0.5), range=16.0, power=1.0, light_radius=radius, shadow_sample_count=5),
        Light(LT_POINT, (2.18, 4.04, 1.40), color_rgb=(0.5, 0.7, 0.9), range=16.0, power=1.0, light_radius=radius, shadow_sample_count=5),
        Light(LT_SPOT, (0.0, 0.0, 1.60), (0.0, 0.0, -1.0), (1.0, 0.0, 0.0), (0.7, 0.1, 0.05), 16.0, 1.0, radius, 5, 45.0, 25.0, 0.5),
    ]


def scene_mesh1() -> Scene:
    """BASELINE config 1."""
    s = Scene("mesh1")
    s.cfg["ao_distance"] = 2.0
    s.meshes.append(Mesh("mesh1", [load_mesh("test-mesh")]))
    s.instances.append(Instance(0, IDENTITY.copy(), 1.0, 1, ""))
    s.lights = _mesh12_lights(0.1)
    return s


def scene_mesh2(normalmap: int = 1) -> Scene:
    """BASELINE config 2."""
    s = Scene("mesh2")
    s.cfg.update(ao_distance=2.0, global_size_factor=4.0, blur_size=0.0, generate_normalmap_data=normalmap)
    s.meshes.append(Mesh("mesh1", [load_mesh("test-set2-mesh1")]))
    s.meshes.append(Mesh("mesh2", [load_mesh("test-set2-mesh2")]))
    s.instances.append(Instance(0, IDENTITY.copy(), 1.0, 1, ""))
    s.instances.append(Instance(1, IDENTITY.copy(), 0.3, 1, ""))
    s.lights = _mesh12_lights(0.2)
    s.lights.append(Light(LT_DIRECT, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), color_rgb=(0.6, 0.55, 0.5), range=1000.0,
                          power=1.0, light_radius=0.2, shadow_sample_count=1))
    return s


def scene_rad1() -> Scene:
    s = Scene("rad1")
    s.cfg.update(ao_distance=2.0, bounce_count=2, sample_fn_kind=1)
    s.meshes.append(Mesh("mesh1", [load_mesh("test-set2-mesh1")]))
    s.instances.append(Instance(0, IDENTITY.copy(), 0.4, 1, ""))
    s.lights.append(Light(LT_POINT, (10.18, 0.0, 1.40), color_rgb=(0.99, 0.97, 0.95), range=24.0, power=1.0,
                          light_radius=0.1, shadow_sample_count=5))
    return s


NAMED = {
    "basic": scene_basic, "hugeoverlap": scene_hugeoverlap, "mesh1": scene_mesh1,
    "mesh2": scene_mesh2, "rad1": scene_rad1,
}
