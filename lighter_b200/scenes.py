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


def scene_corner(normalmap: int = 0) -> Scene:
    """Three separate instances meeting in one concave corner (floor, wall at x=0, wall at y=0) plus a low step lying on
    the floor: lumels next to an edge are pushed by one instance's triangles and then tested against the next instance
    at the MOVED position (ref: lighter.cpp:445-446, lighter_math.cpp:1040-1044).  Not a scenario of the reference's own
    driver; parity fixtures for it come from the live reference."""
    s = Scene("corner")
    s.cfg.update(ao_distance=1.0, blur_size=0.0, generate_normalmap_data=normalmap, global_size_factor=6.0)
    const = lambda v: (lambda p: np.tile(np.asarray(v, np.float64), (len(p), 1)))
    g = 0.04
    floor = _grid_patch(np.array([0.0, 0.0, 0.0]), np.array([4.0, 0, 0]), np.array([0, 4.0, 0]), 5, 5, const((0, 0, 1)), (g, g, 1 - g, 1 - g))
    wallx = _grid_patch(np.array([0.0, 4.0, 0.0]), np.array([0, -4.0, 0]), np.array([0, 0, 3.0]), 4, 3, const((1, 0, 0)), (g, g, 1 - g, 1 - g))
    wally = _grid_patch(np.array([0.0, 0.0, 0.0]), np.array([4.0, 0, 0]), np.array([0, 0, 3.0]), 3, 4, const((0, 1, 0)), (g, g, 1 - g, 1 - g))
    for k, (name, (pos, nrm, uv, idx)) in enumerate((("floor", floor), ("wallx", wallx), ("wally", wally))):
        s.meshes.append(Mesh(name, [Part(pos, nrm, uv.copy(), uv, idx, 1)]))
        s.instances.append(Instance(k, IDENTITY.copy(), 1.0, 1, name))
    # the step: a box 1.2 x 0.9 x 0.35 standing in the corner region, five faces, one instance, rotated a little about z
    x0, x1, y0, y1, z1 = 0.0, 1.2, 0.0, 0.9, 0.35
    faces = [(np.array([x0, y0, 0.0]), np.array([x1 - x0, 0, 0.0]), np.array([0, 0, z1]), (0, -1, 0)),
             (np.array([x1, y0, 0.0]), np.array([0, y1 - y0, 0.0]), np.array([0, 0, z1]), (1, 0, 0)),
             (np.array([x1, y1, 0.0]), np.array([x0 - x1, 0, 0.0]), np.array([0, 0, z1]), (0, 1, 0)),
             (np.array([x0, y1, 0.0]), np.array([0, y0 - y1, 0.0]), np.array([0, 0, z1]), (-1, 0, 0)),
             (np.array([x0, y0, z1]), np.array([x1 - x0, 0, 0.0]), np.array([0, y1 - y0, 0.0]), (0, 0, 1))]
    pos, nrm, uv, idx, base = [], [], [], [], 0
    for k, (o, du, dv, nn) in enumerate(faces):
        p, n_, t, i = _grid_patch(o, du, dv, 2, 2, const(nn), (k * 0.2 + 0.02, 0.02, (k + 1) * 0.2 - 0.02, 0.5))
        pos.append(p); nrm.append(n_); uv.append(t); idx.append(i + base); base += len(p)
    s.meshes.append(Mesh("step", [Part(np.concatenate(pos), np.concatenate(nrm), np.concatenate(uv).copy(), np.concatenate(uv), np.concatenate(idx), 1)]))
    c, sn = np.float32(np.cos(0.3)), np.float32(np.sin(0.3))
    m = np.array([[c, sn, 0, 0], [-sn, c, 0, 0], [0, 0, 1, 0], [0.9, 0.6, 0.0, 1]], np.float32)
    s.instances.append(Instance(3, m, 1.0, 1, "step"))
    s.lights.append(Light(LT_POINT, (2.5, 2.2, 2.4), color_rgb=(0.9, 0.8, 0.6), range=9.0, power=1.0, light_radius=0.15, shadow_sample_count=4))
    s.lights.append(Light(LT_SPOT, (3.4, 0.8, 2.0), (-0.6, 0.1, -0.8), (1.0, 0.0, 0.0), (0.2, 0.5, 0.9), 8.0, 1.0, 0.1, 4, 50.0, 20.0, 0.7))
    return s


NAMED = {
    "corner": scene_corner,
    "basic": scene_basic, "hugeoverlap": scene_hugeoverlap, "mesh1": scene_mesh1,
    "mesh2": scene_mesh2, "rad1": scene_rad1,
}


# --------------------------------------------------------------------------------------------
# synthetic scenes of the BASELINE.json configs 3-5: seeded tile generator
# --------------------------------------------------------------------------------------------
def _height(x, y):
    return np.float32(0.25) * np.sin(np.float32(0.9) * x) * np.cos(np.float32(0.7) * y)


def _grid_patch(origin, du, dv, nu, nv, normal_fn, uv_rect, pos_fn=None):
    """(nu x nv) quads spanning origin + s*du + t*dv, s,t in [0,1]; CCW w.r.t. cross(du, dv).
    uv_rect = (u0, v0, u1, v1) receives the patch.  Returns pos, nrm, uv, idx (local indices)."""
    s, t = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="xy")
    s, t = s.ravel(), t.ravel()
    pos = origin[None, :] + s[:, None] * du[None, :] + t[:, None] * dv[None, :]
    if pos_fn is not None:
        pos = pos_fn(pos)
    nrm = normal_fn(pos)
    u0, v0, u1, v1 = uv_rect
    uv = np.stack([u0 + s * (u1 - u0), v0 + t * (v1 - v0)], 1)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
    a = (j * (nu + 1) + i).ravel()
    b, c, d = a + 1, a + nu + 2, a + nu + 1
    idx = np.stack([a, b, c, c, d, a], 1).ravel()
    return pos.astype(np.float32), nrm.astype(np.float32), uv.astype(np.float32), idx.astype(np.uint32)


def _tile_part(rng, ox, oy, size, floor_n, ceil_n, height, ceiling, pillar, walls, pillar_q=6, wall_q=6, uv_rect=(0.0, 0.0, 1.0, 1.0),
               world_coords=False):
    """One tile: heightfield floor, optional ceiling, optional pillar, boundary walls.  Geometry is in
    tile-local x,y (translated by the instance matrix) unless world_coords."""
    g = 3.0 / 256.0
    U0, V0, U1, V1 = uv_rect

    def rect(a, b, c, d):            # sub-rectangle of this tile's atlas square
        return (U0 + a * (U1 - U0), V0 + b * (V1 - V0), U0 + c * (U1 - U0), V0 + d * (V1 - V0))

    lx, ly = (ox, oy) if world_coords else (0.0, 0.0)
    chunks = []

    def floor_pos(p):
        p = p.copy()
        wx, wy = p[:, 0] + (0.0 if world_coords else ox), p[:, 1] + (0.0 if world_coords else oy)
        p[:, 2] = _height(wx.astype(np.float32), wy.astype(np.float32))
        return p

    def floor_nrm(p):
        wx, wy = p[:, 0] + (0.0 if world_coords else ox), p[:, 1] + (0.0 if world_coords else oy)
        dzdx = 0.25 * 0.9 * np.cos(0.9 * wx) * np.cos(0.7 * wy)
        dzdy = -0.25 * 0.7 * np.sin(0.9 * wx) * np.sin(0.7 * wy)
        n = np.stack([-dzdx, -dzdy, np.ones_like(wx)], 1)
        return n / np.linalg.norm(n, axis=1, keepdims=True)

    const = lambda v: (lambda p: np.tile(np.asarray(v, np.float64), (len(p), 1)))
    chunks.append(_grid_patch(np.array([lx, ly, 0.0]), np.array([size, 0, 0.0]), np.array([0, size, 0.0]), floor_n, floor_n, floor_nrm,
                              rect(g, g, 0.60, 0.60), floor_pos))
    if ceiling:
        chunks.append(_grid_patch(np.array([lx, ly + size, height]), np.array([size, 0, 0.0]), np.array([0, -size, 0.0]), ceil_n, ceil_n,
                                  const((0, 0, -1)), rect(0.62 + g, g, 1 - g, 0.36)))
    if pillar:
        pw = rng.uniform(0.6, 1.6)
        ph = rng.uniform(2.0, max(2.5, height - 0.5))
        cx, cy = lx + size * rng.uniform(0.3, 0.7), ly + size * rng.uniform(0.3, 0.7)
        x0, x1, y0, y1, z0, z1 = cx - pw / 2, cx + pw / 2, cy - pw / 2, cy + pw / 2, -0.3, ph
        faces = [  # origin, du, dv, outward normal
            (np.array([x0, y0, z0]), np.array([pw, 0, 0.0]), np.array([0, 0, z1 - z0]), (0, -1, 0)),
            (np.array([x1, y0, z0]), np.array([0, pw, 0.0]), np.array([0, 0, z1 - z0]), (1, 0, 0)),
            (np.array([x1, y1, z0]), np.array([-pw, 0, 0.0]), np.array([0, 0, z1 - z0]), (0, 1, 0)),
            (np.array([x0, y1, z0]), np.array([0, -pw, 0.0]), np.array([0, 0, z1 - z0]), (-1, 0, 0)),
            (np.array([x0, y0, z1]), np.array([pw, 0, 0.0]), np.array([0, pw, 0.0]), (0, 0, 1)),
        ]
        for k, (o, du, dv, nn) in enumerate(faces):
            chunks.append(_grid_patch(o, du, dv, pillar_q, pillar_q, const(nn), rect(k * 0.19 + g, 0.64, (k + 1) * 0.19 - g, 0.80)))
    wall_defs = [  # west, east, south, north: inward normals
        (np.array([lx, ly + size, -0.3]), np.array([0, -size, 0.0]), (1, 0, 0)),
        (np.array([lx + size, ly, -0.3]), np.array([0, size, 0.0]), (-1, 0, 0)),
        (np.array([lx, ly, -0.3]), np.array([size, 0, 0.0]), (0, 1, 0)),
        (np.array([lx + size, ly + size, -0.3]), np.array([-size, 0, 0.0]), (0, -1, 0)),
    ]
    for k, on in enumerate(walls):
        if on:
            o, du, nn = wall_defs[k]
            chunks.append(_grid_patch(o, du, np.array([0, 0, height + 0.3]), wall_q, wall_q, const(nn), rect(k * 0.25 + g, 0.83, (k + 1) * 0.25 - g, 1 - g)))
    pos, nrm, uv, idx, base = [], [], [], [], 0
    for p, n, t, i in chunks:
        pos.append(p); nrm.append(n); uv.append(t); idx.append(i + base); base += len(p)
    pos, nrm, uv, idx = np.concatenate(pos), np.concatenate(nrm), np.concatenate(uv), np.concatenate(idx)
    return Part(pos, nrm, uv.copy(), uv, idx, 1)


def scene_synthetic(name: str, tiles: int, tile_size: float, target_tris: int, lm_size: int, ceiling: bool, height: float = 6.0,
                    merge: bool = False, seed: int = 20261017, pillar_fraction: float = 0.7) -> Scene:
    """tiles x tiles instances (or one merged instance), each with a forced lm_size^2 lightmap
    (merge: one (tiles*lm_size)^2 lightmap).  Floor resolution is tuned so the total triangle count
    lands within 1 % of target_tris."""
    rng = np.random.Generator(np.random.MT19937(seed))
    nt = tiles * tiles
    pillars = rng.uniform(0, 1, nt) < pillar_fraction
    ceil_n, pq, wq = 8, 6, 6
    fixed = 0
    for k in range(nt):
        i, j = k % tiles, k // tiles
        fixed += (2 * ceil_n * ceil_n if ceiling else 0) + (10 * pq * pq if pillars[k] else 0)
        fixed += 2 * wq * wq * ((i == 0) + (i == tiles - 1) + (j == 0) + (j == tiles - 1))
    per_tile_floor = max(2.0, (target_tris - fixed) / nt)
    m0 = max(1, int(np.floor(np.sqrt(per_tile_floor / 2))))
    floor_n = np.full(nt, m0)
    total = fixed + 2 * m0 * m0 * nt
    k = 0
    while total + (2 * (m0 + 1) ** 2 - 2 * m0 * m0) <= target_tris * 1.005 and k < nt:      # bump tiles one by one
        floor_n[k] = m0 + 1
        total += 2 * (m0 + 1) ** 2 - 2 * m0 * m0
        k += 1
    s = Scene(name)
    s.cfg["size_fn_kind"] = 1
    s.cfg["max_lightmap_size"] = 8192
    merged = []
    for k in range(nt):
        i, j = k % tiles, k // tiles
        ox, oy = i * tile_size, j * tile_size
        walls = (i == 0, i == tiles - 1, j == 0, j == tiles - 1)
        uv_rect = (i / tiles, j / tiles, (i + 1) / tiles, (j + 1) / tiles) if merge else (0.0, 0.0, 1.0, 1.0)
        part = _tile_part(rng, ox, oy, tile_size, int(floor_n[k]), ceil_n, height, ceiling, bool(pillars[k]), walls, pq, wq, uv_rect, world_coords=merge)
        if merge:
            merged.append(part)
            continue
        s.meshes.append(Mesh(f"tile{k}", [part]))
        m = IDENTITY.copy()
        m[3, 0], m[3, 1] = ox, oy
        s.instances.append(Instance(k, m, 1.0, 1, f"t{k}", (lm_size, lm_size)))
    if merge:
        s.meshes.append(Mesh("merged", merged))              # one part per tile
        s.instances.append(Instance(0, IDENTITY.copy(), 1.0, 1, "t0", (lm_size * tiles, lm_size * tiles)))
    return s


def _grid_lights(rng, nx, ny, extent, zlo, zhi, rng_range, n_spot_every=2, ssc=16) -> list:
    out = []
    for k in range(nx * ny):
        i, j = k % nx, k // nx
        x = (i + 0.5 + rng.uniform(-0.3, 0.3)) * extent / nx
        y = (j + 0.5 + rng.uniform(-0.3, 0.3)) * extent / ny
        z = rng.uniform(zlo, zhi)
        col = tuple(float(c) for c in rng.uniform(0.3, 1.0, 3))
        rad = float(rng.uniform(0.1, 0.3))
        if k % n_spot_every == 1:
            tilt = np.radians(rng.uniform(-15, 15, 2))
            d = np.array([np.sin(tilt[0]), np.sin(tilt[1]), -1.0])
            d /= np.linalg.norm(d)
            out.append(Light(LT_SPOT, (x, y, z), tuple(float(c) for c in d), (1.0, 0.0, 0.0), col, rng_range, 1.0, rad, ssc, 45.0, 25.0, 0.5))
        else:
            out.append(Light(LT_POINT, (x, y, z), color_rgb=col, range=rng_range, power=1.0, light_radius=rad, shadow_sample_count=ssc))
    return out


def scene_config3(tiles: int = 8, lm_size: int = 256, target_tris: int = 250_000, merge: bool = False, n_lights: tuple = (4, 8)) -> Scene:
    """BASELINE config 3: closed 40x40x6 interior, 250k tris, 2048^2 texels (64 x 256^2 or merged), 32 lights."""
    s = scene_synthetic("config3" + ("A" if merge else "B"), tiles, 40.0 / 8, target_tris, lm_size, ceiling=True, merge=merge)
    rng = np.random.Generator(np.random.MT19937(20261018))
    s.lights = _grid_lights(rng, n_lights[0], n_lights[1], tiles * 5.0, 4.0, 5.5, 14.0)
    s.cfg.update(ao_distance=0.0, bounce_count=0, blur_size=0.5)
    return s


def scene_config4(tiles: int = 16, lm_size: int = 256, target_tris: int = 1_000_000, bounces: int = 3, merge: bool = False) -> Scene:
    """BASELINE config 4: open-air 400x400 terrain with pillars, 1M tris, 4096^2 texels (256 x 256^2),
    8 local lights + 1 directional (range 60), AO 17 samples, 3 radiosity bounces."""
    s = scene_synthetic("config4", tiles, 25.0, target_tris, lm_size, ceiling=False, height=6.0, merge=merge)
    rng = np.random.Generator(np.random.MT19937(20261019))
    gx, gy = (2, 4) if tiles >= 8 else (1, 2)
    s.lights = _grid_lights(rng, gx, gy, tiles * 25.0, 8.0, 12.0, 60.0)
    d = np.array([0.4, 0.3, 0.85]); d /= np.linalg.norm(d)
    s.lights.append(Light(LT_DIRECT, (0.0, 0.0, 0.0), tuple(float(c) for c in d), color_rgb=(0.6, 0.55, 0.5), range=60.0, power=1.0,
                          light_radius=0.2, shadow_sample_count=16))
    s.cfg.update(ao_distance=2.0, ao_num_samples=17, bounce_count=bounces, blur_size=0.5)
    return s


def scene_config5(n_lights: int = 16, samples: int = 16, tiles: int = 8, lm_size: int = 128, target_tris: int = 500_000) -> Scene:
    """BASELINE config 5 (sweep member): one merged 500k-tri instance, 1024^2 lightmap, L lights."""
    s = scene_synthetic("config5", tiles, 5.0, target_tris, lm_size, ceiling=True, merge=True)
    rng = np.random.Generator(np.random.MT19937(20261020))
    nx = int(np.ceil(np.sqrt(n_lights)))
    ny = (n_lights + nx - 1) // nx
    s.lights = _grid_lights(rng, nx, ny, tiles * 5.0, 4.0, 5.5, 14.0, ssc=samples)[:n_lights]
    s.cfg.update(ao_distance=0.0, bounce_count=0, blur_size=0.5)
    return s


WORKLOADS = {
    # name: (builder, kwargs) -- full-size GPU workloads and the scaled-down siblings the CPU reference can finish
    "config3": (scene_config3, {}),
    "config3_sibling": (scene_config3, dict(tiles=2, lm_size=64, target_tris=250_000 // 16, n_lights=(2, 2))),
    "config4": (scene_config4, {}),
    "config4_sibling": (scene_config4, dict(tiles=2, lm_size=64, target_tris=1_000_000 // 64)),
    "config4_quarter": (scene_config4, dict(tiles=8, lm_size=256, target_tris=250_000)),
    "config5": (scene_config5, dict(n_lights=64, samples=16)),
    "mesh1": (scene_mesh1, {}),
    "mesh2": (scene_mesh2, {}),
}


def workload(name: str) -> Scene:
    fn, kw = WORKLOADS[name]
    return fn(**kw)
