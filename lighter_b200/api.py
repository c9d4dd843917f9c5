"""ctypes mirror of the ltr_* C ABI (include/lighter.h) and the ltrx_* extensions
(include/lighter_b200.h), plus `bake()`: feed a scenes.Scene through the C API exactly the way the
reference's own test driver does (lighter_test.cpp:167-180 polling loop, :187-256 call order).

This module is host-side plumbing only.  All computation happens inside liblighter_b200.so on the
GPU; if the shared library is missing, or there is no CUDA device, calls fail loudly -- there is no
Python or CPU fallback path.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import os
import time

import numpy as np

from . import scenes as _scenes

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblighter_b200.so")

u32 = C.c_uint32
VEC3 = C.c_float * 3
MAT4 = (C.c_float * 4) * 4


class MeshPartInfo(C.Structure):
    _fields_ = [("positions_f3", C.c_void_p), ("normals_f3", C.c_void_p), ("texcoords1_f2", C.c_void_p),
                ("texcoords2_f2", C.c_void_p), ("stride_positions", u32), ("stride_normals", u32),
                ("stride_texcoords1", u32), ("stride_texcoords2", u32), ("indices", C.c_void_p),
                ("vertex_count", u32), ("index_count", u32), ("shadow", C.c_int)]


class MeshInstanceInfo(C.Structure):
    _fields_ = [("matrix", MAT4), ("importance", C.c_float), ("shadow", C.c_int), ("ident", C.c_char_p),
                ("ident_size", C.c_size_t)]


class LightInfo(C.Structure):
    _fields_ = [("type", u32), ("position", VEC3), ("direction", VEC3), ("up_direction", VEC3), ("color_rgb", VEC3),
                ("range", C.c_float), ("power", C.c_float), ("light_radius", C.c_float), ("shadow_sample_count", C.c_int),
                ("spot_angle_out", C.c_float), ("spot_angle_in", C.c_float), ("spot_curve", C.c_float)]


class SampleInfo(C.Structure):
    _fields_ = [("id", u32), ("position", VEC3), ("normal", VEC3), ("out_color", VEC3)]


class SampleRequest(C.Structure):
    _fields_ = [("position", VEC3), ("normal", VEC3), ("tex0u", C.c_float), ("tex0v", C.c_float), ("tex1u", C.c_float),
                ("tex1v", C.c_float), ("part_id", u32), ("mesh_ident", C.c_char_p), ("mesh_ident_size", C.c_size_t),
                ("inst_ident", C.c_char_p), ("inst_ident_size", C.c_size_t), ("out_diffuse_color", VEC3),
                ("out_emissive_color", VEC3)]


class Config(C.Structure):
    pass


SIZE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(Config), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_float, C.c_float,
                      C.POINTER(u32))
SAMPLE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(Config), C.POINTER(SampleRequest))

Config._fields_ = [("userdata", C.c_void_p), ("size_fn", SIZE_FN), ("max_num_threads", C.c_int),
                   ("max_tree_memory", C.c_size_t), ("max_lightmap_size", u32), ("default_width", u32),
                   ("default_height", u32), ("global_size_factor", C.c_float), ("max_correct_dist", C.c_float),
                   ("max_correct_angle", C.c_float), ("clear_color", VEC3), ("ambient_color", VEC3),
                   ("bounce_count", C.c_int), ("sample_fn", SAMPLE_FN), ("ao_distance", C.c_float),
                   ("ao_multiplier", C.c_float), ("ao_falloff", C.c_float), ("ao_effect", C.c_float),
                   ("ao_divergence", C.c_float), ("ao_color_rgb", VEC3), ("ao_num_samples", C.c_int),
                   ("blur_size", C.c_float), ("ds2x", C.c_int), ("generate_normalmap_data", C.c_int)]


class WorkOutputInfo(C.Structure):
    _fields_ = [("lightmap_count", u32), ("sample_count", u32), ("samples", C.POINTER(SampleInfo))]


class WorkOutput(C.Structure):
    _fields_ = [("uid", u32), ("mesh_ident", C.c_char_p), ("mesh_ident_size", C.c_size_t), ("inst_ident", C.c_char_p),
                ("inst_ident_size", C.c_size_t), ("lightmap_rgb", C.POINTER(C.c_float)),
                ("normals_xyzf", C.POINTER(C.c_float)), ("width", u32), ("height", u32)]


class WorkStatus(C.Structure):
    _fields_ = [("completion", C.c_float), ("stage", C.c_char_p)]


class Stats(C.Structure):
    _fields_ = ([(n, C.c_double) for n in ("t_total", "t_prexform", "t_accel", "t_upload", "t_samples", "t_direct",
                                           "t_radiosity", "t_ao", "t_finalize", "t_readback")] +
                [(n, C.c_float) for n in ("gpu_ms_samples", "gpu_ms_direct", "gpu_ms_march", "gpu_ms_radiosity",
                                          "gpu_ms_ao", "gpu_ms_finalize", "gpu_ms_total", "gpu_ms_rad_pairs", "gpu_ms_rad_vis",
                                          "gpu_ms_span")] +
                [(n, C.c_uint64) for n in ("n_lumels_total", "n_lumels_local", "n_triangles", "n_bvh_nodes", "n_marches",
                                           "n_distance_queries", "n_ao_segments", "n_correction_rays", "n_rad_pairs",
                                           "n_rad_segments", "n_rad_links", "n_node_visits", "n_tri_tests",
                                           "n_ray_node_visits", "n_ray_tri_tests", "n_rad_tile_loads",
                                           "kernel_launches", "h2d_bytes", "d2h_bytes", "n_rad_batches", "n_shadow_rays", "n_ray_entry_tests")] +
                [("t_sample_fn", C.c_double)])

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class Lumels(C.Structure):
    _fields_ = [("count", u32), ("width", u32), ("height", u32), ("pos_xyz", C.POINTER(C.c_float)),
                ("nrm_xyz", C.POINTER(C.c_float)), ("loc", C.POINTER(u32)), ("radinfo_xyzw", C.POINTER(C.c_float)),
                ("rgb", C.POINTER(C.c_float))]


def test_bvh_entry(tris: np.ndarray, segs: np.ndarray, bundle_off: np.ndarray, leaf_max: int = 2) -> dict:
    """Host-only: entry sets (csrc/bvh_entry.h) of bundles of segments -- segs (n,6), bundle_off (nb+1) -- with the
    reachability self-check and the any-hit walk from the root vs from the entry set (4-wide node reads summed)."""
    tris = np.ascontiguousarray(tris, np.float32)
    segs = np.ascontiguousarray(segs, np.float32).reshape(-1, 6)
    off = np.ascontiguousarray(bundle_off, np.uint32)
    nb = len(off) - 1
    entries = np.zeros(max(nb, 1), np.uint32)
    vr, ve, et, mm = C.c_uint64(), C.c_uint64(), C.c_uint64(), u32()
    ok = lib().ltrx_test_bvh_entry(_fp(tris), len(tris), leaf_max, _fp(segs), off.ctypes.data, nb, entries.ctypes.data,
                                   C.byref(vr), C.byref(ve), C.byref(et), C.byref(mm))
    return dict(ok=bool(ok), entries=entries[:nb], visits_root=vr.value, visits_entry=ve.value, entry_tests=et.value,
                mismatches=mm.value)


def test_bvh_entry2(tris: np.ndarray, segs: np.ndarray, bundle_off: np.ndarray, leaf_max: int = 2, max_entries: int = 8, shaft: bool = True, batch: int = 0) -> dict:
    """Host-only: version-2 entry sets (csrc/bvh_entry.h: shaft-culled search, leaf entries) of bundles of segments whose first
    end points form one box and whose second end points another -- root walk vs entry walk on every segment, hit / miss AND
    number of triangles tested.  batch > 0: the packet form (per `batch` consecutive segments one walk lists the leaves in the
    batch's shaft; a ray tests only those leaf boxes): visits_entry = node reads of those walks, entry_tests = listed boxes tested."""
    tris = np.ascontiguousarray(tris, np.float32)
    segs = np.ascontiguousarray(segs, np.float32).reshape(-1, 6)
    off = np.ascontiguousarray(bundle_off, np.uint32)
    nb = len(off) - 1
    entries = np.zeros(max(nb, 1), np.uint32)
    stats = (C.c_uint64 * 4)()
    mm, td = u32(), u32()
    ok = lib().ltrx_test_bvh_entry2(_fp(tris), len(tris), leaf_max, _fp(segs), off.ctypes.data, nb, int(max_entries), int(bool(shaft)), int(batch),
                                    entries.ctypes.data, stats, C.byref(mm), C.byref(td))
    return dict(ok=bool(ok), entries=entries[:nb], visits_root=stats[0], visits_entry=stats[1], entry_tests=stats[2], tri_tests=stats[3],
                mismatches=mm.value, test_diffs=td.value)


def test_bvh_entry_cost(tris: np.ndarray, segs: np.ndarray, bundle_off: np.ndarray, leaf_max: int = 2):
    """Host-only: per-segment (4-wide node reads, triangle tests) of the any-hit walk from each bundle's entry set."""
    tris = np.ascontiguousarray(tris, np.float32)
    segs = np.ascontiguousarray(segs, np.float32).reshape(-1, 6)
    off = np.ascontiguousarray(bundle_off, np.uint32)
    nodes, tests = np.zeros(len(segs), np.uint32), np.zeros(len(segs), np.uint32)
    if not lib().ltrx_test_bvh_entry_cost(_fp(tris), len(tris), leaf_max, _fp(segs), off.ctypes.data, len(off) - 1, nodes.ctypes.data, tests.ctypes.data):
        raise RuntimeError("ltrx_test_bvh_entry_cost failed")
    return nodes, tests


def test_rad_cull(rowP, rowN, colP, colN):
    """Host-only: the pair sweep's culling tests (csrc/rad_cull.h) for one block of row lumels against one block of column
    lumels -> (block_ok, row_ok[nrows], pair_fast[nrows, ncols])."""
    rowP, rowN, colP, colN = (np.ascontiguousarray(a, np.float32).reshape(-1, 3) for a in (rowP, rowN, colP, colN))
    ok = C.c_int(0)
    row_ok = np.zeros(len(rowP), np.uint8)
    fast = np.zeros((len(rowP), len(colP)), np.uint8)
    if not lib().ltrx_test_rad_cull(_fp(rowP), _fp(rowN), len(rowP), _fp(colP), _fp(colN), len(colP), C.byref(ok), row_ok.ctypes.data, fast.ctypes.data):
        raise RuntimeError("ltrx_test_rad_cull: empty block")
    return bool(ok.value), row_ok.astype(bool), fast.astype(bool)


class Links(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("count", C.c_uint64), ("row_offset", C.POINTER(C.c_uint64)),
                ("other", C.POINTER(u32)), ("factor", C.POINTER(C.c_float))]


# every symbol include/lighter.h and include/lighter_b200.h declare (tests/test_abi.py checks the .so exports them)
LTR_SYMBOLS = ["ltr_DefaultSizeFunc", "ltr_CreateScene", "ltr_DestroyScene", "ltr_Start", "ltr_Abort", "ltr_GetStatus",
               "ltr_Sleep", "ltr_GetConfig", "ltr_SetConfig", "ltr_CreateMesh", "ltr_MeshAddPart", "ltr_MeshAddInstance",
               "ltr_LightAdd", "ltr_SampleAdd", "ltr_GetWorkOutputInfo", "ltr_GetWorkOutput", "ltr_NextPowerOfTwo"]
LTRX_SYMBOLS = ["ltrx_Version", "ltrx_SampleFnChecker", "ltrx_SetDevice", "ltrx_SetOutputRoot", "ltrx_NcclUniqueId", "ltrx_SetShard", "ltrx_ShardRange", "ltrx_GetStats",
                "ltrx_GetError", "ltrx_Prepare", "ltrx_BakeResident", "ltrx_Finish", "ltrx_OutputHash", "ltrx_SetDebug", "ltrx_GetLumels",
                "ltrx_GetLinks", "ltrx_GetShadowFactors", "ltrx_SetShadowMode", "ltrx_GetShadowMasks", "ltrx_ShadowSampleSegment",
                "ltrx_test_point_tri_distance", "ltrx_test_seg_tri",
                "ltrx_test_scene_queries", "ltrx_test_device_bvh", "ltrx_test_march", "ltrx_test_spiral_dirs", "ltrx_test_reftree", "ltrx_test_bvh", "ltrx_test_bvh_entry", "ltrx_test_bvh_entry2", "ltrx_test_bvh_entry_cost", "ltrx_test_host_prepare", "ltrx_test_rad_cull",
                "ltrx_test_rand_fill"]

_lib = None


def lib() -> C.CDLL:
    """Load liblighter_b200.so (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(lighter_b200 has no CPU or Python fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.ltr_CreateScene.restype = vp
    L.ltr_DestroyScene.argtypes = [vp]
    L.ltr_Start.argtypes = [vp]
    L.ltr_Abort.argtypes = [vp]
    L.ltr_GetStatus.argtypes = [vp, C.POINTER(WorkStatus)]
    L.ltr_Sleep.argtypes = [C.c_int]
    L.ltr_GetConfig.argtypes = [C.POINTER(Config), vp]
    L.ltr_SetConfig.argtypes = [vp, C.POINTER(Config)]
    L.ltr_CreateMesh.restype = vp
    L.ltr_CreateMesh.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.ltr_MeshAddPart.argtypes = [vp, C.POINTER(MeshPartInfo)]
    L.ltr_MeshAddInstance.argtypes = [vp, C.POINTER(MeshInstanceInfo)]
    L.ltr_LightAdd.argtypes = [vp, C.POINTER(LightInfo)]
    L.ltr_SampleAdd.argtypes = [vp, C.POINTER(SampleInfo)]
    L.ltr_GetWorkOutputInfo.argtypes = [vp, C.POINTER(WorkOutputInfo)]
    L.ltr_GetWorkOutput.argtypes = [vp, u32, C.POINTER(WorkOutput)]
    L.ltr_NextPowerOfTwo.restype = u32
    L.ltr_NextPowerOfTwo.argtypes = [u32]
    L.ltr_DefaultSizeFunc.argtypes = [C.POINTER(Config), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_float,
                                      C.c_float, C.POINTER(u32)]
    L.ltrx_Version.restype = C.c_char_p
    L.ltrx_SetDevice.argtypes = [vp, C.c_int]
    L.ltrx_NcclUniqueId.argtypes = [C.c_char_p]
    L.ltrx_SetShard.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.ltrx_SetOutputRoot.argtypes = [vp, C.c_int]
    L.ltrx_ShardRange.restype = None
    L.ltrx_ShardRange.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.ltrx_GetStats.argtypes = [vp, C.POINTER(Stats)]
    L.ltrx_GetError.restype = C.c_char_p
    L.ltrx_GetError.argtypes = [vp]
    L.ltrx_Prepare.argtypes = [vp]
    L.ltrx_test_host_prepare.argtypes = [vp, C.c_void_p]
    L.ltrx_BakeResident.argtypes = [vp, C.POINTER(C.c_float)]
    L.ltrx_Finish.argtypes = [vp]
    L.ltrx_OutputHash.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.ltrx_SetDebug.argtypes = [vp, C.c_int]
    L.ltrx_GetLumels.argtypes = [vp, u32, C.POINTER(Lumels)]
    L.ltrx_GetLinks.argtypes = [vp, C.POINTER(Links)]
    L.ltrx_GetShadowFactors.argtypes = [vp, u32, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_uint64)]
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.ltrx_SetShadowMode.argtypes = [vp, C.c_int]
    L.ltrx_GetShadowMasks.argtypes = [vp, u32, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]
    L.ltrx_ShadowSampleSegment.argtypes = [vp, u32, u32, fp, fp, fp, fp]
    L.ltrx_test_point_tri_distance.argtypes = [fp, fp, u32, fp]
    L.ltrx_test_seg_tri.argtypes = [fp, fp, fp, u32, fp]
    L.ltrx_test_scene_queries.argtypes = [fp, u32, fp, fp, u32, fp, ip, fp, ip]
    L.ltrx_test_march.argtypes = [fp, u32, fp, fp, fp, u32, fp, C.POINTER(u32)]
    L.ltrx_test_device_bvh.argtypes = [fp, u32, C.c_int, C.POINTER(u32), C.POINTER(u32), ip, fp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), ip]
    L.ltrx_test_spiral_dirs.argtypes = [fp, fp, u32, C.c_int, fp]
    L.ltrx_test_reftree.argtypes = [fp, u32, C.c_void_p, u32, C.c_void_p, u32, C.POINTER(u32), C.POINTER(u32)]
    L.ltrx_test_rand_fill.argtypes = [fp, C.c_uint64]
    L.ltrx_test_bvh.argtypes = [fp, u32, C.c_int, C.POINTER(u32), C.POINTER(u32), C.c_void_p, fp]
    L.ltrx_test_bvh_entry_cost.argtypes = [fp, u32, C.c_int, fp, C.c_void_p, u32, C.c_void_p, C.c_void_p]
    L.ltrx_test_rad_cull.argtypes = [fp, fp, u32, fp, fp, u32, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    L.ltrx_test_bvh_entry.argtypes = [fp, u32, C.c_int, fp, C.c_void_p, u32, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64), C.POINTER(u32)]
    L.ltrx_test_bvh_entry2.argtypes = [fp, u32, C.c_int, fp, C.c_void_p, u32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64),
                                       C.POINTER(u32), C.POINTER(u32)]
    _lib = L
    return L


_libc = C.CDLL(ctypes.util.find_library("c") or "libc.so.6")


def srand(seed: int = 1) -> None:
    """Reset the process's libc rand() stream.  The bake consumes rand() exactly as the reference does
    (one draw per light added, one per lumel in the AO pass); srand(1) is the state of a fresh process."""
    _libc.srand(C.c_uint(seed))


def libc_rand() -> int:
    return _libc.rand()


def test_rand_fill(n: int):
    """(values, fast) -- n randf() draws of the libc stream through the library's replay routine."""
    out = np.zeros(max(n, 1), np.float32)
    fast = lib().ltrx_test_rand_fill(_fp(out), n)
    return out[:n], bool(fast)


def shard_range(n: int, rank: int, world: int) -> tuple:
    b, e = C.c_uint64(), C.c_uint64()
    lib().ltrx_ShardRange(n, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# the material rule the reference's rad1 scenario installs (lighter_test.cpp:470-480)
@SAMPLE_FN
def _redwall_sample_fn(cfg, req):
    r = req.contents
    if r.position[0] <= -3 and r.normal[0] > np.float32(0.1):
        r.out_diffuse_color[0], r.out_diffuse_color[1], r.out_diffuse_color[2] = 0.5, 0.05, 0.02
    return 1


class BakeHandle:
    """A scene fed through the C API, kept alive so outputs (owned by the scene) stay valid."""

    def __init__(self, scene: _scenes.Scene, debug: bool = False, device: int | None = None, shard: tuple | None = None,
                 reset_rand: bool = True, shadow_mode: int = 0, output_root_only: bool = False):
        L = lib()
        self.L = L
        self.scene_desc = scene
        self._keep = []
        if reset_rand:
            srand(1)
        self.h = L.ltr_CreateScene()
        cfg = Config()
        L.ltr_GetConfig(C.byref(cfg), self.h)
        c = scene.cfg
        for k in ("max_lightmap_size", "default_width", "default_height", "global_size_factor", "max_correct_dist",
                  "max_correct_angle", "bounce_count", "ao_distance", "ao_multiplier", "ao_falloff", "ao_effect",
                  "ao_divergence", "ao_num_samples", "blur_size", "ds2x", "generate_normalmap_data"):
            setattr(cfg, k, c[k])
        cfg.clear_color = VEC3(*c["clear_color"])
        cfg.ambient_color = VEC3(*c["ambient_color"])
        cfg.ao_color_rgb = VEC3(*c["ao_color"])
        if c["sample_fn_kind"] == 1:
            cfg.sample_fn = _redwall_sample_fn
        elif c["sample_fn_kind"] == 2:          # the library's native example callback (no interpreter in the per-lumel loop)
            cfg.sample_fn = C.cast(L.ltrx_SampleFnChecker, SAMPLE_FN)
        if c["size_fn_kind"] == 1:
            forced = {i.ident.encode(): i.force_size for i in scene.instances if i.force_size[0]}

            @SIZE_FN
            def size_fn(cfgp, mid, midn, iid, iidn, area, imp, out):
                wh = forced.get(C.string_at(iid, iidn) if iid else b"")
                if wh is None:
                    return 0
                out[0], out[1] = wh
                return 1
            self._keep.append(size_fn)
            cfg.size_fn = size_fn
        L.ltr_SetConfig(self.h, C.byref(cfg))
        if device is not None:
            L.ltrx_SetDevice(self.h, device)
        if debug:
            L.ltrx_SetDebug(self.h, 1)
        if shadow_mode and not L.ltrx_SetShadowMode(self.h, shadow_mode):
            raise RuntimeError("ltrx_SetShadowMode rejected the mode")
        if shard is not None:
            rank, world, nccl_id = shard
            if not L.ltrx_SetShard(self.h, rank, world, nccl_id):
                raise RuntimeError("ltrx_SetShard rejected the shard")
            if output_root_only:
                L.ltrx_SetOutputRoot(self.h, 1)

        meshes = []
        for m in scene.meshes:
            ident = m.ident.encode()
            mh = L.ltr_CreateMesh(self.h, ident, len(ident))
            for p in m.parts:
                arrs = [np.ascontiguousarray(p.pos, np.float32), np.ascontiguousarray(p.nrm, np.float32),
                        np.ascontiguousarray(p.uv1, np.float32), np.ascontiguousarray(p.uv2, np.float32),
                        np.ascontiguousarray(p.idx, np.uint32)]
                pi = MeshPartInfo(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data,
                                  12, 12, 8, 8, arrs[4].ctypes.data, len(arrs[0]), len(arrs[4]), p.shadow)
                if not L.ltr_MeshAddPart(mh, C.byref(pi)):
                    raise RuntimeError("ltr_MeshAddPart failed")
            meshes.append(mh)
        for i in scene.instances:
            ii = MeshInstanceInfo()
            mat = np.ascontiguousarray(i.matrix, np.float32)
            C.memmove(ii.matrix, mat.ctypes.data, 64)
            ii.importance, ii.shadow = i.importance, i.shadow
            ident = i.ident.encode()
            ii.ident, ii.ident_size = ident, len(ident)
            L.ltr_MeshAddInstance(meshes[i.mesh], C.byref(ii))
        for lt in scene.lights:
            li = LightInfo(lt.type, VEC3(*lt.position), VEC3(*lt.direction), VEC3(*lt.up_direction), VEC3(*lt.color_rgb),
                           lt.range, lt.power, lt.light_radius, lt.shadow_sample_count, lt.spot_angle_out,
                           lt.spot_angle_in, lt.spot_curve)
            L.ltr_LightAdd(self.h, C.byref(li))
        for pid, pos, nrm in scene.probes:
            si = SampleInfo(pid, VEC3(*pos), VEC3(*nrm), VEC3(0, 0, 0))
            L.ltr_SampleAdd(self.h, C.byref(si))

    # -- the reference's DOWORK loop (lighter_test.cpp:167-180), polled at 0.2 ms ------------------
    def run(self, poll_s: float = 0.0002) -> float:
        st = WorkStatus()
        t0 = time.perf_counter()
        self.L.ltr_Start(self.h)
        self.stages = []
        while self.L.ltr_GetStatus(self.h, C.byref(st)):
            if not self.stages or self.stages[-1] != st.stage:
                self.stages.append(st.stage)
            time.sleep(poll_s)
        wall = time.perf_counter() - t0
        self._raise_on_error()
        return wall

    def _raise_on_error(self):
        err = self.L.ltrx_GetError(self.h)
        if err:
            raise RuntimeError("bake failed: " + err.decode())

    def host_prepare_fingerprints(self) -> list:
        """Host-only: run the host pre-pass (no device needed) and return FNV-1a fingerprints of the arrays it would upload."""
        out = (C.c_uint64 * 12)()
        if not self.L.ltrx_test_host_prepare(self.h, out):
            self._raise_on_error()
            raise RuntimeError("ltrx_test_host_prepare failed")
        return [int(x) for x in out]

    def prepare(self):
        if not self.L.ltrx_Prepare(self.h):
            self._raise_on_error()
            raise RuntimeError("ltrx_Prepare failed")

    def bake_resident(self) -> float:
        ms = C.c_float(0)
        if not self.L.ltrx_BakeResident(self.h, C.byref(ms)):
            self._raise_on_error()
            raise RuntimeError("ltrx_BakeResident failed")
        return ms.value

    def finish(self):
        if not self.L.ltrx_Finish(self.h):
            self._raise_on_error()
            raise RuntimeError("ltrx_Finish failed")

    def output_hash(self) -> str:
        """FNV-1a-64 (hex) of all lightmaps + probe colours of the last bake, computed by the library over its own output arrays."""
        h = C.c_uint64(0)
        if not self.L.ltrx_OutputHash(self.h, C.byref(h)):
            return ""
        return f"{h.value:016x}"

    def stats(self) -> dict:
        s = Stats()
        self.L.ltrx_GetStats(self.h, C.byref(s))
        return s.as_dict()

    def outputs(self) -> dict:
        info = WorkOutputInfo()
        self.L.ltr_GetWorkOutputInfo(self.h, C.byref(info))
        lms = []
        for i in range(info.lightmap_count):
            wo = WorkOutput()
            if not self.L.ltr_GetWorkOutput(self.h, i, C.byref(wo)):
                raise RuntimeError("ltr_GetWorkOutput failed")
            n = wo.width * wo.height
            rgb = np.ctypeslib.as_array(wo.lightmap_rgb, (n * 3,)).reshape(wo.height, wo.width, 3).copy() if n else np.zeros((0, 0, 3), np.float32)
            nrm = None
            if wo.normals_xyzf:
                nrm = np.ctypeslib.as_array(wo.normals_xyzf, (n * 4,)).reshape(wo.height, wo.width, 4).copy()
            lms.append(dict(uid=wo.uid, width=wo.width, height=wo.height, rgb=rgb, normals=nrm))
        probes = np.array([[info.samples[i].out_color[k] for k in range(3)] for i in range(info.sample_count)], np.float32).reshape(-1, 3)
        return dict(lightmaps=lms, probes=probes)

    def lumels(self, instance: int) -> dict:
        lm = Lumels()
        if not self.L.ltrx_GetLumels(self.h, instance, C.byref(lm)):
            raise RuntimeError("ltrx_GetLumels: no stage dump (bake with debug=True)")
        n = lm.count

        def arr(p, k, dt=np.float32):
            if n == 0:
                return np.zeros((0, k) if k > 1 else (0,), dt)
            a = np.ctypeslib.as_array(p, (n * k,)).copy()
            return a.reshape(n, k) if k > 1 else a
        return dict(n=n, width=lm.width, height=lm.height, pos=arr(lm.pos_xyz, 3), nrm=arr(lm.nrm_xyz, 3),
                    loc=arr(lm.loc, 1, np.uint32), radinfo=arr(lm.radinfo_xyzw, 4), rgb=arr(lm.rgb, 3))

    def links(self) -> dict:
        lk = Links()
        if not self.L.ltrx_GetLinks(self.h, C.byref(lk)):
            return dict(rows=0, row_offset=np.zeros(1, np.uint64), other=np.zeros(0, np.uint32), factor=np.zeros(0, np.float32))
        ro = np.ctypeslib.as_array(lk.row_offset, (lk.rows + 1,)).copy()
        other = np.ctypeslib.as_array(lk.other, (lk.count,)).copy() if lk.count else np.zeros(0, np.uint32)
        fac = np.ctypeslib.as_array(lk.factor, (lk.count,)).copy() if lk.count else np.zeros(0, np.float32)
        return dict(rows=lk.rows, row_offset=ro, other=other, factor=fac)

    def shadow_factors(self, light: int) -> np.ndarray:
        p, n = C.POINTER(C.c_float)(), C.c_uint64()
        if not self.L.ltrx_GetShadowFactors(self.h, light, C.byref(p), C.byref(n)):
            raise RuntimeError("no shadow factors kept (bake with debug=True)")
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def shadow_masks(self, light: int) -> np.ndarray:
        """Sampled-shadow mode, debug bakes: per local lumel, bit s set = sample s of `light` is blocked."""
        p, n = C.POINTER(C.c_uint64)(), C.c_uint64()
        if not self.L.ltrx_GetShadowMasks(self.h, light, C.byref(p), C.byref(n)):
            raise RuntimeError("no shadow masks kept (bake with debug=True, shadow_mode=1)")
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def shadow_segment(self, light: int, sample: int, pos, nrm):
        """Host evaluation of the kernel's own segment function for one lumel (bit-identical end points)."""
        pos = np.ascontiguousarray(pos, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
        a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        if not self.L.ltrx_ShadowSampleSegment(self.h, light, sample, _fp(pos), _fp(nrm), _fp(a), _fp(b)):
            raise RuntimeError("ltrx_ShadowSampleSegment: bad light/sample or scene not baked")
        return a, b

    def close(self):
        if self.h:
            self.L.ltr_DestroyScene(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def bake(scene: _scenes.Scene, debug: bool = False, **kw) -> dict:
    """One full bake through the public API; returns lightmaps, probe colours, wall time, stats."""
    with BakeHandle(scene, debug=debug, **kw) as b:
        wall = b.run()
        out = b.outputs()
        out["wall_s"] = wall
        out["stats"] = b.stats()
        out["stages"] = [s.decode() for s in b.stages if s]
        if debug:
            out["instances"] = [b.lumels(i) for i in range(len(scene.instances) + 1)]
            out["links"] = b.links()
        return out


# ---- kernel-level entry points -----------------------------------------------------------------
def test_point_tri_distance(pts: np.ndarray, tris: np.ndarray) -> np.ndarray:
    pts = np.ascontiguousarray(pts, np.float32); tris = np.ascontiguousarray(tris, np.float32)
    out = np.zeros(len(pts), np.float32)
    if not lib().ltrx_test_point_tri_distance(_fp(pts), _fp(tris), len(pts), _fp(out)):
        raise RuntimeError("ltrx_test_point_tri_distance failed (no CUDA device?)")
    return out


def test_seg_tri(a: np.ndarray, b: np.ndarray, tris: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32); tris = np.ascontiguousarray(tris, np.float32)
    out = np.zeros(len(a), np.float32)
    if not lib().ltrx_test_seg_tri(_fp(a), _fp(b), _fp(tris), len(a), _fp(out)):
        raise RuntimeError("ltrx_test_seg_tri failed (no CUDA device?)")
    return out


def test_scene_queries(tris: np.ndarray, a: np.ndarray, b: np.ndarray) -> dict:
    tris = np.ascontiguousarray(tris, np.float32); a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    n = len(a)
    dist, closest = np.zeros(n, np.float32), np.zeros(n, np.float32)
    anyhit, ctri = np.zeros(n, np.int32), np.zeros(n, np.int32)
    ip = C.POINTER(C.c_int)
    if not lib().ltrx_test_scene_queries(_fp(tris), len(tris), _fp(a), _fp(b), n, _fp(dist), anyhit.ctypes.data_as(ip),
                                         _fp(closest), ctri.ctypes.data_as(ip)):
        raise RuntimeError("ltrx_test_scene_queries failed (no CUDA device?)")
    return dict(dist=dist, anyhit=anyhit, closest=closest, closest_tri=ctri)


def test_device_bvh(tris: np.ndarray, leaf_max: int = 2) -> dict:
    """The scene BVH built on the device against the host builder on the same triangles (csrc/gpu_bvh.cu vs csrc/bvh.cpp)."""
    tris = np.ascontiguousarray(tris, np.float32)
    nn, nn4, mn, mn4, ml = u32(), u32(), u32(), u32(), u32()
    h, hh, ms = C.c_int(), C.c_int(), C.c_float()
    if not lib().ltrx_test_device_bvh(_fp(tris), len(tris), leaf_max, C.byref(nn), C.byref(nn4), C.byref(h), C.byref(ms), C.byref(mn), C.byref(mn4),
                                      C.byref(ml), C.byref(hh)):
        raise RuntimeError("ltrx_test_device_bvh failed (no CUDA device?)")
    return dict(n_nodes=nn.value, n_nodes4=nn4.value, height=h.value, host_height=hh.value, build_ms=ms.value,
                mismatch_nodes=mn.value, mismatch_nodes4=mn4.value, mismatch_leaves=ml.value)


def test_march(tris: np.ndarray, frm: np.ndarray, to: np.ndarray, k: np.ndarray) -> tuple:
    tris = np.ascontiguousarray(tris, np.float32); frm = np.ascontiguousarray(frm, np.float32)
    to = np.ascontiguousarray(to, np.float32); k = np.ascontiguousarray(k, np.float32)
    out, steps = np.zeros(len(frm), np.float32), np.zeros(len(frm), np.uint32)
    if not lib().ltrx_test_march(_fp(tris), len(tris), _fp(frm), _fp(to), _fp(k), len(frm), _fp(out), steps.ctypes.data_as(C.POINTER(u32))):
        raise RuntimeError("ltrx_test_march failed (no CUDA device?)")
    return out, steps


def test_spiral_dirs(nrm: np.ndarray, randoff: np.ndarray, samples: int) -> np.ndarray:
    nrm = np.ascontiguousarray(nrm, np.float32); randoff = np.ascontiguousarray(randoff, np.float32)
    out = np.zeros((len(nrm), samples, 3), np.float32)
    if not lib().ltrx_test_spiral_dirs(_fp(nrm), _fp(randoff), len(nrm), samples, _fp(out)):
        raise RuntimeError("ltrx_test_spiral_dirs failed (no CUDA device?)")
    return out


def test_reftree(tris: np.ndarray) -> tuple:
    """Host-only: build the reference-order box tree over triangle boxes -> (nodes[n,8] as u32 words, items)."""
    tris = np.ascontiguousarray(tris, np.float32)
    cap_n, cap_i = 2 * len(tris) + 8, 3 * len(tris) + 8
    nodes = np.zeros((cap_n, 8), np.uint32)
    items = np.zeros(cap_i, np.int32)
    nn, ni = u32(), u32()
    if not lib().ltrx_test_reftree(_fp(tris), len(tris), nodes.ctypes.data, cap_n, items.ctypes.data, cap_i, C.byref(nn), C.byref(ni)):
        raise RuntimeError("ltrx_test_reftree: capacity")
    return nodes[:nn.value], items[:ni.value]


def test_bvh(tris: np.ndarray, leaf_max: int = 4) -> dict:
    """Host-only: build the flat scene BVH and run its structural self-check."""
    tris = np.ascontiguousarray(tris, np.float32)
    nn, depth = u32(), u32()
    order = np.zeros(max(len(tris), 1), np.uint32)
    bounds = np.zeros(6, np.float32)
    ok = lib().ltrx_test_bvh(_fp(tris), len(tris), leaf_max, C.byref(nn), C.byref(depth), order.ctypes.data, _fp(bounds))
    return dict(ok=bool(ok), n_nodes=nn.value, depth=depth.value, order=order[:len(tris)], bounds=bounds)
