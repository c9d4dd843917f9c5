"""TEST INFRASTRUCTURE: ctypes bindings for the CPU oracle (oracle/liboracle.so, the plain-C
restatement) and, when built, for the reference's own primitives (oracle/_ref/libref_prims.so).
Imported only by tests/, tools/make_golden.py, __graft_entry__.smoke() and bench.py's CPU leg."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_PRIMS_SO = os.path.join(_HERE, "_ref", "libref_prims.so")

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
up = C.POINTER(C.c_uint32)


def _f(a):
    return a.ctypes.data_as(fp)


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dt)


class Oracle:
    """The plain-C restatement."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle oracle`")
        L = self.L = C.CDLL(ORACLE_SO)
        L.o_point_tri_distance.restype = C.c_float
        L.o_point_tri_distance.argtypes = [fp, fp]
        L.o_seg_tri.restype = C.c_float
        L.o_seg_tri.argtypes = [fp, fp, fp]
        L.o_scene_distance.restype = C.c_float
        L.o_scene_distance.argtypes = [fp, C.c_int, fp]
        L.o_visibility_test.argtypes = [fp, C.c_int, fp, fp]
        L.o_distance_test.restype = C.c_float
        L.o_distance_test.argtypes = [fp, C.c_int, fp, fp, ip]
        L.o_anyhit_raw.argtypes = [fp, C.c_int, fp, fp]
        L.o_closest_raw.restype = C.c_float
        L.o_closest_raw.argtypes = [fp, C.c_int, fp, fp, ip]
        L.o_march.restype = C.c_float
        L.o_march.argtypes = [fp, C.c_int, fp, fp, C.c_float, up]
        L.o_spiral_dir.argtypes = [fp, C.c_float, C.c_int, C.c_int, fp]
        L.o_direct_lumel.argtypes = [fp, C.c_int, fp, fp, fp, fp, fp]
        L.o_ao_lumel.argtypes = [fp, C.c_int, fp, fp, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, fp, fp]
        L.o_rad_links.restype = C.c_long
        L.o_rad_links.argtypes = [fp, C.c_int, fp, fp, C.c_int, C.c_long, up, up, fp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.o_rad_row.restype = C.c_long
        L.o_rad_row.argtypes = [fp, C.c_int, fp, fp, fp, fp, C.c_int, C.c_int, C.POINTER(C.c_uint8), fp, C.POINTER(C.c_long)]
        L.o_rad_bounce.argtypes = [up, up, fp, C.c_long, C.c_int, fp, fp, fp, C.c_int, fp]
        L.o_scatter_dilate.argtypes = [fp, up, C.c_int, C.c_int, C.c_int, fp]
        L.o_gauss_kernel.argtypes = [fp, C.c_int, C.c_float]
        L.o_blur.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
        L.o_downsample2x.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int]
        L.o_raster_tri.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_float, fp, fp, fp, fp]

    # -- primitives over arrays -------------------------------------------------------------------
    def point_tri_distance(self, pts, tris):
        pts, tris = _c(pts), _c(tris)
        return np.array([self.L.o_point_tri_distance(_f(pts[i]), _f(tris[i])) for i in range(len(pts))], np.float32)

    def seg_tri(self, a, b, tris):
        a, b, tris = _c(a), _c(b), _c(tris)
        return np.array([self.L.o_seg_tri(_f(a[i]), _f(b[i]), _f(tris[i])) for i in range(len(a))], np.float32)

    def scene_distance(self, tris, pts):
        tris, pts = _c(tris), _c(pts)
        return np.array([self.L.o_scene_distance(_f(tris), len(tris), _f(pts[i])) for i in range(len(pts))], np.float32)

    def visibility_test(self, tris, a, b):
        """ltr_Scene::VisibilityTest (lighter.cpp:138-147): 1 = the segment, pulled in 0.001 at both ends, is BLOCKED."""
        tris, a, b = _c(tris), _c(a), _c(b)
        return np.array([self.L.o_visibility_test(_f(tris), len(tris), _f(a[i]), _f(b[i])) for i in range(len(a))], np.int32)

    def anyhit_raw(self, tris, a, b):
        tris, a, b = _c(tris), _c(a), _c(b)
        return np.array([self.L.o_anyhit_raw(_f(tris), len(tris), _f(a[i]), _f(b[i])) for i in range(len(a))], np.int32)

    def closest_raw(self, tris, a, b):
        tris, a, b = _c(tris), _c(a), _c(b)
        out, tid = np.zeros(len(a), np.float32), np.zeros(len(a), np.int32)
        t = C.c_int()
        for i in range(len(a)):
            out[i] = self.L.o_closest_raw(_f(tris), len(tris), _f(a[i]), _f(b[i]), C.byref(t))
            tid[i] = t.value
        return out, tid

    def march(self, tris, frm, to, k):
        tris, frm, to = _c(tris), _c(frm), _c(to)
        out, steps = np.zeros(len(frm), np.float32), np.zeros(len(frm), np.uint32)
        s = C.c_uint32()
        for i in range(len(frm)):
            out[i] = self.L.o_march(_f(tris), len(tris), _f(frm[i]), _f(to[i]), float(k[i]), C.byref(s))
            steps[i] = s.value
        return out, steps

    def spiral_dirs(self, nrm, randoff, samples):
        nrm = _c(nrm)
        out = np.zeros((len(nrm), samples, 3), np.float32)
        for i in range(len(nrm)):
            for s in range(samples):
                self.L.o_spiral_dir(_f(nrm[i]), float(randoff[i]), s, samples, _f(out[i, s]))
        return out

    # -- stages ------------------------------------------------------------------------------------
    @staticmethod
    def pack_light(lt) -> np.ndarray:
        d = np.array(lt.direction, np.float32)
        l2 = np.float32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        if l2 != 0:
            d = d * (np.float32(1.0) / np.sqrt(l2, dtype=np.float32))       # Normalized(), lighter.cpp:1288
        return np.array([lt.type, *lt.position, *d, *lt.color_rgb, lt.range, lt.power, lt.light_radius, lt.spot_angle_out,
                         lt.spot_angle_in, lt.spot_curve], np.float32)

    def direct_light(self, tris, lights, pos, nrm, ambient=(0.0, 0.0, 0.0)):
        """Per-lumel direct light, lights accumulated in order (all instances lit by all lights)."""
        tris, pos, nrm = _c(tris), _c(pos), _c(nrm)
        rgb = np.tile(np.array(ambient, np.float32), (len(pos), 1))
        packed = [self.pack_light(lt) for lt in lights]
        fv = C.c_float()
        for i in range(len(pos)):
            for pl in packed:
                self.L.o_direct_lumel(_f(tris), len(tris), _f(pl), _f(pos[i]), _f(nrm[i]), _f(rgb[i]), C.byref(fv))
        return rgb

    def ambient_occlusion(self, tris, pos, nrm, randoff, rgb, cfg):
        tris, pos, nrm = _c(tris), _c(pos), _c(nrm)
        rgb = _c(rgb).copy()
        aoc = np.array(cfg["ao_color"], np.float32)
        for i in range(len(pos)):
            self.L.o_ao_lumel(_f(tris), len(tris), _f(pos[i]), _f(nrm[i]), float(randoff[i]), int(cfg["ao_num_samples"]),
                              float(cfg["ao_distance"]), float(cfg["ao_multiplier"]), float(cfg["ao_falloff"]), float(cfg["ao_effect"]),
                              _f(aoc), _f(rgb[i]))
        return rgb

    def rad_links(self, tris, pos, nrm, cap=1 << 22):
        tris, pos, nrm = _c(tris), _c(pos), _c(nrm)
        li, lj, lf = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.float32)
        pt, sg = C.c_long(), C.c_long()
        n = self.L.o_rad_links(_f(tris), len(tris), _f(pos), _f(nrm), len(pos), cap, li.ctypes.data_as(up), lj.ctypes.data_as(up), _f(lf),
                               C.byref(pt), C.byref(sg))
        assert n <= cap
        return li[:n], lj[:n], lf[:n], pt.value, sg.value

    def rad_row(self, tris, self_pos, self_nrm, pos, nrm, rank_of_self):
        """One row of the link double loop (lighter.cpp:728-760): lumel `self` against the candidates pos/nrm (ascending global
        order, self excluded; rank_of_self = how many of them have a lower global index).  Returns (linked mask, factors, segments)."""
        tris, pos, nrm = _c(tris), _c(pos), _c(nrm)
        sp, sn = _c(self_pos), _c(self_nrm)
        linked, fac, sg = np.zeros(len(pos), np.uint8), np.zeros(len(pos), np.float32), C.c_long()
        self.L.o_rad_row(_f(tris), len(tris), _f(sp), _f(sn), _f(pos), _f(nrm), len(pos), int(rank_of_self),
                         linked.ctypes.data_as(C.POINTER(C.c_uint8)), _f(fac), C.byref(sg))
        return linked.astype(bool), fac, sg.value

    def rad_bounce(self, li, lj, lf, diffuse, emit, area, bounces):
        li, lj, lf = _c(li, np.uint32), _c(lj, np.uint32), _c(lf)
        diffuse, emit, area = _c(diffuse), _c(emit), _c(area)
        total = np.zeros_like(emit)
        self.L.o_rad_bounce(li.ctypes.data_as(up), lj.ctypes.data_as(up), _f(lf), len(li), len(emit), _f(diffuse), _f(emit), _f(area), bounces, _f(total))
        return total

    def finalize(self, rgb, loc, w, h, blur_size=0.0, ds2x=0):
        rgb, loc = _c(rgb), _c(loc, np.uint32)
        img = np.zeros((h, w, 3), np.float32)
        self.L.o_scatter_dilate(_f(rgb), loc.ctypes.data_as(up), len(rgb), w, h, _f(img))
        if blur_size:
            ext = int(np.ceil(blur_size))
            k = np.zeros(2 * ext + 1, np.float32)
            self.L.o_gauss_kernel(_f(k), ext, float(blur_size))
            self.L.o_blur(_f(img), w, h, ext, _f(k))
        if ds2x:
            dw, dh = max(w // 2, 1), max(h // 2, 1)
            out = np.zeros((dh, dw, 3), np.float32)
            self.L.o_downsample2x(_f(out), dw, dh, _f(img), w, h)
            img = out
        return img

    def raster_tri(self, w, h, margin, p, va, vb, vc, imgs=None):
        if imgs is None:
            imgs = (np.zeros((h, w, 3), np.float32), np.zeros((h, w, 3), np.float32), np.zeros((h, w, 4), np.float32))
        p, va, vb, vc = _c(p), _c(va), _c(vb), _c(vc)
        self.L.o_raster_tri(_f(imgs[0]), _f(imgs[1]), _f(imgs[2]), w, h, float(margin), _f(p), _f(va), _f(vb), _f(vc))
        return imgs


class RefPrims:
    """The reference's OWN primitives (compiled from /root/reference into oracle/_ref)."""

    def __init__(self):
        if not os.path.exists(REF_PRIMS_SO):
            raise RuntimeError("oracle/_ref/libref_prims.so missing: run `make -C oracle ref` where /root/reference exists")
        L = self.L = C.CDLL(REF_PRIMS_SO)
        L.refp_point_tri_distance.restype = C.c_float
        L.refp_point_tri_distance.argtypes = [fp, fp]
        L.refp_point_proj_on_tri.argtypes = [fp, fp]
        L.refp_seg_tri.restype = C.c_float
        L.refp_seg_tri.argtypes = [fp, fp, fp]
        L.refp_ray_aabb.argtypes = [fp, fp, fp, fp]
        L.refp_spiral_dir.argtypes = [fp, C.c_float, C.c_int, C.c_int, fp]
        L.refp_triangle_area3.restype = C.c_float
        L.refp_triangle_area3.argtypes = [fp]
        L.refp_sample_area.restype = C.c_float
        L.refp_sample_area.argtypes = [fp, fp]
        L.refp_transform.argtypes = [fp, fp, fp, C.c_int, fp, fp]
        L.refp_raster_tri.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_float, fp, fp, fp, fp]
        L.refp_gauss_kernel.argtypes = [fp, C.c_int, C.c_float]
        L.refp_convolve_transpose.argtypes = [fp, fp, C.c_uint, C.c_uint, C.c_int, fp]
        L.refp_downsample2x.argtypes = [fp, C.c_uint, C.c_uint, fp, C.c_uint, C.c_uint]
        L.refp_tritree_create.restype = C.c_void_p
        L.refp_tritree_create.argtypes = [fp, C.c_int]
        L.refp_tritree_destroy.argtypes = [C.c_void_p]
        L.refp_tritree_tri_count.argtypes = [C.c_void_p]
        L.refp_tritree_distance.restype = C.c_float
        L.refp_tritree_distance.argtypes = [C.c_void_p, fp, C.c_float]
        L.refp_tritree_anyhit.argtypes = [C.c_void_p, fp, fp]
        L.refp_tritree_closest.restype = C.c_float
        L.refp_tritree_closest.argtypes = [C.c_void_p, fp, fp, ip]
        L.refp_tritree_offset.argtypes = [C.c_void_p, fp, fp, C.c_float]
        L.refp_tritree_node_count.argtypes = [C.c_void_p]
        L.refp_tritree_item_count.argtypes = [C.c_void_p]
        L.refp_tritree_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]

    def point_tri_distance(self, pts, tris):
        pts, tris = _c(pts), _c(tris)
        return np.array([self.L.refp_point_tri_distance(_f(pts[i]), _f(tris[i])) for i in range(len(pts))], np.float32)

    def seg_tri(self, a, b, tris):
        a, b, tris = _c(a), _c(b), _c(tris)
        return np.array([self.L.refp_seg_tri(_f(a[i]), _f(b[i]), _f(tris[i])) for i in range(len(a))], np.float32)

    def spiral_dirs(self, nrm, randoff, samples):
        nrm = _c(nrm)
        out = np.zeros((len(nrm), samples, 3), np.float32)
        for i in range(len(nrm)):
            for s in range(samples):
                self.L.refp_spiral_dir(_f(nrm[i]), float(randoff[i]), s, samples, _f(out[i, s]))
        return out

    def tree_queries(self, tris, a, b):
        """TriTree distance at a (cap 2.0), any-hit and closest-hit on the raw segment a-b."""
        tris, a, b = _c(tris), _c(a), _c(b)
        h = self.L.refp_tritree_create(_f(tris), len(tris))
        n = len(a)
        dist, closest = np.zeros(n, np.float32), np.zeros(n, np.float32)
        anyhit, tid = np.zeros(n, np.int32), np.zeros(n, np.int32)
        t = C.c_int()
        for i in range(n):
            dist[i] = self.L.refp_tritree_distance(h, _f(a[i]), 2.0)
            anyhit[i] = self.L.refp_tritree_anyhit(h, _f(a[i]), _f(b[i]))
            closest[i] = self.L.refp_tritree_closest(h, _f(a[i]), _f(b[i]), C.byref(t))
            tid[i] = t.value
        self.L.refp_tritree_destroy(h)
        return dict(dist=dist, anyhit=anyhit, closest=closest, closest_tri=tid)

    def raster_tri(self, w, h, margin, p, va, vb, vc, imgs=None):
        if imgs is None:
            imgs = (np.zeros((h, w, 3), np.float32), np.zeros((h, w, 3), np.float32), np.zeros((h, w, 4), np.float32))
        p, va, vb, vc = _c(p), _c(va), _c(vb), _c(vc)
        self.L.refp_raster_tri(_f(imgs[0]), _f(imgs[1]), _f(imgs[2]), w, h, float(margin), _f(p), _f(va), _f(vb), _f(vc))
        return imgs

    def blur(self, img, blur_size):
        h, w, _ = img.shape
        ext = int(np.ceil(blur_size))
        k = np.zeros(2 * ext + 1, np.float32)
        self.L.refp_gauss_kernel(_f(k), ext, float(blur_size))
        src = _c(img).copy()
        tmp = np.zeros((w, h, 3), np.float32)
        self.L.refp_convolve_transpose(_f(src), _f(tmp), w, h, ext, _f(k))
        out = np.zeros((h, w, 3), np.float32)
        self.L.refp_convolve_transpose(_f(tmp), _f(out), h, w, ext, _f(k))
        return out

    def downsample2x(self, img):
        h, w, _ = img.shape
        dw, dh = max(w // 2, 1), max(h // 2, 1)
        src = _c(img).copy()
        out = np.zeros((dh, dw, 3), np.float32)
        self.L.refp_downsample2x(_f(out), dw, dh, _f(src), w, h)
        return out
