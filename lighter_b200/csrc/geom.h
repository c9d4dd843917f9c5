/*
 * geom.h -- the geometric primitives of the bake path, host+device.
 *
 * Each function is a fresh restatement of the arithmetic the reference performs (cited per
 * function) with the operand order preserved so results are bit-identical; the data flow around
 * them (prepared triangles, flat BVH, one thread per query) is new.
 */
#pragma once
#include "vmath.h"

/* The branch-free slab test of the any-hit walks (gpu_internal.cuh) and an axis the segment does not move along (d == 0
 * exactly: lumels on the shared edge of two tiles, wall lumels).  Such an axis constrains nothing as long as the origin lies
 * inside the box's range on it, FACES INCLUDED (the reference skips the axis altogether: RayAABBTest, lighter_math.cpp:618-650).
 * A huge finite inverse does that without a branch: (lo - o) * 1e30 and (hi - o) * 1e30 keep their signs, so the axis
 * contributes (-huge, +huge) inside and an empty interval outside -- except ON a face, where the product is 0.  On the lower
 * face that is harmless (the interval starts at 0 anyway); on the upper face it ends the interval at 0 and the box is lost.
 * So the upper faces are tested against lb_slab_origin_hi: the origin itself on an axis the segment moves along, one ulp
 * below it on an axis it does not -- hi - o' is then a positive ulp on the face and the axis drops out.  Per ray, outside
 * the loop; the loop's instruction count is unchanged.  Round 1 did not do this and lost the hits of segments lying in a
 * tile's boundary plane (16 of 22 561 links on the config-4 sibling). */
LB_HD float lb_slab_inv(float d) { return d != 0 ? 1.0f / d : 1e30f; }    /* nvcc already compiles this to the reciprocal path: __frcp_rn gives the same SASS length */
LB_HD float lb_slab_origin_hi(float o, float d) { return d != 0 ? o : nextafterf(o, -INFINITY); }

/* ------------------------------------------------------------------------------------------
 * Point / triangle distance  (ref: lighter_math.cpp:875-913)
 *
 * Case analysis on a triangle t0,t1,t2: vertex regions first (point projects behind an edge start
 * and beyond the previous edge end), then edge slabs (outside the edge in the plane, projection
 * inside the edge span), else the plane distance.
 *
 * PreparedTri holds every per-triangle quantity of that analysis that does not depend on the query
 * point, computed once with the reference's exact expressions (prepare_tri below), so the per-query
 * cost drops from 7 normalisations + 4 cross products to a handful of dot products while the result
 * stays bit-identical.  160 bytes = 10 float4, 16B-aligned.
 * ------------------------------------------------------------------------------------------ */
struct PreparedTri {
    V3 t0;  float t0min;     /* dot(tan0,t0) */
    V3 t1;  float t1min;     /* dot(tan1,t1) */
    V3 t2;  float t2min;     /* dot(tan2,t2) */
    V3 tan0; float t0max;    /* dot(tan0,t1) */
    V3 tan1; float t1max;    /* dot(tan1,t2) */
    V3 tan2; float t2max;    /* dot(tan2,t0) */
    V3 nrm;  float dn;       /* dot(nrm,t0) */
    V3 en0;  float de0;      /* dot(en0,t0) */
    V3 en1;  float de1;      /* dot(en1,t1) */
    V3 en2;  float de2;      /* dot(en2,t2) */
};

LB_HD void prepare_tri(V3 t0, V3 t1, V3 t2, PreparedTri &o)
{
    V3 nrm = norm3(cross3(t1 - t0, t2 - t0));
    V3 tan0 = norm3(t1 - t0), tan1 = norm3(t2 - t1), tan2 = norm3(t0 - t2);
    o.t0 = t0; o.t1 = t1; o.t2 = t2;
    o.tan0 = tan0; o.tan1 = tan1; o.tan2 = tan2;
    o.t0min = dot3(tan0, t0); o.t0max = dot3(tan0, t1);
    o.t1min = dot3(tan1, t1); o.t1max = dot3(tan1, t2);
    o.t2min = dot3(tan2, t2); o.t2max = dot3(tan2, t0);
    o.nrm = nrm; o.dn = dot3(nrm, t0);
    o.en0 = norm3(cross3(t1 - t0, nrm)); o.de0 = dot3(o.en0, t0);
    o.en1 = norm3(cross3(t2 - t1, nrm)); o.de1 = dot3(o.en1, t1);
    o.en2 = norm3(cross3(t0 - t2, nrm)); o.de2 = dot3(o.en2, t2);
}

LB_HD float point_tri_distance_prepared(V3 pt, const PreparedTri &T)
{
    float t0p = dot3(T.tan0, pt), t1p = dot3(T.tan1, pt), t2p = dot3(T.tan2, pt);
    if (T.t0min >= t0p && T.t2max <= t2p) return len3(pt - T.t0);
    if (T.t1min >= t1p && T.t0max <= t0p) return len3(pt - T.t1);
    if (T.t2min >= t2p && T.t1max <= t1p) return len3(pt - T.t2);
    float pd = fabsf(dot3(T.nrm, pt) - T.dn);
    float ptd0 = dot3(T.en0, pt) - T.de0;
    float ptd1 = dot3(T.en1, pt) - T.de1;
    float ptd2 = dot3(T.en2, pt) - T.de2;
    if (ptd0 >= 0 && t0p >= T.t0min && t0p <= T.t0max) return sqrtf(pd * pd + ptd0 * ptd0);
    if (ptd1 >= 0 && t1p >= T.t1min && t1p <= T.t1max) return sqrtf(pd * pd + ptd1 * ptd1);
    if (ptd2 >= 0 && t2p >= T.t2min && t2p <= T.t2max) return sqrtf(pd * pd + ptd2 * ptd2);
    return pd;
}

LB_HD float point_tri_distance(V3 pt, V3 t0, V3 t1, V3 t2)
{
    PreparedTri T;
    prepare_tri(t0, t1, t2, T);
    return point_tri_distance_prepared(pt, T);
}

/* Does the point project onto the triangle, with SMALL_FLOAT slack (ref: lighter_math.cpp:915-952) */
LB_HD bool point_proj_on_tri(V3 pt, const PreparedTri &T)
{
    const float e = LB_SMALL;
    float t0p = dot3(T.tan0, pt), t1p = dot3(T.tan1, pt), t2p = dot3(T.tan2, pt);
    if (T.t0min - e > t0p && T.t2max + e < t2p) return false;
    if (T.t1min - e > t1p && T.t0max + e < t0p) return false;
    if (T.t2min - e > t2p && T.t1max + e < t1p) return false;
    float ptd0 = dot3(T.en0, pt) - T.de0;
    float ptd1 = dot3(T.en1, pt) - T.de1;
    float ptd2 = dot3(T.en2, pt) - T.de2;
    if (ptd0 > e && t0p - e > T.t0min && t0p + e < T.t0max) return false;
    if (ptd1 > e && t1p - e > T.t1min && t1p + e < T.t1max) return false;
    if (ptd2 > e && t2p - e > T.t2min && t2p + e < T.t2max) return false;
    return true;
}

/* ------------------------------------------------------------------------------------------
 * Segment / triangle intersection parameter (ref: lighter_math.cpp:315-354).
 * Returns r in [0,1] for a hit, 2.0f for a miss.  Two-sided.  RayTri caches the per-triangle terms.
 * 64 bytes = 4 float4.
 * ------------------------------------------------------------------------------------------ */
struct RayTri {
    V3 p1; float uu;
    V3 u;  float uv;
    V3 v;  float vv;
    V3 n;  float D;       /* n = cross(u,v) un-normalised; D = uv*uv - uu*vv */
};

LB_HD void prepare_raytri(V3 p1, V3 p2, V3 p3, RayTri &o)
{
    o.p1 = p1; o.u = p2 - p1; o.v = p3 - p1;
    o.n = cross3(o.u, o.v);
    o.uu = dot3(o.u, o.u); o.uv = dot3(o.u, o.v); o.vv = dot3(o.v, o.v);
    o.D = o.uv * o.uv - o.uu * o.vv;
}

#define LB_NO_HIT 2.0f

/* ------------------------------------------------------------------------------------------
 * Sampled soft shadows (EXTENSION mode; the reference only has a disabled directional-light
 * sketch of it, lighter.cpp:566-584, and "TODO" for point/spot lights, lighter.cpp:1311-1314).
 *
 * Every light is an area source sampled on a golden-angle disk: sample s of n sits at radius
 * light_radius*sqrt((s+.5)/n), angle (s+randoff)*137.508 deg (randoff = the rand() the reference
 * draws per ltr_LightAdd, lighter.cpp:1300).  Point/spot: the disk is centred on the light and
 * faces the lumel.  Directional: the disk sits at unit distance along the light direction, i.e.
 * the direction is jittered within atan(light_radius).  The shadow ray starts at the reference's
 * shadow offset SP + SN*0.005 (lighter_int.hpp:966) and is tested with VisibilityTest semantics
 * (any hit, both ends pulled in by 0.001, lighter.cpp:138-147).
 *
 * `smp` is the host-libm table entry of (light, s): directional = the unit sample direction;
 * point/spot = (disk x, disk y, 0).  Only exact IEEE operations below: host and device agree bit
 * for bit, which is what lets the tests check hit/miss against the oracle.
 * ------------------------------------------------------------------------------------------ */
#define LB_SHADOW_OFFSET 0.005f
LB_HD void shadow_sample_segment(unsigned type, V3 Lpos, float range, V3 smp, V3 SP, V3 SN, V3 &from, V3 &to)
{
    from = SP + SN * LB_SHADOW_OFFSET;
    if (type == 3u) { to = from + smp * range; return; }
    const V3 d = norm3(Lpos - from);
    const V3 diffvec = mk3(d.y, -d.z, d.x);
    const V3 up = norm3(cross3(d, diffvec));
    const V3 rt = cross3(d, up);
    to = Lpos + rt * smp.x + up * smp.y;
}

/* USEFUL = true: the caller guarantees !near_zero3(T.n) -- every triangle of the scene BVH: the collision data only takes
 * triangles that pass exactly this test on exactly this cross product (lighter.cpp:349-384 CheckIsUseful; bake.cpp
 * for_useful_tris) -- so the per-test check is dead there. */
template <bool USEFUL = false>
LB_HD float seg_tri_prepared(V3 l1, V3 dir /* = l2 - l1 */, const RayTri &T)
{
    if (!USEFUL && near_zero3(T.n)) return LB_NO_HIT;
    V3 w0 = l1 - T.p1;
    float a = -dot3(T.n, w0);
    float b = dot3(T.n, dir);
    if (fabsf(b) < LB_SMALL) return LB_NO_HIT;
    float r = a / b;
    if (r < 0.0f || r > 1.0f) return LB_NO_HIT;
    V3 I = l1 + r * dir;
    V3 w = I - T.p1;
    float wu = dot3(w, T.u), wv = dot3(w, T.v);
    float s = (T.uv * wv - T.vv * wu) / T.D;
    if (s < 0.0f || s > 1.0f) return LB_NO_HIT;
    float t = (T.uv * wu - T.uu * wv) / T.D;
    if (t < 0.0f || (s + t) > 1.0f) return LB_NO_HIT;
    return r;
}

LB_HD float seg_tri(V3 l1, V3 l2, V3 p1, V3 p2, V3 p3)
{
    RayTri T;
    prepare_raytri(p1, p2, p3, T);
    return seg_tri_prepared(l1, l2 - l1, T);
}

/* ------------------------------------------------------------------------------------------
 * Ray set-up and slab test in the reference's form (ref: lighter_int.hpp:680-702,
 * lighter_math.cpp:618-650): inverse of the NORMALISED direction with 0 for zero components
 * (those axes are skipped), accept when tmax >= tmin and len >= tmin.
 * ------------------------------------------------------------------------------------------ */
struct RefRay { V3 org; float len; V3 inv; };

LB_HD RefRay make_ref_ray(V3 r0, V3 r1)
{
    RefRay r;
    V3 d = r1 - r0;
    r.org = r0;
    r.len = len3(d);
    d = norm3(d);
    r.inv = mk3(d.x ? 1 / d.x : 0, d.y ? 1 / d.y : 0, d.z ? 1 / d.z : 0);
    return r;
}

LB_HD bool ref_ray_box(const RefRay &r, V3 lo, V3 hi)
{
    float tmin = -3.402823466e+38f, tmax = 3.402823466e+38f;
    if (r.inv.x != 0.0f) {
        float a = (lo.x - r.org.x) * r.inv.x, b = (hi.x - r.org.x) * r.inv.x;
        tmin = fmaxr(tmin, fminr(a, b)); tmax = fminr(tmax, fmaxr(a, b));
    }
    if (r.inv.y != 0.0f) {
        float a = (lo.y - r.org.y) * r.inv.y, b = (hi.y - r.org.y) * r.inv.y;
        tmin = fmaxr(tmin, fminr(a, b)); tmax = fminr(tmax, fmaxr(a, b));
    }
    if (r.inv.z != 0.0f) {
        float a = (lo.z - r.org.z) * r.inv.z, b = (hi.z - r.org.z) * r.inv.z;
        tmin = fmaxr(tmin, fminr(a, b)); tmax = fminr(tmax, fmaxr(a, b));
    }
    return tmax >= tmin && r.len >= tmin;
}

/* Hit normal used by the overlap correction (ref: lighter_math.cpp:529, lighter_int.hpp:614-617) */
LB_HD V3 tri_back_normal(V3 p1, V3 p2, V3 p3) { return norm3(cross3(p3 - p1, p2 - p1)); }
