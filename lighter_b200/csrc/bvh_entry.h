/*
 * bvh_entry.h -- entry sets: where a BUNDLE of local queries may start walking the scene BVH (host + device).
 *
 * The reference starts every ray at the top of its two-level structure (lighter.cpp:112-136: every instance, then
 * AABBTree::RayQuery from node 0, lighter_int.hpp:742-762).  On the flat scene tree of a 1 M-triangle, 400 x 400 unit
 * scene a radiosity segment is at most 17.84 units long, so roughly the first third of the nodes a ray visits are the
 * descent from the root to its own neighbourhood -- the same nodes for every ray of a spatially compact bundle.
 *
 * bvh4_entry_search() does that descent ONCE for a bundle, given the bundle's bounding box: it keeps a small frontier of
 * nodes whose boxes overlap the box, and repeatedly replaces the largest one by those of its children that overlap, until
 * the frontier is full or every frontier node has a leaf child in range.  The frontier (node + its own box) is the entry
 * set: a ray of the bundle tests the <= BVH_ENTRY_MAX entry boxes and walks only the sub-trees it touches.
 *
 * Correctness: a node is only ever replaced by ALL of its children that overlap the (padded) bundle box, and a node with
 * an overlapping leaf child is never replaced, so every triangle whose leaf box overlaps the bundle box stays reachable.
 * A ray inside the bundle box can only accept leaf boxes that overlap it, hence the set of triangles a ray tests is the
 * same as on a walk from the root -- results are unchanged (checked on the host by tests/test_host.py::test_bvh_entry_set
 * and on the device by the any-hit agreement test).
 */
#pragma once
#include <string.h>
#include "bvh.h"

#ifndef BVH_ENTRY_MAX
#define BVH_ENTRY_MAX 8       /* warp-cooperative search: entry i lives in lane i of an 8-lane group, so <= 8 */
#endif

struct BvhEntrySet {
    int n;
    int node[BVH_ENTRY_MAX];                      /* node indices of the tree searched (Bvh4Node or BvhNode) */
    float lox[BVH_ENTRY_MAX], loy[BVH_ENTRY_MAX], loz[BVH_ENTRY_MAX];
    float hix[BVH_ENTRY_MAX], hiy[BVH_ENTRY_MAX], hiz[BVH_ENTRY_MAX];
};

/* Pad a bundle box so that the slab slack of the per-ray box tests (2e-6 in the segment parameter, gpu_internal.cuh) and
 * the rounding of the shortened end points stay inside it. */
LB_HD void bvh_entry_pad(float &lx, float &ly, float &lz, float &hx, float &hy, float &hz)
{
    const float ext = (hx - lx) + (hy - ly) + (hz - lz);                 /* >= the longest segment in the box */
    const float mag = fmaxf(fmaxf(fmaxf(fabsf(lx), fabsf(hx)), fmaxf(fabsf(ly), fabsf(hy))), fmaxf(fabsf(lz), fabsf(hz)));
    const float pad = 1e-3f + 4e-6f * ext + 2e-6f * mag;
    lx -= pad; ly -= pad; lz -= pad; hx += pad; hy += pad; hz += pad;
}

/* node accessors: the search is the same on the binary tree (closest-hit walks) and on its 4-wide collapse (any-hit walks) */
struct Bvh4Access {
    typedef Bvh4Node Node;
    static constexpr int W = 4;
    static LB_HD int code(const Node &n, int c) { return n.c[c]; }
    static LB_HD bool empty(int code) { return code == BVH4_EMPTY; }
    static LB_HD void box(const Node &n, int c, float &lx, float &ly, float &lz, float &hx, float &hy, float &hz)
    { lx = n.lox[c]; ly = n.loy[c]; lz = n.loz[c]; hx = n.hix[c]; hy = n.hiy[c]; hz = n.hiz[c]; }
};
struct Bvh2Access {
    typedef BvhNode Node;
    static constexpr int W = 2;
    static LB_HD int code(const Node &n, int c) { return c ? n.c1 : n.c0; }
    static LB_HD bool empty(int) { return false; }
    static LB_HD void box(const Node &n, int c, float &lx, float &ly, float &lz, float &hx, float &hy, float &hz)
    {
        if (c) { lx = n.lo1x; ly = n.lo1y; lz = n.lo1z; hx = n.hi1x; hy = n.hi1y; hz = n.hi1z; }
        else   { lx = n.lo0x; ly = n.lo0y; lz = n.lo0z; hx = n.hi0x; hy = n.hi0y; hz = n.hi0z; }
    }
};

/* Scalar, executed by one thread per bundle.  `iters` (optional) receives the number of nodes read. */
template <class A>
LB_HD void bvh_entry_search_t(const typename A::Node *nodes, float qlx, float qly, float qlz, float qhx, float qhy, float qhz,
                              BvhEntrySet &E, int max_entries = BVH_ENTRY_MAX, int *iters = nullptr)
{
    E.n = 1;
    E.node[0] = 0;
    E.lox[0] = E.loy[0] = E.loz[0] = -INFINITY;
    E.hix[0] = E.hiy[0] = E.hiz[0] = INFINITY;
    unsigned fin = 0;                             /* bit i: entry i is final (leaf child in range, or no room for its children) */
    int it = 0;
    for (; it < 64; ++it) {
        int pick = -1;
        float best = -1.f;
        for (int i = 0; i < E.n; ++i) {
            if ((fin >> i) & 1u) continue;
            const float s = (E.hix[i] - E.lox[i]) + (E.hiy[i] - E.loy[i]) + (E.hiz[i] - E.loz[i]);
            if (s > best) { best = s; pick = i; }
        }
        if (pick < 0) break;
        const typename A::Node &N = nodes[E.node[pick]];
        int nh = 0, hcode[A::W];
        float hb[A::W][6];
        bool leaf = false;
        for (int c = 0; c < A::W; ++c) {
            const int code = A::code(N, c);
            if (A::empty(code)) continue;
            float lx, ly, lz, hx, hy, hz;
            A::box(N, c, lx, ly, lz, hx, hy, hz);
            if (lx <= qhx && hx >= qlx && ly <= qhy && hy >= qly && lz <= qhz && hz >= qlz) {
                if (code < 0) leaf = true;
                hcode[nh] = code;
                hb[nh][0] = lx; hb[nh][1] = ly; hb[nh][2] = lz; hb[nh][3] = hx; hb[nh][4] = hy; hb[nh][5] = hz;
                ++nh;
            }
        }
        if (leaf || E.n - 1 + nh > max_entries) { fin |= 1u << pick; continue; }
        if (nh == 0) {                            /* nothing below this node is in range: drop it (the last entry takes its slot) */
            const int last = E.n - 1;
            if (pick != last) {
                E.node[pick] = E.node[last];
                E.lox[pick] = E.lox[last]; E.loy[pick] = E.loy[last]; E.loz[pick] = E.loz[last];
                E.hix[pick] = E.hix[last]; E.hiy[pick] = E.hiy[last]; E.hiz[pick] = E.hiz[last];
                fin = (fin & ~(1u << pick)) | (((fin >> last) & 1u) << pick);
            }
            fin &= ~(1u << last);
            E.n = last;
            continue;
        }
        for (int k = 0; k < nh; ++k) {            /* first child in place, the others appended (their final bits are clear) */
            const int at = k == 0 ? pick : E.n++;
            E.node[at] = hcode[k];
            E.lox[at] = hb[k][0]; E.loy[at] = hb[k][1]; E.loz[at] = hb[k][2];
            E.hix[at] = hb[k][3]; E.hiy[at] = hb[k][4]; E.hiz[at] = hb[k][5];
        }
    }
    if (iters) *iters = it;
}

LB_HD void bvh4_entry_search(const Bvh4Node *nodes, float qlx, float qly, float qlz, float qhx, float qhy, float qhz,
                             BvhEntrySet &E, int max_entries = BVH_ENTRY_MAX, int *iters = nullptr)
{
    bvh_entry_search_t<Bvh4Access>(nodes, qlx, qly, qlz, qhx, qhy, qhz, E, max_entries, iters);
}
LB_HD void bvh2_entry_search(const BvhNode *nodes, float qlx, float qly, float qlz, float qhx, float qhy, float qhz,
                             BvhEntrySet &E, int max_entries = BVH_ENTRY_MAX, int *iters = nullptr)
{
    bvh_entry_search_t<Bvh2Access>(nodes, qlx, qly, qlz, qhx, qhy, qhz, E, max_entries, iters);
}

/* ------------------------------------------------------------------------------------------------------------------
 * Version 2 (the radiosity visibility kernel): shaft-culled search, leaf entries, up to BVH_ENTRY2_MAX entries.
 *
 * A visibility bundle is not an arbitrary cloud of segments: every segment runs from a lumel of one row warp (box R) to a
 * lumel of a few column tiles (box C), so all of them lie in the convex hull of R and C -- the SHAFT -- which for a
 * diagonal bundle is a small part of the bundle's bounding box.  The search drops a child whose box lies outside one of
 * twelve half-spaces that contain the shaft (per axis pair, the four lines through corresponding corners of the two
 * rectangles), so the same number of frontier slots reaches deeper into the tree.  Two more changes: a LEAF in range
 * becomes an entry itself (version 1 stopped refining a node as soon as one of its children was a leaf), and an entry is
 * 32 bytes, so that a ray reads it with two 16-byte shared-memory loads (the frontier may hold up to 16; 8 measured best).
 *
 * Why results cannot change.  (1) A ray tests a triangle iff the plain slab test accepts the triangle's leaf box and every
 * ancestor box on the way down.  Child boxes are contained in their parents' (min / max, no rounding) and rounding is
 * monotone, so the slab values of an ancestor bracket the leaf's: accepting the leaf box implies accepting every ancestor
 * -- starting below the root skips tests that cannot fail.  (2) A dropped box lies farther than `margin` outside a
 * half-space n.(p - r) <= m, where m is the maximum of the left side over the eight corners of R and C, i.e. over the
 * whole shaft, for WHATEVER direction n was chosen (a badly chosen n only makes the plane useless, never wrong).  The
 * margin (2e-3 units + 2e-6 of the largest coordinate, times |n|_1) is two orders of magnitude above everything that
 * separates the float slab test from the geometric one: the 2e-6 slack in the segment parameter (<= 4e-5 units), the
 * rounding of the shortened end points and of the slab products (<= 1e-4 at |p| = 400), and the rounding of the plane
 * evaluation itself.  Host model + self-check (root walk == entry walk on every segment): ltrx_test_bvh_entry2,
 * tests/test_host.py; on the device the lightmap hash and the link sets against the reference.
 * ------------------------------------------------------------------------------------------------------------------ */
#ifndef BVH_ENTRY2_MAX
#define BVH_ENTRY2_MAX 8      /* <= 16.  Measured on B200, config 4 (profiles/r02_ab_runs.md): 4: 222 ms, 6: 214, 8: 211, 10: 212, 12: 214, 16: 221 --
                               * an entry box costs a ray what a quarter of a node visit costs, so a longer frontier stops paying at 8 */
#endif
#define BVH_SHAFT_PLANES 12

struct BvhShaft {                                 /* plane k: axis pair (k / 4): 0 = (x,y), 1 = (y,z), 2 = (z,x) */
    float nu[BVH_SHAFT_PLANES], nv[BVH_SHAFT_PLANES], ru[BVH_SHAFT_PLANES], rv[BVH_SHAFT_PLANES], lim[BVH_SHAFT_PLANES];
};

struct __attribute__((aligned(16))) BvhEntrySet2 {
    float lo[BVH_ENTRY2_MAX][4];                  /* x, y, z, child code as raw bits (>= 0 inner node, < 0 leaf code) */
    float hi[BVH_ENTRY2_MAX][4];                  /* x, y, z, unused */
    float rc[12];                                 /* the bundle's boxes R and C (lo xyz, hi xyz each): scratch of the warp-cooperative search */
    int n, pad[3];
};

/* plane k of the shaft between boxes R and C (6 floats each: lo xyz, hi xyz); maxabs = largest |coordinate| around */
LB_HD void bvh_shaft_plane(const float *R, const float *C, float maxabs, int k, float &nu, float &nv, float &ru, float &rv, float &lim)
{
    const int a = k >> 2, u = a, v = a == 2 ? 0 : a + 1, su = (k >> 1) & 1, sv = k & 1;
    ru = R[u + 3 * su]; rv = R[v + 3 * sv];
    const float cu = C[u + 3 * su], cv = C[v + 3 * sv];
    nu = -(cv - rv); nv = cu - ru;
    /* outward: R's opposite corner must be on the inner side */
    if (nu * (R[u + 3 * (1 - su)] - ru) + nv * (R[v + 3 * (1 - sv)] - rv) > 0.f) { nu = -nu; nv = -nv; }
    float m = 0.f;                                /* r itself gives 0 */
    for (int j = 0; j < 8; ++j) {
        const float *B = (j & 4) ? C : R;
        m = fmaxf(m, nu * (B[u + 3 * (j & 1)] - ru) + nv * (B[v + 3 * ((j >> 1) & 1)] - rv));
    }
    lim = m + (fabsf(nu) + fabsf(nv)) * (2e-3f + 2e-6f * maxabs);
    if (!(lim == lim) || !(fabsf(lim) < 3e38f)) { nu = nv = 0.f; lim = 0.f; }         /* inf / nan boxes: a plane that rejects nothing */
}

LB_HD void bvh_shaft_build(const float *R, const float *C, float maxabs, BvhShaft &S)
{
    for (int k = 0; k < BVH_SHAFT_PLANES; ++k) bvh_shaft_plane(R, C, maxabs, k, S.nu[k], S.nv[k], S.ru[k], S.rv[k], S.lim[k]);
}

/* is the box certainly outside the shaft? */
LB_HD bool bvh_shaft_outside(const BvhShaft &S, float lx, float ly, float lz, float hx, float hy, float hz)
{
    const float lo[3] = { lx, ly, lz }, hi[3] = { hx, hy, hz };
    for (int k = 0; k < BVH_SHAFT_PLANES; ++k) {
        const int a = k >> 2, u = a, v = a == 2 ? 0 : a + 1;
        const float bu = S.nu[k] > 0.f ? lo[u] : hi[u], bv = S.nv[k] > 0.f ? lo[v] : hi[v];
        if (S.nu[k] * (bu - S.ru[k]) + S.nv[k] * (bv - S.rv[k]) > S.lim[k]) return true;
    }
    return false;
}

/* Scalar model of the version-2 search on the 4-wide tree (the kernels run the warp-cooperative form in gpu_internal.cuh,
 * which makes the same picks).  q: padded bundle box; S: shaft planes or nullptr. */
LB_HD void bvh4_entry_search2(const Bvh4Node *nodes, float qlx, float qly, float qlz, float qhx, float qhy, float qhz, const BvhShaft *S,
                              BvhEntrySet2 &E, int max_entries = BVH_ENTRY2_MAX, int *iters = nullptr)
{
    int code[BVH_ENTRY2_MAX];
    E.n = 1;
    code[0] = 0;
    E.lo[0][0] = E.lo[0][1] = E.lo[0][2] = -INFINITY;
    E.hi[0][0] = E.hi[0][1] = E.hi[0][2] = INFINITY;
    unsigned fin = 0;                             /* bit i: entry i is final (a leaf, or no room for its children) */
    int it = 0;
    for (; it < 128; ++it) {
        int pick = -1;
        float best = -1.f;
        for (int i = 0; i < E.n; ++i) {
            if ((fin >> i) & 1u) continue;
            const float s = (E.hi[i][0] - E.lo[i][0]) + (E.hi[i][1] - E.lo[i][1]) + (E.hi[i][2] - E.lo[i][2]);
            if (s > best) { best = s; pick = i; }
        }
        if (pick < 0) break;
        const Bvh4Node &N = nodes[code[pick]];
        int nh = 0, hcode[4];
        float hb[4][6];
        for (int c = 0; c < 4; ++c) {
            if (N.c[c] == BVH4_EMPTY) continue;
            const float lx = N.lox[c], ly = N.loy[c], lz = N.loz[c], hx = N.hix[c], hy = N.hiy[c], hz = N.hiz[c];
            if (!(lx <= qhx && hx >= qlx && ly <= qhy && hy >= qly && lz <= qhz && hz >= qlz)) continue;
            if (S && bvh_shaft_outside(*S, lx, ly, lz, hx, hy, hz)) continue;
            hcode[nh] = N.c[c];
            hb[nh][0] = lx; hb[nh][1] = ly; hb[nh][2] = lz; hb[nh][3] = hx; hb[nh][4] = hy; hb[nh][5] = hz;
            ++nh;
        }
        if (E.n - 1 + nh > max_entries) { fin |= 1u << pick; continue; }
        if (nh == 0) {                            /* nothing below this node is in range: drop it (the last entry takes its slot) */
            const int last = E.n - 1;
            if (pick != last) {
                code[pick] = code[last];
                for (int k = 0; k < 3; ++k) { E.lo[pick][k] = E.lo[last][k]; E.hi[pick][k] = E.hi[last][k]; }
                fin = (fin & ~(1u << pick)) | (((fin >> last) & 1u) << pick);
            }
            fin &= ~(1u << last);
            E.n = last;
            continue;
        }
        for (int k = 0; k < nh; ++k) {            /* first child in place, the others appended; a leaf is final at once */
            const int at = k == 0 ? pick : E.n++;
            code[at] = hcode[k];
            for (int j = 0; j < 3; ++j) { E.lo[at][j] = hb[k][j]; E.hi[at][j] = hb[k][3 + j]; }
            fin = hcode[k] < 0 ? (fin | (1u << at)) : (fin & ~(1u << at));
        }
    }
    for (int i = 0; i < E.n; ++i) { memcpy(&E.lo[i][3], &code[i], 4); E.hi[i][3] = 0.f; }
    if (iters) *iters = it;
}
