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
