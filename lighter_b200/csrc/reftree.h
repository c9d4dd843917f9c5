/*
 * reftree.h -- "reference-order" box tree.
 *
 * Two results of the reference depend on the ORDER in which its per-instance box tree hands out
 * triangles, not just on the set: the concave-edge sample offset mutates the sample position while
 * it traverses (lighter_math.cpp:991-1044), and closest-hit ties keep the first triangle met
 * (lighter_math.cpp:841-853).  A third depends on node granularity: light -> instance culling
 * returns every item of any overlapped leaf of the instance tree (lighter_int.hpp:786-806,
 * lighter.cpp:80-98).  For exactly those three uses we rebuild a tree with the same topology and
 * item order as the reference would (algorithm: lighter_math.cpp:674-781 -- median split of the
 * "splittable" boxes along the longest axis of their union, boxes with >= 1/3 of the node volume
 * stay in the inner node, depth <= 16, leaf <= 4; node 0 is the root, the first child follows its
 * parent, items are stored as <count> ids...).  Every hot query (distance march, shadow/AO/radiosity
 * segments) uses the flat scene BVH instead (bvh.h), whose topology is free (SURVEY.md finding 3).
 */
#pragma once
#include <stdint.h>
#include <vector>
#include "vmath.h"

struct RefNode {           /* 32 bytes, uploaded as is */
    V3 lo, hi;
    int32_t ch;            /* second child; first child is node+1; -1 = no children */
    int32_t ido;           /* offset of "<count> ids..." in items, -1 = none */
};

struct RefTree {
    std::vector<RefNode> nodes;
    std::vector<int32_t> items;

    /* threads > 1: the halves of big nodes are built concurrently (same bytes as the serial build; reftree.cpp) */
    void build(const Box3 *boxes, size_t count, int threads = 1);

    /* box query in pre-order (node items, first child, second child); calls f(ids,count) */
    template <class F> void query(V3 qlo, V3 qhi, F &f, int32_t node = 0) const
    {
        const RefNode &N = nodes[node];
        if (qlo.x > N.hi.x || qhi.x < N.lo.x || qlo.y > N.hi.y || qhi.y < N.lo.y || qlo.z > N.hi.z || qhi.z < N.lo.z) return;
        if (N.ido != -1) f(&items[N.ido + 1], items[N.ido]);
        if (N.ch != -1) { query(qlo, qhi, f, node + 1); query(qlo, qhi, f, N.ch); }
    }
    /* every item, in storage order (ref: lighter_int.hpp:808-814) */
    template <class F> void all(F &f) const
    {
        for (size_t i = 0; i < items.size(); i += 1 + items[i]) f(&items[i + 1], items[i]);
    }
};
