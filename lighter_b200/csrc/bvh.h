/*
 * bvh.h -- the flat scene BVH used by every hot query (distance march, any-hit, closest-hit).
 *
 * Replaces the reference's two-level structure (instance box tree -> per-instance TriTree,
 * lighter.cpp:112-289, lighter_int.hpp:706-833) by ONE tree over the world-space triangles of all
 * shadow-casting instances.  Results of min/any/closest queries do not depend on tree topology
 * (SURVEY.md finding 3), so the topology is chosen for the GPU: binned-SAH binary tree, children's
 * boxes stored in the parent (one 64-byte, 16B-aligned node = 4 float4 loads decides both
 * children), leaves of <= LEAF_MAX triangles referenced as a contiguous range of the re-ordered
 * triangle arrays.
 */
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <new>
#include <utility>
#include <vector>
#include "vmath.h"

struct BvhNode {                 /* 64 bytes */
    float lo0x, lo0y, lo0z, hi0x;
    float hi0y, hi0z, lo1x, lo1y;
    float lo1z, hi1x, hi1y, hi1z;
    int32_t c0, c1;              /* >=0 inner node index; <0: leaf, ~c = (first_tri << 3) | count */
    int32_t pad0, pad1;
};

#define BVH_LEAF_MAX 2      /* measured on B200 (configs 3, 4): leaves of <=2 beat 4 by 10-12 % and 7 by 25-30 % on the traversal kernels */
#define BVH_STACK 64

/* The same tree with every other level removed: a node holds the boxes of up to four grand-children (SoA, so each
 * coordinate of the four boxes is one float4 load).  Used by the any-hit walks, which visit every overlapped node anyway:
 * half the iterations, half the loop/stack overhead per box test (the walks are instruction-issue bound). */
#define BVH4_EMPTY 0x7fffffff
struct Bvh4Node {                /* 128 bytes */
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    int32_t c[4];                /* >=0 inner Bvh4Node index; <0 leaf (same code as BvhNode); BVH4_EMPTY unused slot */
    int32_t pad[4];
};

/* A/B variant (LB_VIS_Q8=1, profiles/r02_ab_runs.md): the 4-wide node compressed to 64 bytes -- the four child boxes as 8-bit
 * offsets from the node's own corner in power-of-two steps (north_star's "compressed nodes"; the reference's node is
 * lighter_int.hpp:708-714).  Planes are rounded OUTWARDS and widened by one more step, so the decoded boxes contain the float
 * boxes with at least one step to spare (that step absorbs the rounding of the walk's fused decode, gpu_internal.cuh). */
struct Bvh4QNode {               /* 64 bytes = 4 x 16 */
    float ox, oy, oz;            /* corner: one step below the lowest child plane */
    uint32_t exps;               /* biased float exponents of the three steps: ex | ey << 8 | ez << 16 */
    uint32_t qlox, qloy, qloz, qhix;   /* byte c of each word = child c */
    uint32_t qhiy, qhiz;
    int32_t c[4];                /* as Bvh4Node::c */
    uint32_t pad[2];
};

/* Allocator of the builder's big arrays: elements are default-INITIALISED (resize() does not zero 70 MB that the flatten
 * passes overwrite anyway) and large blocks ask for transparent huge pages (THP "madvise" mode; see bvh.cpp). */
void *lb_big_alloc(size_t bytes);
template <class T> struct LbBigAlloc {
    typedef T value_type;
    LbBigAlloc() = default;
    template <class U> LbBigAlloc(const LbBigAlloc<U> &) {}
    T *allocate(size_t n) { return static_cast<T *>(lb_big_alloc(n * sizeof(T))); }
    void deallocate(T *p, size_t) { free(p); }
    template <class U> void construct(U *) {}                                           /* default-init: left as is */
    template <class U, class A0, class... A> void construct(U *p, A0 &&a0, A &&...a) { ::new ((void *)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
    template <class U> bool operator==(const LbBigAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const LbBigAlloc<U> &) const { return false; }
};
template <class T> using BigVec = std::vector<T, LbBigAlloc<T>>;

struct SceneBvh {
    BigVec<BvhNode> nodes;               /* nodes[0] = root */
    BigVec<Bvh4Node> nodes4;             /* nodes4[0] = root of the collapsed tree (build_bvh4) */
    BigVec<uint32_t> order;              /* order[k] = original triangle index stored at slot k */
    Box3 bounds;
    int depth = 0;
};

/* tris: 9 floats per triangle (world space).  threads <= 0: hardware concurrency. */
void build_scene_bvh(const float *tris9, size_t count, SceneBvh &out, int leaf_max, int threads);
/* out.nodes4 from out.nodes by serial collapse (slots = grand-children; a leaf child stays a slot).  build_scene_bvh already
 * fills nodes4 (in parallel, straight from its build tree); this is the reference implementation used for degenerate trees. */
void build_bvh4(SceneBvh &bvh);
