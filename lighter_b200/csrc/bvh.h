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

struct SceneBvh {
    std::vector<BvhNode> nodes;          /* nodes[0] = root */
    std::vector<uint32_t> order;         /* order[k] = original triangle index stored at slot k */
    Box3 bounds;
    int depth = 0;
};

/* tris: 9 floats per triangle (world space).  threads <= 0: hardware concurrency. */
void build_scene_bvh(const float *tris9, size_t count, SceneBvh &out, int leaf_max, int threads);
