/*
 * gpu_bvh.cu -- the flat scene BVH (bvh.h) built ON THE DEVICE.
 *
 * Replaces Job_ColInfo_Inner / AABBTree::SetAABBs / _MakeNode (lighter.cpp:349-384, lighter_math.cpp:674-781) for every hot
 * query: the reference builds one median-split tree per instance on its thread pool; the topology of an acceleration
 * structure is free (SURVEY finding 3: min / any / closest queries equal brute force on any conservative tree), so the tree
 * here is chosen for the GPU and built by it -- the host build was 70-96 ms of a 0.55 s bake on 16 cores and was repeated by
 * every rank of a multi-GPU bake.
 *
 * Algorithm: the SAME binned SAH as the host builder in bvh.cpp (16 bins x 3 axes over the centroid bounds, cost =
 * area x count of the two sides, split at the cheapest non-empty bin boundary, leaves of <= leaf_max triangles), run
 * level-synchronously:
 *
 *   every level      nodes of the level are a contiguous range of the build-node array; a triangle ("prim": box + id, 32 B)
 *                    knows its node, prims of a node are contiguous in a ping-pong prim array;
 *   big nodes        (> 32 prims) bin kernel over the prims (warp-aggregated u32 atomics on order-preserving float codes:
 *                    min / max / count are exact and order independent), split kernel (one thread per node sweeps its 48
 *                    bins), partition kernel over the prims (warp-aggregated cursors; the children's box and centroid
 *                    bounds are reduced on the way);
 *   small nodes      (<= 32 prims) one warp per node does all three steps in registers / shared memory, no global atomics;
 *   after the last   sub-tree sizes bottom-up, PRE-ORDER slots top-down (first child = next slot, as bvh.cpp lays the tree
 *   level            out), then one thread per inner node writes its BvhNode and, at even depth, its 4-wide Bvh4Node.
 *
 * The arithmetic of binning and of the cost sweep is bvh.cpp's, operation for operation (compiled -fmad=false), and the bin
 * statistics are exact, so the tree equals the host builder's for non-degenerate input (tests/test_gpu_bvh.py compares them
 * node by node); leaves list their triangles in ascending original index, which makes the result independent of the order
 * in which atomics land.  Degenerate input: coincident centroids split by position; below SAH_DEPTH_MAX levels every node
 * splits by position, which bounds the height (the traversal stacks are fixed, BVH_STACK).
 */
#include "gpu_internal.cuh"

#include <stdlib.h>

#define GB_NB 16                      /* SAH bins per axis (bvh.cpp: NB) */
#define GB_SMALL 32u                  /* nodes up to this many prims are built by one warp */
#define GB_MAX_LEVELS 72
#define GB_HEIGHT_MAX 40              /* what the fixed traversal stacks allow (4-wide walk: 3 pushes per two levels) */

struct GbPrim { float4 lo, hi; };     /* lo.w = original triangle index (bits) */

struct GbBin { uint32_t lo[3], hi[3], cnt, pad; };          /* order-preserving codes; empty = lo 0xffffffff, hi 0, cnt 0 */

struct GbNode {
    uint32_t blo[3], bhi[3];          /* box of the prims (codes) */
    uint32_t clo[3], chi[3];          /* box of their centroids (codes) */
    uint32_t first, count;
    int32_t left, right;              /* -1: not split (a leaf when count <= leaf_max) */
    uint32_t lcount, lcur, rcur;
    int32_t slot;                     /* bin block of a big node, -1 otherwise */
    uint32_t axis, split, bypos, depth;
    int32_t inner, inner_even, height;
    int32_t out_slot, out_slot4;
    int32_t pad[3];
};

struct GbState {
    uint32_t node_count;
    uint32_t slot_count[2];
    uint32_t lvl_begin[GB_MAX_LEVELS + 2];
};

__device__ __forceinline__ uint32_t gb_enc(float f)
{
    f += 0.0f;                                                 /* -0 -> +0: one code per value */
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float gb_dec(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__device__ __forceinline__ int gb_clamp_bin(int b) { return b < 0 ? 0 : (b >= GB_NB ? GB_NB - 1 : b); }

/* bins of a prim's centroid on the three axes: the arithmetic of bvh.cpp bin_range (unused axis: scale 0 -> bin 0) */
struct GbBinning { float lo[3], scale[3]; bool use[3]; };
__device__ __forceinline__ GbBinning gb_binning(const GbNode &N)
{
    GbBinning B;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float l = gb_dec(N.clo[a]), h = gb_dec(N.chi[a]);
        const float ext = h - l;
        B.lo[a] = l;
        B.use[a] = ext > 0;
        B.scale[a] = B.use[a] ? (float)GB_NB / ext : 0.f;
    }
    return B;
}
__device__ __forceinline__ void gb_bins_of(const GbBinning &B, const GbPrim &P, int bin[3])
{
    const float cx = (P.lo.x + P.hi.x) * 0.5f, cy = (P.lo.y + P.hi.y) * 0.5f, cz = (P.lo.z + P.hi.z) * 0.5f;
    bin[0] = gb_clamp_bin((int)((cx - B.lo[0]) * B.scale[0]));
    bin[1] = gb_clamp_bin((int)((cy - B.lo[1]) * B.scale[1]));
    bin[2] = gb_clamp_bin((int)((cz - B.lo[2]) * B.scale[2]));
}

__device__ __forceinline__ float gb_half_area(const float lo[3], const float hi[3])
{
    const float e0 = hi[0] - lo[0], e1 = hi[1] - lo[1], e2 = hi[2] - lo[2];
    return e0 * e1 + e1 * e2 + e2 * e0;
}

/* The cost sweep of one axis over its 16 bins (bvh.cpp split(): non-empty bins only, positions high to low, strict <).
 * Returns false when the axis offers no split; otherwise cost / split bin / prims on the left. */
__device__ __forceinline__ bool gb_sweep_axis(const GbBin *bins /* GB_NB of this axis */, float &cost_out, int &split_out, uint32_t &lcount_out)
{
    float la[GB_NB]; uint32_t lc[GB_NB]; int bs[GB_NB];
    int K = 0;
    float alo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f }, ahi[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
    uint32_t c = 0;
    for (int b = 0; b < GB_NB; ++b) {
        const uint32_t n = bins[b].cnt;
        if (!n) continue;
#pragma unroll
        for (int k = 0; k < 3; ++k) { alo[k] = fminf(alo[k], gb_dec(bins[b].lo[k])); ahi[k] = fmaxf(ahi[k], gb_dec(bins[b].hi[k])); }
        c += n;
        la[K] = gb_half_area(alo, ahi); lc[K] = c; bs[K] = b; ++K;
    }
    if (K < 2) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) { alo[k] = 3.402823466e+38f; ahi[k] = -3.402823466e+38f; }
    c = 0;
    float best = 3.402823466e+38f;
    int best_b = -1;
    uint32_t best_l = 0;
    for (int k = K - 1; k > 0; --k) {
        const int b = bs[k];
#pragma unroll
        for (int q = 0; q < 3; ++q) { alo[q] = fminf(alo[q], gb_dec(bins[b].lo[q])); ahi[q] = fmaxf(ahi[q], gb_dec(bins[b].hi[q])); }
        c += bins[b].cnt;
        const float cost = la[k - 1] * (float)lc[k - 1] + gb_half_area(alo, ahi) * (float)c;
        if (cost < best) { best = cost; best_b = b; best_l = lc[k - 1]; }
    }
    if (best_b < 0) return false;
    cost_out = best; split_out = best_b; lcount_out = best_l;
    return true;
}

__device__ __forceinline__ void gb_child_init(GbNode &C, uint32_t first, uint32_t count, uint32_t depth)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) { C.blo[k] = 0xffffffffu; C.bhi[k] = 0u; C.clo[k] = 0xffffffffu; C.chi[k] = 0u; }
    C.first = first; C.count = count; C.left = C.right = -1;
    C.lcount = C.lcur = C.rcur = 0; C.slot = -1; C.axis = C.split = C.bypos = 0; C.depth = depth;
    C.inner = C.inner_even = C.height = 0; C.out_slot = C.out_slot4 = -1;
}

/* ---- level 0: prims from the triangles, root bounds ------------------------------------------------------------------- */
__global__ void gb_init_kernel(const float *__restrict__ tris9, uint32_t n, GbPrim *__restrict__ prim, int32_t *__restrict__ pnode, GbNode *nodes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint32_t e[12];
#pragma unroll
    for (int k = 0; k < 3; ++k) { e[k] = 0xffffffffu; e[3 + k] = 0u; e[6 + k] = 0xffffffffu; e[9 + k] = 0u; }
    if (live) {
        const float *q = tris9 + 9ull * i;
        const V3 a = mk3(q[0], q[1], q[2]), b = mk3(q[3], q[4], q[5]), c = mk3(q[6], q[7], q[8]);
        const V3 lo = min3(a, min3(b, c)), hi = max3(a, max3(b, c));
        GbPrim P;
        P.lo = make_float4(lo.x, lo.y, lo.z, __uint_as_float(i));
        P.hi = make_float4(hi.x, hi.y, hi.z, 0.f);
        prim[i] = P;
        pnode[i] = 0;
        const float cx = (lo.x + hi.x) * 0.5f, cy = (lo.y + hi.y) * 0.5f, cz = (lo.z + hi.z) * 0.5f;
        e[0] = gb_enc(lo.x); e[1] = gb_enc(lo.y); e[2] = gb_enc(lo.z); e[3] = gb_enc(hi.x); e[4] = gb_enc(hi.y); e[5] = gb_enc(hi.z);
        e[6] = gb_enc(cx); e[7] = gb_enc(cy); e[8] = gb_enc(cz); e[9] = e[6]; e[10] = e[7]; e[11] = e[8];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        e[k] = __reduce_min_sync(0xffffffffu, e[k]); e[3 + k] = __reduce_max_sync(0xffffffffu, e[3 + k]);
        e[6 + k] = __reduce_min_sync(0xffffffffu, e[6 + k]); e[9 + k] = __reduce_max_sync(0xffffffffu, e[9 + k]);
    }
    if ((threadIdx.x & 31u) == 0) {
        GbNode &R = nodes[0];
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&R.blo[k], e[k]); atomicMax(&R.bhi[k], e[3 + k]); atomicMin(&R.clo[k], e[6 + k]); atomicMax(&R.chi[k], e[9 + k]); }
    }
}

__global__ void gb_setup_kernel(GbNode *nodes, GbState *S, uint32_t n)
{
    if (blockIdx.x || threadIdx.x) return;
    gb_child_init(nodes[0], 0, n, 0);
    nodes[0].slot = n > GB_SMALL ? 0 : -1;
    S->node_count = 1;
    S->slot_count[0] = 1; S->slot_count[1] = 0;
    for (int l = 0; l < GB_MAX_LEVELS + 2; ++l) S->lvl_begin[l] = l == 0 ? 0u : 1u;
}

__global__ void gb_fill_bins_kernel(GbBin *bins, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GbBin b;
    b.lo[0] = b.lo[1] = b.lo[2] = 0xffffffffu; b.hi[0] = b.hi[1] = b.hi[2] = 0u; b.cnt = 0u; b.pad = 0u;
    bins[i] = b;
}

/* ---- big nodes, step 1: binning ------------------------------------------------------------------------------------------ */
__global__ void gb_bin_kernel(const GbPrim *__restrict__ prim, const int32_t *__restrict__ pnode, uint32_t n, const GbNode *__restrict__ nodes,
                              GbBin *__restrict__ bins, uint16_t *__restrict__ pbin, uint32_t sah_depth_max, uint32_t leaf_max)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    int32_t node = i < n ? pnode[i] : -1;
    bool act = false;
    int bin[3] = { 0, 0, 0 };
    uint32_t e[6] = { 0, 0, 0, 0, 0, 0 };
    int32_t slot = -1;
    if (node >= 0) {
        const GbNode &N = nodes[node];
        if (N.count > GB_SMALL && N.slot >= 0 && N.depth < sah_depth_max) {
            act = true;
            slot = N.slot;
            const GbPrim P = prim[i];
            gb_bins_of(gb_binning(N), P, bin);
            pbin[i] = (uint16_t)(bin[0] | (bin[1] << 4) | (bin[2] << 8));
            e[0] = gb_enc(P.lo.x); e[1] = gb_enc(P.lo.y); e[2] = gb_enc(P.lo.z); e[3] = gb_enc(P.hi.x); e[4] = gb_enc(P.hi.y); e[5] = gb_enc(P.hi.z);
        }
    }
    (void)leaf_max;
    const int32_t s0 = __shfl_sync(0xffffffffu, slot, 0);
    const bool uniform = __all_sync(0xffffffffu, act && slot == s0);
    if (uniform) {
        GbBin *B = bins + (size_t)slot * (3 * GB_NB);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const unsigned m = __match_any_sync(0xffffffffu, bin[a]);
            const uint32_t l0 = __reduce_min_sync(m, e[0]), l1 = __reduce_min_sync(m, e[1]), l2 = __reduce_min_sync(m, e[2]);
            const uint32_t h0 = __reduce_max_sync(m, e[3]), h1 = __reduce_max_sync(m, e[4]), h2 = __reduce_max_sync(m, e[5]);
            if (lane == (unsigned)__ffs(m) - 1u) {
                GbBin &b = B[a * GB_NB + bin[a]];
                atomicMin(&b.lo[0], l0); atomicMin(&b.lo[1], l1); atomicMin(&b.lo[2], l2);
                atomicMax(&b.hi[0], h0); atomicMax(&b.hi[1], h1); atomicMax(&b.hi[2], h2);
                atomicAdd(&b.cnt, (uint32_t)__popc(m));
            }
        }
    } else if (act) {
        GbBin *B = bins + (size_t)slot * (3 * GB_NB);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            GbBin &b = B[a * GB_NB + bin[a]];
            atomicMin(&b.lo[0], e[0]); atomicMin(&b.lo[1], e[1]); atomicMin(&b.lo[2], e[2]);
            atomicMax(&b.hi[0], e[3]); atomicMax(&b.hi[1], e[4]); atomicMax(&b.hi[2], e[5]);
            atomicAdd(&b.cnt, 1u);
        }
    }
}

/* ---- big nodes, step 2: pick the split, create the children ------------------------------------------------------------ */
__global__ void gb_split_kernel(GbNode *nodes, GbState *S, uint32_t level, GbBin *bins_cur, uint32_t sah_depth_max, uint32_t leaf_max, uint32_t max_slots)
{
    const uint32_t lb = S->lvl_begin[level], le = S->lvl_begin[level + 1];
    for (uint32_t j = lb + blockIdx.x * blockDim.x + threadIdx.x; j < le; j += gridDim.x * blockDim.x) {
        GbNode &N = nodes[j];
        if (N.count <= GB_SMALL) continue;                        /* small nodes: gb_small_kernel; leaves: nothing to do */
        int best_axis = -1, best_split = 0;
        uint32_t best_l = 0;
        if (N.slot >= 0 && N.depth < sah_depth_max) {
            GbBin *B = bins_cur + (size_t)N.slot * (3 * GB_NB);
            const GbBinning bn = gb_binning(N);
            float best_cost = 3.402823466e+38f;
            for (int a = 0; a < 3; ++a) {
                if (!bn.use[a]) continue;
                float cost; int sp; uint32_t lc;
                if (gb_sweep_axis(B + a * GB_NB, cost, sp, lc) && cost < best_cost) { best_cost = cost; best_axis = a; best_split = sp; best_l = lc; }
            }
            for (int b = 0; b < 3 * GB_NB; ++b)                      /* leave the block clean for the node that gets it next */
                if (B[b].cnt) { B[b].lo[0] = B[b].lo[1] = B[b].lo[2] = 0xffffffffu; B[b].hi[0] = B[b].hi[1] = B[b].hi[2] = 0u; B[b].cnt = 0u; }
        }
        if (best_axis < 0) { N.bypos = 1; best_l = N.count / 2; }      /* coincident centroids, or below the SAH depth limit */
        N.axis = (uint32_t)(best_axis < 0 ? 0 : best_axis); N.split = (uint32_t)best_split; N.lcount = best_l;
        const uint32_t c0 = atomicAdd(&S->node_count, 2u);
        N.left = (int32_t)c0; N.right = (int32_t)c0 + 1;
        GbNode L, R;
        gb_child_init(L, N.first, best_l, N.depth + 1);
        gb_child_init(R, N.first + best_l, N.count - best_l, N.depth + 1);
        if (L.count > GB_SMALL && L.depth < sah_depth_max) { const uint32_t s = atomicAdd(&S->slot_count[(level + 1) & 1u], 1u); L.slot = s < max_slots ? (int32_t)s : -1; }
        if (R.count > GB_SMALL && R.depth < sah_depth_max) { const uint32_t s = atomicAdd(&S->slot_count[(level + 1) & 1u], 1u); R.slot = s < max_slots ? (int32_t)s : -1; }
        nodes[c0] = L; nodes[c0 + 1] = R;
    }
    (void)leaf_max;
}

/* ---- small nodes: one warp does binning, sweep and partition ----------------------------------------------------------- */
#define GB_SMALL_WARPS 8
__global__ void __launch_bounds__(GB_SMALL_WARPS * 32)
gb_small_kernel(const GbPrim *__restrict__ prim_in, GbPrim *__restrict__ prim_out, int32_t *__restrict__ pnode_out, GbNode *nodes, GbState *S, uint32_t level,
                uint32_t sah_depth_max, uint32_t leaf_max)
{
    __shared__ GbBin s_bins[GB_SMALL_WARPS][3 * GB_NB];
    __shared__ float s_cost[GB_SMALL_WARPS][3];
    __shared__ int s_split[GB_SMALL_WARPS][3];
    __shared__ uint32_t s_lc[GB_SMALL_WARPS][3];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t lb = S->lvl_begin[level], le = S->lvl_begin[level + 1];
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t j = lb + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); j < le; j += warps) {
        GbNode &N = nodes[j];
        const uint32_t count = N.count, first = N.first, depth = N.depth;
        if (count > GB_SMALL || count <= leaf_max) continue;         /* warp-uniform */
        const bool live = lane < count;
        GbPrim P;
        P.lo = make_float4(0.f, 0.f, 0.f, 0.f); P.hi = P.lo;
        if (live) P = prim_in[first + lane];
        int best_axis = -1, best_split = 0;
        uint32_t lcount = 0;
        int bin[3] = { 0, 0, 0 };
        if (depth < sah_depth_max) {
            const GbBinning bn = gb_binning(N);
            for (unsigned b = lane; b < 3u * GB_NB; b += 32u) {
                GbBin &x = s_bins[w][b];
                x.lo[0] = x.lo[1] = x.lo[2] = 0xffffffffu; x.hi[0] = x.hi[1] = x.hi[2] = 0u; x.cnt = 0u;
            }
            __syncwarp();
            if (live) {
                gb_bins_of(bn, P, bin);
                const uint32_t e0 = gb_enc(P.lo.x), e1 = gb_enc(P.lo.y), e2 = gb_enc(P.lo.z), e3 = gb_enc(P.hi.x), e4 = gb_enc(P.hi.y), e5 = gb_enc(P.hi.z);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    GbBin &x = s_bins[w][a * GB_NB + bin[a]];
                    atomicMin(&x.lo[0], e0); atomicMin(&x.lo[1], e1); atomicMin(&x.lo[2], e2);
                    atomicMax(&x.hi[0], e3); atomicMax(&x.hi[1], e4); atomicMax(&x.hi[2], e5);
                    atomicAdd(&x.cnt, 1u);
                }
            }
            __syncwarp();
            if (lane < 3) {
                float cost = 3.402823466e+38f; int sp = -1; uint32_t lc = 0;
                const bool ok = bn.use[lane] && gb_sweep_axis(&s_bins[w][lane * GB_NB], cost, sp, lc);
                s_cost[w][lane] = ok ? cost : 3.402823466e+38f; s_split[w][lane] = ok ? sp : -1; s_lc[w][lane] = lc;
            }
            __syncwarp();
            float best_cost = 3.402823466e+38f;
            for (int a = 0; a < 3; ++a)
                if (s_split[w][a] >= 0 && s_cost[w][a] < best_cost) { best_cost = s_cost[w][a]; best_axis = a; best_split = s_split[w][a]; lcount = s_lc[w][a]; }
            __syncwarp();
        }
        bool left;
        if (best_axis < 0) { lcount = count / 2; left = lane < lcount; }
        else left = (best_axis == 0 ? bin[0] : best_axis == 1 ? bin[1] : bin[2]) < best_split;
        const unsigned live_m = __ballot_sync(0xffffffffu, live);
        const unsigned mL = __ballot_sync(0xffffffffu, live && left), mR = live_m & ~mL;
        lcount = (uint32_t)__popc(mL);                                /* equals the swept count */
        uint32_t c0 = 0;
        if (lane == 0) c0 = atomicAdd(&S->node_count, 2u);
        c0 = __shfl_sync(0xffffffffu, c0, 0);
        /* children bounds: reduce over the lanes of each side */
        const float cx = (P.lo.x + P.hi.x) * 0.5f, cy = (P.lo.y + P.hi.y) * 0.5f, cz = (P.lo.z + P.hi.z) * 0.5f;
        const uint32_t v[12] = { gb_enc(P.lo.x), gb_enc(P.lo.y), gb_enc(P.lo.z), gb_enc(P.hi.x), gb_enc(P.hi.y), gb_enc(P.hi.z),
                                 gb_enc(cx), gb_enc(cy), gb_enc(cz), gb_enc(cx), gb_enc(cy), gb_enc(cz) };
        if (live) {
            const unsigned m = left ? mL : mR;
            uint32_t r[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) r[k] = ((k / 3) & 1) ? __reduce_max_sync(m, v[k]) : __reduce_min_sync(m, v[k]);
            if (lane == (unsigned)__ffs(m) - 1u) {
                GbNode C;
                gb_child_init(C, left ? first : first + lcount, left ? lcount : count - lcount, depth + 1);
#pragma unroll
                for (int k = 0; k < 3; ++k) { C.blo[k] = r[k]; C.bhi[k] = r[3 + k]; C.clo[k] = r[6 + k]; C.chi[k] = r[9 + k]; }
                nodes[c0 + (left ? 0u : 1u)] = C;
            }
            const uint32_t at = left ? first + (uint32_t)__popc(mL & ((1u << lane) - 1u)) : first + lcount + (uint32_t)__popc(mR & ((1u << lane) - 1u));
            prim_out[at] = P;
            pnode_out[at] = (int32_t)(c0 + (left ? 0u : 1u));
        }
        if (lane == 0) {
            N.left = (int32_t)c0; N.right = (int32_t)c0 + 1; N.lcount = lcount;
            N.axis = (uint32_t)(best_axis < 0 ? 0 : best_axis); N.split = (uint32_t)best_split; N.bypos = best_axis < 0 ? 1u : 0u;
        }
        __syncwarp();
    }
}

/* ---- big nodes, step 3: partition; leaves: final order; everything else: retire --------------------------------------- */
__global__ void gb_partition_kernel(const GbPrim *__restrict__ prim_in, const int32_t *__restrict__ pnode_in, GbPrim *__restrict__ prim_out,
                                    int32_t *__restrict__ pnode_out, uint32_t n, GbNode *nodes, const uint16_t *__restrict__ pbin,
                                    uint32_t *__restrict__ order, uint32_t leaf_max)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const int32_t node = i < n ? pnode_in[i] : -1;
    bool big = false, left = false;
    GbPrim P;
    P.lo = make_float4(0.f, 0.f, 0.f, 0.f); P.hi = P.lo;
    if (node >= 0) {
        GbNode &N = nodes[node];
        if (N.count <= leaf_max) {                                     /* leaf: its slots of the final order (sorted later) */
            const GbPrim Q = prim_in[i];
            order[N.first + atomicAdd(&N.lcur, 1u)] = __float_as_uint(Q.lo.w);
            pnode_out[i] = -1;
        } else if (N.count > GB_SMALL) {
            big = true;
            P = prim_in[i];
            if (N.bypos) left = (i - N.first) < N.lcount;
            else left = ((pbin[i] >> (4u * N.axis)) & 15u) < N.split;
        }
        /* small inner nodes: moved by gb_small_kernel */
    } else if (i < n) pnode_out[i] = -1;
    const int32_t n0 = __shfl_sync(0xffffffffu, node, 0);
    const bool uniform = __all_sync(0xffffffffu, big && node == n0);
    const float cx = (P.lo.x + P.hi.x) * 0.5f, cy = (P.lo.y + P.hi.y) * 0.5f, cz = (P.lo.z + P.hi.z) * 0.5f;
    const uint32_t v[12] = { gb_enc(P.lo.x), gb_enc(P.lo.y), gb_enc(P.lo.z), gb_enc(P.hi.x), gb_enc(P.hi.y), gb_enc(P.hi.z),
                             gb_enc(cx), gb_enc(cy), gb_enc(cz), gb_enc(cx), gb_enc(cy), gb_enc(cz) };
    if (uniform) {
        GbNode &N = nodes[node];
        const unsigned mL = __ballot_sync(0xffffffffu, left), mR = ~mL;
        const unsigned m = left ? mL : mR;
        uint32_t r[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) r[k] = ((k / 3) & 1) ? __reduce_max_sync(m, v[k]) : __reduce_min_sync(m, v[k]);
        uint32_t base = 0;
        const bool leader = lane == (unsigned)__ffs(m) - 1u;
        if (leader) {
            base = left ? atomicAdd(&N.lcur, (uint32_t)__popc(mL)) : atomicAdd(&N.rcur, (uint32_t)__popc(mR));
            GbNode &C = nodes[left ? N.left : N.right];
#pragma unroll
            for (int k = 0; k < 3; ++k) { atomicMin(&C.blo[k], r[k]); atomicMax(&C.bhi[k], r[3 + k]); atomicMin(&C.clo[k], r[6 + k]); atomicMax(&C.chi[k], r[9 + k]); }
        }
        const uint32_t baseL = __shfl_sync(0xffffffffu, base, mL ? __ffs(mL) - 1 : 0), baseR = __shfl_sync(0xffffffffu, base, mR ? __ffs(mR) - 1 : 0);
        const uint32_t at = left ? N.first + baseL + (uint32_t)__popc(mL & ((1u << lane) - 1u)) : N.first + N.lcount + baseR + (uint32_t)__popc(mR & ((1u << lane) - 1u));
        prim_out[at] = P;
        pnode_out[at] = left ? N.left : N.right;
    } else if (big) {
        GbNode &N = nodes[node];
        GbNode &C = nodes[left ? N.left : N.right];
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&C.blo[k], v[k]); atomicMax(&C.bhi[k], v[3 + k]); atomicMin(&C.clo[k], v[6 + k]); atomicMax(&C.chi[k], v[9 + k]); }
        const uint32_t at = left ? N.first + atomicAdd(&N.lcur, 1u) : N.first + N.lcount + atomicAdd(&N.rcur, 1u);
        prim_out[at] = P;
        pnode_out[at] = left ? N.left : N.right;
    }
}

__global__ void gb_advance_kernel(GbState *S, uint32_t level)
{
    if (blockIdx.x || threadIdx.x) return;
    S->lvl_begin[level + 2] = S->node_count;
    for (uint32_t l = level + 3; l < GB_MAX_LEVELS + 2; ++l) S->lvl_begin[l] = S->node_count;
    S->slot_count[level & 1u] = 0;
}

/* ---- after the last level ------------------------------------------------------------------------------------------------ */
__global__ void gb_sizes_kernel(GbNode *nodes, const GbState *S, uint32_t level)
{
    const uint32_t lb = S->lvl_begin[level], le = S->lvl_begin[level + 1];
    for (uint32_t j = lb + blockIdx.x * blockDim.x + threadIdx.x; j < le; j += gridDim.x * blockDim.x) {
        GbNode &N = nodes[j];
        if (N.left < 0) { N.inner = 0; N.inner_even = 0; N.height = 0; continue; }
        const GbNode &L = nodes[N.left], &R = nodes[N.right];
        N.inner = 1 + L.inner + R.inner;
        N.inner_even = 1 + (L.inner - L.inner_even) + (R.inner - R.inner_even);
        N.height = 1 + max(L.height, R.height);
    }
}

__global__ void gb_slots_kernel(GbNode *nodes, const GbState *S, uint32_t level)
{
    const uint32_t lb = S->lvl_begin[level], le = S->lvl_begin[level + 1];
    for (uint32_t j = lb + blockIdx.x * blockDim.x + threadIdx.x; j < le; j += gridDim.x * blockDim.x) {
        GbNode &N = nodes[j];
        if (N.left < 0) continue;
        if (level == 0) { N.out_slot = 0; N.out_slot4 = 0; }
        GbNode &L = nodes[N.left], &R = nodes[N.right];
        L.out_slot = N.out_slot + 1;
        R.out_slot = N.out_slot + 1 + L.inner;
        if ((level & 1u) == 0u) {                                     /* even depth: this node becomes a 4-wide node; its inner grand-children get theirs */
            int32_t next = N.out_slot4 + 1;
            const int32_t kids[2] = { N.left, N.right };
            for (int k = 0; k < 2; ++k) {
                GbNode &C = nodes[kids[k]];
                if (C.left < 0) continue;
                GbNode &G0 = nodes[C.left], &G1 = nodes[C.right];
                if (G0.left >= 0) { G0.out_slot4 = next; next += G0.inner_even; }
                if (G1.left >= 0) { G1.out_slot4 = next; next += G1.inner_even; }
            }
        }
    }
}

__device__ __forceinline__ int32_t gb_leaf_code(uint32_t first, uint32_t count) { return ~(int32_t)((first << 3) | count); }

__global__ void gb_emit_kernel(const GbNode *__restrict__ nodes, uint32_t n_nodes, BvhNode *__restrict__ out2, Bvh4Node *__restrict__ out4, uint32_t *__restrict__ order)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_nodes) return;
    const GbNode &N = nodes[j];
    if (N.left < 0) {                                                  /* leaf: ascending original index */
        uint32_t *o = order + N.first;
        for (uint32_t a = 1; a < N.count; ++a) {
            const uint32_t x = o[a];
            uint32_t b = a;
            while (b > 0 && o[b - 1] > x) { o[b] = o[b - 1]; --b; }
            o[b] = x;
        }
        return;
    }
    const GbNode &L = nodes[N.left], &R = nodes[N.right];
    BvhNode o;
    o.lo0x = gb_dec(L.blo[0]); o.lo0y = gb_dec(L.blo[1]); o.lo0z = gb_dec(L.blo[2]);
    o.hi0x = gb_dec(L.bhi[0]); o.hi0y = gb_dec(L.bhi[1]); o.hi0z = gb_dec(L.bhi[2]);
    o.lo1x = gb_dec(R.blo[0]); o.lo1y = gb_dec(R.blo[1]); o.lo1z = gb_dec(R.blo[2]);
    o.hi1x = gb_dec(R.bhi[0]); o.hi1y = gb_dec(R.bhi[1]); o.hi1z = gb_dec(R.bhi[2]);
    o.c0 = L.left < 0 ? gb_leaf_code(L.first, L.count) : L.out_slot;
    o.c1 = R.left < 0 ? gb_leaf_code(R.first, R.count) : R.out_slot;
    o.pad0 = o.pad1 = 0;
    out2[N.out_slot] = o;
    if (N.depth & 1u) return;
    Bvh4Node q;
    int ng = 0;
    const int32_t kids[2] = { N.left, N.right };
    int32_t g[4];
    for (int k = 0; k < 2; ++k) {
        const GbNode &C = nodes[kids[k]];
        if (C.left >= 0) { g[ng++] = C.left; g[ng++] = C.right; } else g[ng++] = kids[k];
    }
    for (int s = 0; s < 4; ++s) {
        q.pad[s] = 0;
        if (s >= ng) {
            q.lox[s] = q.loy[s] = q.loz[s] = 3.402823466e+38f; q.hix[s] = q.hiy[s] = q.hiz[s] = -3.402823466e+38f;
            q.c[s] = BVH4_EMPTY;
            continue;
        }
        const GbNode &G = nodes[g[s]];
        q.lox[s] = gb_dec(G.blo[0]); q.loy[s] = gb_dec(G.blo[1]); q.loz[s] = gb_dec(G.blo[2]);
        q.hix[s] = gb_dec(G.bhi[0]); q.hiy[s] = gb_dec(G.bhi[1]); q.hiz[s] = gb_dec(G.bhi[2]);
        q.c[s] = G.left >= 0 ? G.out_slot4 : gb_leaf_code(G.first, G.count);
    }
    out4[N.out_slot4] = q;
}

/* ---- host driver ----------------------------------------------------------------------------------------------------------- */
#define GB_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, errlen, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto fail; } } while (0)

static int gb_build_once(cudaStream_t st, const float *d_tris9, uint32_t n, int leaf_max, int num_sms, uint32_t sah_depth_max,
                         LbDeviceBvh *out, char *err, size_t errlen)
{
    GbPrim *prim[2] = { nullptr, nullptr };
    int32_t *pnode[2] = { nullptr, nullptr };
    uint16_t *pbin = nullptr;
    GbNode *nodes = nullptr;
    GbState *S = nullptr;
    GbBin *bins[2] = { nullptr, nullptr };
    GbState hs;
    GbNode root;
    const uint32_t max_slots = n / (GB_SMALL + 1u) + 2u;              /* big nodes of one level are disjoint and hold > 32 prims each */
    const unsigned gp = grid_for(n, 256), gn = (unsigned)num_sms * 8u;
    uint32_t levels = 0;
    int cur = 0;
    out->nodes = nullptr; out->nodes4 = nullptr; out->order = nullptr; out->n_nodes = out->n_nodes4 = 0; out->height = 0; out->launches = 0;

    GB_CU(lb_malloc(&prim[0], (size_t)n * sizeof(GbPrim))); GB_CU(lb_malloc(&prim[1], (size_t)n * sizeof(GbPrim)));
    GB_CU(lb_malloc(&pnode[0], (size_t)n * 4)); GB_CU(lb_malloc(&pnode[1], (size_t)n * 4));
    GB_CU(lb_malloc(&pbin, (size_t)n * 2));
    GB_CU(lb_malloc(&nodes, (size_t)(2ull * n + 2) * sizeof(GbNode)));
    GB_CU(lb_malloc(&S, sizeof(GbState)));
    GB_CU(lb_malloc(&bins[0], (size_t)max_slots * 3 * GB_NB * sizeof(GbBin))); GB_CU(lb_malloc(&bins[1], (size_t)max_slots * 3 * GB_NB * sizeof(GbBin)));
    GB_CU(lb_malloc(&out->order, (size_t)n * 4));
    for (int b = 0; b < 2; ++b) {
        const size_t nb = (size_t)max_slots * 3 * GB_NB;
        gb_fill_bins_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(bins[b], nb);
    }
    gb_setup_kernel<<<1, 1, 0, st>>>(nodes, S, n);
    gb_init_kernel<<<gp, 256, 0, st>>>(d_tris9, n, prim[0], pnode[0], nodes);
    GB_CU(cudaMemsetAsync(pnode[1], 0xff, (size_t)n * 4, st));
    out->launches += 4;
    for (uint32_t level = 0; level < GB_MAX_LEVELS; ++level) {
        gb_bin_kernel<<<gp, 256, 0, st>>>(prim[cur], pnode[cur], n, nodes, bins[level & 1u], pbin, sah_depth_max, (uint32_t)leaf_max);
        gb_split_kernel<<<gn, 256, 0, st>>>(nodes, S, level, bins[level & 1u], sah_depth_max, (uint32_t)leaf_max, max_slots);
        gb_small_kernel<<<gn, GB_SMALL_WARPS * 32, 0, st>>>(prim[cur], prim[cur ^ 1], pnode[cur ^ 1], nodes, S, level, sah_depth_max, (uint32_t)leaf_max);
        gb_partition_kernel<<<gp, 256, 0, st>>>(prim[cur], pnode[cur], prim[cur ^ 1], pnode[cur ^ 1], n, nodes, pbin, out->order, (uint32_t)leaf_max);
        gb_advance_kernel<<<1, 1, 0, st>>>(S, level);
        out->launches += 5;
        cur ^= 1;
        levels = level + 1;
        if (level >= 15 && (level & 3u) == 3u) {                       /* every fourth level from the 16th: has the tree stopped growing? */
            GB_CU(cudaMemcpyAsync(&hs, S, sizeof(hs), cudaMemcpyDeviceToHost, st));
            GB_CU(cudaStreamSynchronize(st));
            if (hs.lvl_begin[level + 1] == hs.lvl_begin[level + 2]) break;
        }
    }
    GB_CU(cudaGetLastError());
    GB_CU(cudaMemcpyAsync(&hs, S, sizeof(hs), cudaMemcpyDeviceToHost, st));
    GB_CU(cudaStreamSynchronize(st));
    if (hs.lvl_begin[levels] != hs.lvl_begin[levels + 1]) { snprintf(err, errlen, "device BVH build: tree still growing after %u levels", levels); goto fail; }
    while (levels > 1 && hs.lvl_begin[levels - 1] == hs.lvl_begin[levels]) --levels;      /* levels = number of non-empty levels */
    for (int l = (int)levels - 1; l >= 0; --l) { gb_sizes_kernel<<<gn, 256, 0, st>>>(nodes, S, (uint32_t)l); ++out->launches; }
    for (uint32_t l = 0; l < levels; ++l) { gb_slots_kernel<<<gn, 256, 0, st>>>(nodes, S, l); ++out->launches; }
    GB_CU(cudaMemcpyAsync(&root, nodes, sizeof(root), cudaMemcpyDeviceToHost, st));
    GB_CU(cudaStreamSynchronize(st));
    if (root.left < 0 || root.inner < 1) { snprintf(err, errlen, "device BVH build: root was not split (%u triangles)", n); goto fail; }
    out->n_nodes = (uint32_t)root.inner; out->n_nodes4 = (uint32_t)root.inner_even; out->height = root.height;
    GB_CU(lb_malloc(&out->nodes, (size_t)out->n_nodes * sizeof(BvhNode)));
    GB_CU(lb_malloc(&out->nodes4, (size_t)out->n_nodes4 * sizeof(Bvh4Node)));
    gb_emit_kernel<<<grid_for(hs.node_count, 256), 256, 0, st>>>(nodes, hs.node_count, out->nodes, out->nodes4, out->order);
    ++out->launches;
    GB_CU(cudaGetLastError());
    GB_CU(cudaStreamSynchronize(st));
    lb_free(prim[0]); lb_free(prim[1]); lb_free(pnode[0]); lb_free(pnode[1]); lb_free(pbin); lb_free(nodes); lb_free(S); lb_free(bins[0]); lb_free(bins[1]);
    return 0;
fail:
    lb_free(prim[0]); lb_free(prim[1]); lb_free(pnode[0]); lb_free(pnode[1]); lb_free(pbin); lb_free(nodes); lb_free(S); lb_free(bins[0]); lb_free(bins[1]);
    lb_free(out->nodes); lb_free(out->nodes4); lb_free(out->order);
    out->nodes = nullptr; out->nodes4 = nullptr; out->order = nullptr;
    return 1;
}

int lb_build_bvh_device(cudaStream_t st, const float *d_tris9, uint32_t n, int leaf_max, int num_sms, LbDeviceBvh *out, char *err, size_t errlen)
{
    if (leaf_max < 1) leaf_max = 1;
    if (leaf_max > 7) leaf_max = 7;
    if (n <= (uint32_t)leaf_max) { snprintf(err, errlen, "device BVH build needs more than leaf_max triangles"); return 1; }
    if (n > 0x0fffffffu) { snprintf(err, errlen, "too many triangles for the leaf code (first << 3)"); return 1; }
    uint32_t sah_depth_max = 32;
    if (const char *e = getenv("LTR_BVH_SAH_DEPTH")) sah_depth_max = (uint32_t)atoi(e);        /* tests: force the position splits */
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (gb_build_once(st, d_tris9, n, leaf_max, num_sms, sah_depth_max, out, err, errlen)) return 1;
        if (out->height <= GB_HEIGHT_MAX) return 0;
        /* degenerate geometry drove the SAH into a long chain: rebuild with position splits from higher up (height <= limit + log2 n) */
        lb_free(out->nodes); lb_free(out->nodes4); lb_free(out->order);
        out->nodes = nullptr; out->nodes4 = nullptr; out->order = nullptr;
        sah_depth_max = attempt == 0 ? 8u : 0u;
    }
    snprintf(err, errlen, "device BVH build: height %d exceeds the traversal stack", out->height);
    return 1;
}
