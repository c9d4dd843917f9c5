/*
 * gpu_internal.cuh -- device-side context, error plumbing and the BVH traversal routines shared by
 * the stage kernels.  sm_100a only; compiled with -fmad=false (parity, see vmath.h).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "gpu.h"
#include "bvh_entry.h"

#ifndef LB_VIS_SKIPEMPTY
#define LB_VIS_SKIPEMPTY 1                /* 1 = the any-hit node loop branches around the two last slots of a 4-wide node when both are unused (209.9 -> 208.6 ms on config 4) */
#endif
#ifndef LB_VIS_LEAFQ
#define LB_VIS_LEAFQ 1                    /* 1 = the any-hit walk of the 4-wide tree queues leaf codes instead of triangle indices (bvh4_anyhit_core) */
#endif
#ifndef LB_VIS_SMEMNODES
#define LB_VIS_SMEMNODES 0                /* 1 = the records of a chunk's entry nodes are copied to shared memory and the walk's first visit of each reads them there (north_star's "shared-memory node caching"): A/B variant */
#endif
#define LB_SM_TAG 0x40000000
#ifndef LB_VIS_Q8
#define LB_VIS_Q8 0                       /* 1 = the visibility walk reads 64-byte quantised nodes (Bvh4QNode): A/B variant, see profiles/r02_ab_runs.md */
#endif
#define LB_BLOCK 128                      /* traversal kernels: 4 warps per CTA, many CTAs per SM */
#define LB_PAD 1024                       /* slack elements on per-lumel arrays so shards can be padded to equal size */

struct DevBuf {                           /* growable device allocation */
    void *p = nullptr;
    size_t cap = 0;
};

struct ltrgpu_Ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_span0 = nullptr, ev_span1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    char err[512] = {0};
    void *stage_buf[2] = { nullptr, nullptr };    /* pinned staging for pageable uploads (lb_upload_staged) */
    cudaEvent_t stage_ev[2] = { nullptr, nullptr };

    ltrgpu_Params params;
    /* ---- uploaded scene ---- */
    uint32_t n_inst = 0, n_verts = 0, n_rtris = 0, n_rnodes = 0, n_ritems = 0, n_rtree_tris = 0;
    uint32_t n_bvh_nodes = 0, n_tris = 0, n_lights = 0, n_probes = 0;
    int bvh_height = 0;                       /* inner levels of the scene BVH (validated against BVH_STACK at upload) */
    float bvh_build_ms = 0.f;                 /* device build: CUDA-event time of the builder, 0 for a host-built tree */
    uint64_t n_texels = 0;
    ltrgpu_Inst *d_inst = nullptr;
    ltrgpu_Inst *h_inst = nullptr;
    V3 *d_wpos = nullptr, *d_wnrm = nullptr;
    float2 *d_vtex = nullptr, *d_ltex = nullptr;
    ltrgpu_RasterTri *d_rtris = nullptr;
    RefNode *d_rnodes = nullptr;
    int32_t *d_ritems = nullptr;
    float *d_rtree_tris = nullptr;
    float4 *d_rtree_boxes = nullptr;          /* 2 float4 per reference-order triangle: its box (conservative pre-test) */
    PreparedTri *d_rtree_ptris = nullptr;     /* the same triangles with the point-query terms precomputed (lumel_fix_kernel) */
    BvhNode *d_bvh = nullptr;
    Bvh4Node *d_bvh4 = nullptr;               /* 4-wide collapse of d_bvh for the any-hit walks */
    Bvh4QNode *d_bvh4q = nullptr;             /* LB_VIS_Q8 builds only: the same nodes quantised to 64 bytes */
    uint32_t n_bvh4_nodes = 0;
    PreparedTri *d_ptris = nullptr;
    RayTri *d_raytris = nullptr;
    /* flat BVH over the triangles of ALL instance trees (lumel_classify_kernel): aliases the scene BVH when every instance
     * casts shadows, a second device-built tree otherwise, NULL when neither exists */
    BvhNode *d_lbvh = nullptr;
    PreparedTri *d_lbvh_ptris = nullptr;
    bool lbvh_owned = false;
    uint32_t *d_tri_orig = nullptr;
    ltrgpu_Light *d_lights = nullptr;
    ltrgpu_Light *h_lights = nullptr;
    uint8_t *d_light_inst = nullptr;
    V3 *d_probe_pos = nullptr, *d_probe_nrm = nullptr;
    float *d_ao_cos = nullptr, *d_ao_sin = nullptr;
    float *d_blur_kernel = nullptr;
    int blur_ext = 0;

    /* ---- lumels (global array: probes first, then instances in order) ---- */
    uint64_t n_lumels = 0;
    uint64_t *h_inst_lumel_off = nullptr;     /* n_inst+1 */
    uint32_t *d_texkey = nullptr;             /* winner key per texel */
    uint32_t *d_texidx = nullptr;             /* exclusive scan of lumel flags */
    float4 *d_lpos = nullptr, *d_lnrm = nullptr, *d_lrad = nullptr, *d_lrgb = nullptr;
    float4 *d_lnmap = nullptr;                /* normal/focus per lumel (normal map mode) */
    uint32_t *d_lloc = nullptr, *d_linst = nullptr;

    /* ---- shard ---- */
    uint64_t sh_begin = 0, sh_end = 0;
    int rank = 0, world = 1;
    ltrgpu_allgather_fn allgather = nullptr;
    ltrgpu_gatherv_fn gatherv = nullptr;
    ltrgpu_alltoallv_fn alltoallv = nullptr;
    void *allgather_user = nullptr;

    /* ---- direct light ---- */
    uint64_t fvis_tab_base = 0; uint32_t fvis_tab_n = 0;   /* the factor table covers lumels [base, base + n) */
    float *d_fvis = nullptr;                  /* [n_lights][local lumels] shadow factors */
    unsigned long long *d_smask = nullptr;    /* sampled mode: [n_lights][local lumels] blocked-sample bit masks */
    float4 *d_light_samples = nullptr;        /* sampled mode: per (light, sample) table */
    uint2 *d_active = nullptr;                /* (local lumel, light) pairs to march */
    uint32_t *d_active_count = nullptr;

    /* ---- radiosity ---- */
    uint64_t rad_rows = 0, rad_links = 0, rad_k0 = 0, rad_n = 0;
    uint32_t *d_rad_sidx = nullptr;           /* Morton order: sorted position -> original lumel index */
    uint64_t *d_rad_rowoff = nullptr;
    uint32_t *d_rad_other = nullptr;
    float *d_rad_factor = nullptr;

    /* ---- finalize ---- */
    float *d_image = nullptr, *d_image_tmp = nullptr;   /* concatenated rgb images */
    unsigned char *d_mask = nullptr, *d_mask_tmp = nullptr;
    float *d_normals = nullptr;               /* concatenated xyzf images (normal map mode) */
    uint64_t *h_out_off = nullptr;            /* per instance: offset (in texels) of the OUTPUT image */
    uint32_t *h_out_w = nullptr, *h_out_h = nullptr;
    float *d_out = nullptr;                   /* output images after optional ds2x */

    /* ---- early scene-BVH build (second stream; ltrgpu_upload_tris_early / ltrgpu_build_bvh_early) ---- */
    uint32_t early_tris = 0;                  /* > 0: d_rtree_tris already holds that many triangles of this scene */
    bool early_bvh = false;                   /* the scene BVH over them has been built (d_bvh, d_bvh4, d_tri_orig, d_ptris, d_raytris) */
    cudaEvent_t ev_early = nullptr;           /* end of that build on the second stream */
    unsigned early_launches = 0;
    char early_err[256] = {0};

    /* ---- sample_fn batching (second stream; touched by the material thread only, see gpu.h) ---- */
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_lumels = nullptr, ev_req[2] = { nullptr, nullptr };
    ltrgpu_SampleReq *d_req[2] = { nullptr, nullptr };
    uint32_t req_cap[2] = { 0, 0 };
    char aux_err[256] = {0};
    unsigned long long aux_d2h_bytes = 0, aux_launches = 0;   /* added to the bake's counters by ltrgpu_get_counters */

    /* ---- counters ---- */
    unsigned long long *d_counters = nullptr; /* see CNT_* */
    ltrgpu_Counters host_counters;
};

enum { CNT_MARCHES = 0, CNT_DIST_QUERIES, CNT_AO_SEGMENTS, CNT_CORR_RAYS, CNT_RAD_PAIRS, CNT_RAD_SEGMENTS,
       CNT_RAD_LINKS, CNT_NODE_VISITS, CNT_TRI_TESTS, CNT_RAY_NODE_VISITS, CNT_RAY_TRI_TESTS, CNT_RAD_TILE_LOADS, CNT_SHADOW_RAYS, CNT_RAY_ENTRY_TESTS, CNT_COUNT };

#define CU_TRY(ctx, call)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call,    \
                     cudaGetErrorString(e_));                                                      \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define CU_LAUNCH_CHECK(ctx)                                                                       \
    do {                                                                                           \
        (ctx)->host_counters.kernel_launches++;                                                    \
        CU_TRY(ctx, cudaGetLastError());                                                           \
    } while (0)

/* Caching device allocator (gpu_scene.cu): a bake allocates and drops dozens of large buffers; going
 * to cudaMalloc/cudaFree each time costs tens of milliseconds and a device-wide sync per call, so
 * freed blocks are kept per device and reused by size class.  Trimmed by ltrgpu_destroy. */
cudaError_t lb_malloc(void **p, size_t bytes);
template <class T> static inline cudaError_t lb_malloc(T **p, size_t bytes) { return lb_malloc((void **)p, bytes); }
void lb_free(void *p);
void lb_trim(void);

template <class T> static inline int dev_alloc(ltrgpu_Ctx *ctx, T **p, size_t count)
{
    if (*p) { lb_free(*p); *p = nullptr; }
    if (count == 0) count = 1;
    CU_TRY(ctx, lb_malloc((void **)p, count * sizeof(T)));
    return 0;
}
int lb_upload_staged(ltrgpu_Ctx *ctx, void *dst, const void *src, size_t bytes);      /* gpu_scene.cu */
template <class T> static inline int dev_upload(ltrgpu_Ctx *ctx, T **p, const void *src, size_t count)
{
    if (dev_alloc(ctx, p, count)) return 1;
    if (count && src) {
        if (lb_upload_staged(ctx, *p, src, count * sizeof(T))) return 1;
        ctx->host_counters.h2d_bytes += count * sizeof(T);
    }
    return 0;
}
template <class T> static inline void dev_free(T **p) { if (*p) { lb_free(*p); *p = nullptr; } }

static inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

/* The flat scene BVH built on the device (gpu_bvh.cu): binary nodes, their 4-wide collapse and the triangle order, all in
 * lb_malloc'ed device memory owned by the caller.  n must exceed leaf_max (smaller scenes are one wrapped leaf: bvh.cpp). */
struct LbDeviceBvh {
    BvhNode *nodes; Bvh4Node *nodes4; uint32_t *order;
    uint32_t n_nodes, n_nodes4;
    int height;                 /* levels of inner nodes */
    unsigned launches;
};
int lb_build_bvh_device(cudaStream_t st, const float *d_tris9, uint32_t n, int leaf_max, int num_sms, LbDeviceBvh *out, char *err, size_t errlen);

/* ------------------------------------------------------------------------------------------
 * device helpers
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ V3 ld3(const float4 &v) { return mk3(v.x, v.y, v.z); }

/* libm-class functions: evaluated in double and rounded once, which reproduces glibc's (nearly
 * always correctly rounded) float results; see DESIGN.md "transcendentals". */
__device__ __forceinline__ float ref_powf(float x, float y)
{
    if (y == 1.0f) return x;
    if (y == 0.0f) return 1.0f;
    return (float)pow((double)x, (double)y);
}
__device__ __forceinline__ float ref_acosf(float x) { return (float)acos((double)x); }
__device__ __forceinline__ float ref_sinf(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float ref_cosf(float x) { return (float)cos((double)x); }

struct TravStats { unsigned nodes, tris, entries; };

__device__ __forceinline__ float box_dist2(V3 p, float lx, float ly, float lz, float hx, float hy, float hz)
{
    float dx = fmaxf(fmaxf(lx - p.x, p.x - hx), 0.f);
    float dy = fmaxf(fmaxf(ly - p.y, p.y - hy), 0.f);
    float dz = fmaxf(fmaxf(lz - p.z, p.z - hz), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ void load_prepared(const PreparedTri *src, PreparedTri &T)
{
    const float4 *s = reinterpret_cast<const float4 *>(src);
    float4 *d = reinterpret_cast<float4 *>(&T);
#pragma unroll
    for (int i = 0; i < 10; ++i) d[i] = __ldg(s + i);
}

/*
 * Nearest-triangle distance, clamped to `radius` (ref semantics: ltr_Scene::Distance,
 * lighter.cpp:150-188 = min(MAX_PENUMBRA_SIZE, min over triangles of PointTriangleDistance)).
 * Returns early with a value < stop_below as soon as one is found (the march only needs to know
 * that h < 0.001, lighter.cpp:200-201).  Pruning is conservative: a sub-tree is skipped only when
 * its box is farther than the best distance so far.
 */
template <int FLUSH = 3, bool HINT = false>
__device__ __forceinline__ float bvh_distance(const BvhNode *__restrict__ nodes, const PreparedTri *__restrict__ tris,
                                              V3 p, float radius, float stop_below, TravStats &ts, int *hint = nullptr)
{
    /* Two-phase walk: the node loop (nearer child first, sub-trees farther than the best distance so far pruned) only
     * QUEUES the triangles of the leaves it reaches; once FLUSH are pending, or the walk is over, they are evaluated
     * together.  ncu on the test-as-you-go version: half of the warp instructions were point/triangle evaluations
     * running with 5 of 32 lanes, because lanes reach their leaves at different iterations; evaluated in a batch the
     * lanes of a warp (neighbouring lumels at the same march step) do it side by side.  Pruning lags by at most
     * FLUSH triangles. */
    constexpr int TQ = FLUSH + 14;                   /* FLUSH - 1 pending + two leaves of up to 7 triangles */
    int   stack_n[BVH_STACK];
    float stack_d[BVH_STACK];
    int   tq[TQ];
    float tqd[TQ];
    int sp = 0, nq = 0;
    float best = radius;
    /* Pruning bound.  The reference's point/triangle distance (geom.h, lighter_math.cpp:875-913) subtracts dot products of
     * world coordinates, so its value carries an ABSOLUTE rounding error that grows with the coordinates (~1e-6 x |p|): two
     * triangles sharing the nearest edge report distances a few ulps apart and the reference, which looks at both, keeps
     * the smaller.  A box is therefore skipped only when it is farther than best + slack, slack = 1e-5 x (1 + |p|_inf) --
     * ten times that error (a relative slack on the square, the round-1 form, let the warm-started walk keep 2 of
     * 69.8 M config-4 factors one ulp above the brute-force minimum: profiles/r02_ab_runs.md). */
    const float slack = 1e-5f * (1.0f + fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z)));
#define LB_PRUNE2(B) (((B) + slack) * ((B) + slack))
    float best2 = LB_PRUNE2(best);
    /* HINT (the shadow march): *hint is the triangle that was nearest at the previous step of the same march (-1: none).
     * It is evaluated first, so the walk starts with the bound it usually ends with and only visits what is nearer than
     * that -- the minimum over the triangles within `radius` does not depend on the order they are looked at, so the
     * result is the same float.  On return *hint is the nearest triangle found (-1 when nothing is within radius). */
    int hint_in = -1, best_slot = -1;
    if (HINT) {
        hint_in = *hint;
        if (hint_in >= 0) {
            PreparedTri T;
            load_prepared(tris + hint_in, T);
            ts.tris++;
            const float d = point_tri_distance_prepared(p, T);
            if (d < best) {
                best = d; best2 = LB_PRUNE2(best); best_slot = hint_in;
                if (best < stop_below) return best;
            }
        }
    }
    int node = 0;
    for (;;) {
        while (node >= 0) {
            const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
            float4 a = __ldg(n4), b = __ldg(n4 + 1), c = __ldg(n4 + 2);
            int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 3));
            ts.nodes++;
            float d0 = box_dist2(p, a.x, a.y, a.z, a.w, b.x, b.y);
            float d1 = box_dist2(p, b.z, b.w, c.x, c.y, c.z, c.w);
            int c0 = k.x, c1 = k.y;
            if (d1 < d0) { float td = d0; d0 = d1; d1 = td; int tc = c0; c0 = c1; c1 = tc; }   /* c0 = nearer */
            int next = -1;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                int cc = side ? c1 : c0;
                float dd = side ? d1 : d0;
                if (dd > best2) continue;
                if (cc < 0) {
                    unsigned code = ~cc;
                    unsigned first = code >> 3, cnt = code & 7u;
                    for (unsigned t = 0; t < cnt; ++t) { tq[nq] = (int)(first + t); tqd[nq] = dd; ++nq; }
                } else if (next < 0) {
                    next = cc;
                } else {
                    stack_n[sp] = cc; stack_d[sp] = dd; ++sp;
                }
            }
            node = next;
            while (node < 0 && sp > 0) { --sp; if (stack_d[sp] <= best2) node = stack_n[sp]; }
            if (nq >= FLUSH) break;
        }
        while (nq > 0) {
            --nq;
            if (tqd[nq] > best2) continue;               /* the best distance has improved since this leaf was queued */
            if (HINT && tq[nq] == hint_in) continue;     /* evaluated before the walk */
            PreparedTri T;
            load_prepared(tris + tq[nq], T);
            ts.tris++;
            float d = point_tri_distance_prepared(p, T);
            if (d < best) {
                best = d;
                best2 = LB_PRUNE2(best);
                if (HINT) best_slot = tq[nq];
                if (best < stop_below) return best;
            }
        }
        if (node < 0) {
            /* the stack may still hold sub-trees that were skipped at pop time with an older bound: they stay skipped
             * (the bound only shrinks), so an empty `node` with nothing pending means the walk is complete */
            if (HINT) *hint = best_slot;
            return best;
        }
    }
#undef LB_PRUNE2
}

/*
 * The same nearest-distance query on the 4-wide tree (bvh.h Bvh4Node): one node read decides four boxes, so the chain of
 * dependent node loads a query waits for is half as long -- and the march is bound by exactly that wait (ncu: 29 % of its
 * stall samples sit on the two lines that consume a freshly loaded node).  The nearest inner child in range is walked next,
 * the others are pushed with their box distances; leaves are queued as leaf CODES and expanded in the evaluation phase.
 * Unused slots hold an inverted box of +-FLT_MAX: their box distance overflows to +inf and they are pruned like any far box.
 * The result is the minimum over the same set of triangles as on the binary tree (everything whose box is not farther than
 * best + slack), hence the same float.
 */
template <int LEAVES = 2, bool HINT = false>
__device__ __forceinline__ float bvh4_distance(const Bvh4Node *__restrict__ nodes, const PreparedTri *__restrict__ tris,
                                               V3 p, float radius, float stop_below, TravStats &ts, int *hint = nullptr)
{
    constexpr int TQ = LEAVES + 3;                   /* LEAVES - 1 pending + the four children of one node */
    int   stack_n[BVH_STACK];
    float stack_d[BVH_STACK];
    unsigned tq[TQ];
    float tqd[TQ];
    int sp = 0, nq = 0;
    float best = radius;
    const float slack = 1e-5f * (1.0f + fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z)));      /* see bvh_distance */
#define LB_PRUNE2(B) (((B) + slack) * ((B) + slack))
    float best2 = LB_PRUNE2(best);
    int hint_in = -1, best_slot = -1;
    if (HINT) {
        hint_in = *hint;
        if (hint_in >= 0) {
            PreparedTri T;
            load_prepared(tris + hint_in, T);
            ts.tris++;
            const float d = point_tri_distance_prepared(p, T);
            if (d < best) {
                best = d; best2 = LB_PRUNE2(best); best_slot = hint_in;
                if (best < stop_below) return best;
            }
        }
    }
    int node = 0;
    for (;;) {
        while (node >= 0) {
            const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
            const float4 lx = __ldg(n4), ly = __ldg(n4 + 1), lz = __ldg(n4 + 2), hx = __ldg(n4 + 3), hy = __ldg(n4 + 4), hz = __ldg(n4 + 5);
            const int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 6));
            ts.nodes += 2;
            int next = -1;
            float nextd = 0.f;
#define LB_D4_CHILD(LX, LY, LZ, HX, HY, HZ, C)                                                                    \
            {                                                                                                     \
                const float dd = box_dist2(p, (LX), (LY), (LZ), (HX), (HY), (HZ));                                \
                if (!(dd > best2)) {                                                                              \
                    if ((C) < 0) { tq[nq] = ~(unsigned)(C); tqd[nq] = dd; ++nq; }                                 \
                    else if (next < 0) { next = (C); nextd = dd; }                                                \
                    else if (dd < nextd) { stack_n[sp] = next; stack_d[sp] = nextd; ++sp; next = (C); nextd = dd; } \
                    else { stack_n[sp] = (C); stack_d[sp] = dd; ++sp; }                                           \
                }                                                                                                 \
            }
            LB_D4_CHILD(lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, k.x)
            LB_D4_CHILD(lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, k.y)
            LB_D4_CHILD(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, k.z)
            LB_D4_CHILD(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, k.w)
#undef LB_D4_CHILD
            node = next;
            while (node < 0 && sp > 0) { --sp; if (stack_d[sp] <= best2) node = stack_n[sp]; }
            if (nq >= LEAVES) break;
        }
        while (nq > 0) {
            --nq;
            if (tqd[nq] > best2) continue;               /* the best distance has improved since this leaf was queued */
            const unsigned code = tq[nq];
            for (unsigned t = 0, slot = code >> 3; t < (code & 7u); ++t, ++slot) {
                if (HINT && (int)slot == hint_in) continue;                /* evaluated before the walk */
                PreparedTri T;
                load_prepared(tris + slot, T);
                ts.tris++;
                const float d = point_tri_distance_prepared(p, T);
                if (d < best) {
                    best = d;
                    best2 = LB_PRUNE2(best);
                    if (HINT) best_slot = (int)slot;
                    if (best < stop_below) return best;
                }
            }
        }
        if (node < 0) {
            if (HINT) *hint = best_slot;
            return best;
        }
    }
#undef LB_PRUNE2
}

/*
 * ref: lighter.cpp:190-207 (CalcInvShadowFactor): sphere tracing from the lumel to the light, every step a nearest-distance
 * query capped at 2 (MAX_PENUMBRA_SIZE), steps capped at 1 (MAX_PENUMBRA_STEP).  Shared by direct_march_kernel and the test
 * entry point (ltrx_test_march).
 *
 * Measured and NOT adopted (round 2, profiles/r02_ab_runs.md "open-air windows"): after a capped answer a lane asked ONE box
 * question -- "is every leaf box farther than 2.01 from the bounding box of my next K positions?" -- and replayed the K
 * predictable steps (h = 2, t += 1) without walking the tree.  Bit-identical factors and step counts, 18 % fewer node
 * visits, and 3.2x SLOWER (config 4: 131.6 vs 41.4 ms, config 3: 291 vs 90): lanes that jump K steps ahead of their
 * neighbours leave the lock step in which the 32 marches of a warp read the same nodes, and the walks of a warp stop
 * sharing cache lines.  The cost of a march is its near-surface steps, not the open-air ones (2-3 node visits each).
 */
#ifndef LB_MARCH_BVH4
#define LB_MARCH_BVH4 0          /* 1 = the march's distance queries walk the 4-wide tree (bvh4_distance).  Same floats (device test + hashes), no gain:
                                  * config 4 march 37.8 -> 38.7 ms, config 3 69.8 -> 68.3 (profiles/r02_ab_runs.md): half the node reads, twice the box
                                  * arithmetic per read, and the nearer-child-first order of the binary walk prunes better.  Kept for A/B */
#endif
__device__ __forceinline__ float march_shadow(const BvhNode *__restrict__ bvh, const Bvh4Node *__restrict__ bvh4, const PreparedTri *__restrict__ tris,
                                              V3 from, V3 to, float k, unsigned &queries, TravStats &ts)
{
    V3 rd = norm3(to - from);
    float maxt = len3(to - from);
    float res = 1.0f;
#ifndef LB_MARCH_HINT
#define LB_MARCH_HINT 1          /* warm start of every distance query with the previous step's nearest triangle (bvh_distance HINT) */
#endif
    int hint = -1;
    for (float t = 0.001f; t < maxt;) {
#if LB_MARCH_BVH4
        float h = bvh4_distance<2, LB_MARCH_HINT != 0>(bvh4, tris, from + rd * t, 2.0f, 0.001f, ts, &hint);
#else
        float h = bvh_distance<3, LB_MARCH_HINT != 0>(bvh, tris, from + rd * t, 2.0f, 0.001f, ts, &hint);
#endif
        ++queries;
        if (h < 0.001f) return 0.0f;
        res = fminr(res, h / fminr(t * k, 2.0f));
        h = fminr(h, 1.0f);
        t += h;
    }
    return res;
}

/* Segment set-up for BVH traversal: parametrised over [0,1] on l1 -> l2 so that box entry
 * distances compare directly with the hit parameter of seg_tri_prepared. */
struct SegRay { V3 o, d, inv; };

__device__ __forceinline__ SegRay make_seg(V3 l1, V3 l2)
{
    SegRay r;
    r.o = l1; r.d = l2 - l1;
    r.inv = mk3(r.d.x != 0 ? 1.0f / r.d.x : 0.f, r.d.y != 0 ? 1.0f / r.d.y : 0.f, r.d.z != 0 ? 1.0f / r.d.z : 0.f);
    return r;
}

/* entry parameter of the segment into the box, or +inf when it misses [0,tmax] (conservative) */
__device__ __forceinline__ float seg_box(const SegRay &r, float lx, float ly, float lz, float hx, float hy, float hz, float tmax)
{
    float t0 = 0.f, t1 = tmax;
    if (r.d.x != 0) { float a = (lx - r.o.x) * r.inv.x, b = (hx - r.o.x) * r.inv.x; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    else if (r.o.x < lx || r.o.x > hx) return INFINITY;
    if (r.d.y != 0) { float a = (ly - r.o.y) * r.inv.y, b = (hy - r.o.y) * r.inv.y; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    else if (r.o.y < ly || r.o.y > hy) return INFINITY;
    if (r.d.z != 0) { float a = (lz - r.o.z) * r.inv.z, b = (hz - r.o.z) * r.inv.z; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    else if (r.o.z < lz || r.o.z > hz) return INFINITY;
    /* widen by a few ulps so rounding in the slab products can never drop a touching box */
    return (t0 <= t1 * 1.0000005f + 1e-7f) ? t0 : INFINITY;
}

__device__ __forceinline__ void load_raytri(const RayTri *src, RayTri &T)
{
    const float4 *s = reinterpret_cast<const float4 *>(src);
    float4 *d = reinterpret_cast<float4 *>(&T);
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = __ldg(s + i);
}

/*
 * Segment query against the scene BVH.
 *   ANY = true : returns 1.0f-below value as soon as any triangle is hit (ref: VisibilityTest's
 *                any-hit, lighter.cpp:112-147 / lighter_math.cpp:804-832), else LB_NO_HIT
 *   ANY = false: closest-hit parameter (ref: lighter_math.cpp:835-871), ties -> lowest original
 *                triangle index; *hit_slot receives the BVH-order triangle slot or -1
 * The caller shortens the segment (SMALL_FLOAT at both ends) exactly as the reference does.
 */
template <bool ANY>
__device__ __forceinline__ float bvh_segment_core(const BvhNode *__restrict__ nodes, const RayTri *__restrict__ tris, const uint32_t *__restrict__ tri_orig,
                                                  const SegRay &r, int (&stack_n)[BVH_STACK], float (&stack_t)[BVH_STACK], int sp, int node,
                                                  int *hit_slot, TravStats &ts)
{
    float best = LB_NO_HIT;
    float tmax = 1.0f;
    int best_slot = -1;
    unsigned best_orig = 0xffffffffu;
    for (;;) {
        if (node < 0) {                                      /* next postponed node that can still hold a closer hit */
            for (;;) {
                if (sp == 0) { if (hit_slot) *hit_slot = best_slot; return best; }
                --sp;
                if (stack_t[sp] <= tmax) { node = stack_n[sp]; break; }
            }
        }
        const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
        float4 a = __ldg(n4), b = __ldg(n4 + 1), c = __ldg(n4 + 2);
        int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 3));
        ts.nodes++;
        float e0 = seg_box(r, a.x, a.y, a.z, a.w, b.x, b.y, tmax);
        float e1 = seg_box(r, b.z, b.w, c.x, c.y, c.z, c.w, tmax);
        int c0 = k.x, c1 = k.y;
        if (e1 < e0) { float te = e0; e0 = e1; e1 = te; int tc = c0; c0 = c1; c1 = tc; }
        int next = -1;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            int cc = side ? c1 : c0;
            float ee = side ? e1 : e0;
            if (ee > tmax) continue;                         /* also rejects INFINITY */
            if (cc < 0) {
                unsigned code = ~cc;
                unsigned first = code >> 3, cnt = code & 7u;
                for (unsigned t = 0; t < cnt; ++t) {
                    RayTri T;
                    load_raytri(tris + first + t, T);
                    ts.tris++;
                    float h = seg_tri_prepared(r.o, r.d, T);
                    if (ANY) {
                        if (h < 1.0f) return h;
                    } else if (h < LB_NO_HIT) {
                        unsigned orig = __ldg(tri_orig + first + t);
                        if (h < best || (h == best && orig < best_orig)) {
                            best = h; best_slot = (int)(first + t); best_orig = orig;
                            tmax = fminf(1.0f, best * 1.0000005f + 1e-7f);   /* keep exact ties reachable */
                        }
                    }
                }
            } else if (next < 0) {
                next = cc;
            } else {
                stack_n[sp] = cc; stack_t[sp] = ee; ++sp;
            }
        }
        node = next;
    }
}

template <bool ANY>
__device__ __forceinline__ float bvh_segment(const BvhNode *__restrict__ nodes, const RayTri *__restrict__ tris,
                                             const uint32_t *__restrict__ tri_orig, V3 l1, V3 l2, int *hit_slot, TravStats &ts)
{
    int stack_n[BVH_STACK];
    float stack_t[BVH_STACK];
    const SegRay r = make_seg(l1, l2);
    return bvh_segment_core<ANY>(nodes, tris, tri_orig, r, stack_n, stack_t, 0, 0, hit_slot, ts);
}

/* The same query started from the entry set of the segment's bundle (bvh_entry.h, searched on the BINARY tree): entry
 * boxes the segment misses are skipped, the others are walked nearest-box-last-pushed; the result does not depend on the
 * order (closest hit, ties -> lowest original triangle index). */
template <bool ANY>
__device__ __forceinline__ float bvh_segment_entries(const BvhNode *__restrict__ nodes, const RayTri *__restrict__ tris, const uint32_t *__restrict__ tri_orig,
                                                     const BvhEntrySet &E, V3 l1, V3 l2, int *hit_slot, TravStats &ts)
{
    int stack_n[BVH_STACK];
    float stack_t[BVH_STACK];
    int sp = 0;
    const SegRay r = make_seg(l1, l2);
    const int n = E.n;
    for (int i = 0; i < n; ++i) {
        const float e = seg_box(r, E.lox[i], E.loy[i], E.loz[i], E.hix[i], E.hiy[i], E.hiz[i], 1.0f);
        if (e <= 1.0f) { stack_n[sp] = E.node[i]; stack_t[sp] = e; ++sp; }
    }
    ts.entries += (unsigned)n;
    return bvh_segment_core<ANY>(nodes, tris, tri_orig, r, stack_n, stack_t, sp, -1, hit_slot, ts);
}

/*
 * Any-hit specialisation for the radiosity visibility rays (billions per bake, ~98 % of them misses,
 * and the kernel is instruction-issue bound -- ncu: IPC 3.3 of 4): no child ordering (a miss must
 * visit every overlapped node anyway), no entry-distance stack, branch-free handling of zero
 * direction components (huge finite inverse, upper faces tested against an origin one ulp down:
 * geom.h lb_slab_inv / lb_slab_origin_hi), absolute slack on the slab comparison.  Box tests may use any
 * conservative arithmetic; the triangle test keeps the reference's exact operand order.
 */
template <int FLUSH = 10>  /* postponed triangles that trigger the test phase: high for rays that mostly miss, low for rays that are often blocked */
__device__ __forceinline__ bool bvh_anyhit(const BvhNode *__restrict__ nodes, const RayTri *__restrict__ tris, V3 l1, V3 l2, TravStats &ts)
{
    /* Two-phase ("while-while") walk: the node loop only COLLECTS the triangles of the leaves it meets; they are
     * tested afterwards, when the lanes of the warp have all run out of nodes (or a lane's queue is nearly full).
     * ncu on the one-phase version: node tests ran with 28 of 32 lanes, but the triangle tests -- a quarter of the
     * instructions -- with 9, because lanes reach their leaves at different iterations.  A miss (98 % of the
     * radiosity rays) costs exactly the same work either way; a hit is found a little later. */
    constexpr int TQ = FLUSH + 14;                   /* FLUSH - 1 pending + two leaves of up to 7 triangles */
    int stack_n[BVH_STACK];
    int tq[TQ];
    int sp = 0, nq = 0;
    const V3 d = l2 - l1;
    const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
    const V3 l1h = mk3(lb_slab_origin_hi(l1.x, d.x), lb_slab_origin_hi(l1.y, d.y), lb_slab_origin_hi(l1.z, d.z));
    int node = 0;
    for (;;) {
        while (node >= 0) {
            const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
            const float4 a = __ldg(n4), b = __ldg(n4 + 1), c = __ldg(n4 + 2);
            const int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 3));
            ts.nodes++;
            bool hit0, hit1;
            {
                float x0 = (a.x - l1.x) * ix, x1 = (a.w - l1h.x) * ix, y0 = (a.y - l1.y) * iy, y1 = (b.x - l1h.y) * iy, z0 = (a.z - l1.z) * iz, z1 = (b.y - l1h.z) * iz;
                float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
                hit0 = t0 <= t1 + 2e-6f;
            }
            {
                float x0 = (b.z - l1.x) * ix, x1 = (c.y - l1h.x) * ix, y0 = (b.w - l1.y) * iy, y1 = (c.z - l1h.y) * iy, z0 = (c.x - l1.z) * iz, z1 = (c.w - l1h.z) * iz;
                float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
                hit1 = t0 <= t1 + 2e-6f;
            }
            int next = -1;
            if (hit0) {
                if (k.x < 0) { const unsigned code = ~k.x; for (unsigned t = 0, first = code >> 3; t < (code & 7u); ++t) tq[nq++] = (int)(first + t); }
                else next = k.x;
            }
            if (hit1) {
                if (k.y < 0) { const unsigned code = ~k.y; for (unsigned t = 0, first = code >> 3; t < (code & 7u); ++t) tq[nq++] = (int)(first + t); }
                else if (next < 0) next = k.y;
                else stack_n[sp++] = k.y;
            }
            node = next >= 0 ? next : (sp ? stack_n[--sp] : -1);
            if (nq >= FLUSH) break;                      /* enough postponed work: test it (a blocked ray stops here) */
        }
        while (nq) {
            RayTri T;
            load_raytri(tris + tq[--nq], T);
            ts.tris++;
            if (seg_tri_prepared(l1, d, T) < 1.0f) return true;
        }
        if (node < 0) return false;
    }
}

/*
 * The same any-hit walk on the 4-wide tree (bvh.h Bvh4Node): one iteration tests four boxes (six float4 loads for the
 * boxes, one for the child codes).  A visit is counted as two node units (128 bytes = two 64-byte binary nodes).
 */
/* The walk proper, with the leaf queue handed in: `tq` may already hold nq leaf codes (version-2 entry sets queue the leaf
 * entries a ray touches); TQN >= max(nq on entry, (FLUSH + 1) / 2 - 1) + 4. */
template <int FLUSH, int TQN>
__device__ __forceinline__ bool bvh4_anyhit_core_q(const Bvh4Node *__restrict__ nodes, const RayTri *__restrict__ tris, const V3 l1, const V3 l1h, const V3 d,
                                                   const float ix, const float iy, const float iz, int (&stack_n)[BVH_STACK], int sp, int node,
                                                   unsigned (&tq)[TQN], int nq, TravStats &ts, const Bvh4Node *sm_nodes = nullptr)
{
    /* The queue holds LEAF CODES (first << 3 | count), one store per leaf met; the triangles of a leaf are enumerated in the
     * test phase.  (The round-1 form stored every triangle index: a chain of up to seven store + branch steps per leaf in
     * the node loop.)  The test phase starts after (FLUSH + 1) / 2 queued leaves (leaves hold <= 2 triangles by default). */
    constexpr int LQ = (FLUSH + 1) / 2;
    {
    for (;;) {
        while (node >= 0) {
#if LB_VIS_SMEMNODES
            /* generic loads: an entry node (LB_SM_TAG + slot) comes from its shared-memory copy, every other node from global memory */
            const float4 *n4 = reinterpret_cast<const float4 *>(node >= LB_SM_TAG ? sm_nodes + (node - LB_SM_TAG) : nodes + node);
            const float4 lx = n4[0], ly = n4[1], lz = n4[2], hx = n4[3], hy = n4[4], hz = n4[5];
            const int4 k = *reinterpret_cast<const int4 *>(n4 + 6);
#else
            const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
            const float4 lx = __ldg(n4), ly = __ldg(n4 + 1), lz = __ldg(n4 + 2), hx = __ldg(n4 + 3), hy = __ldg(n4 + 4), hz = __ldg(n4 + 5);
            const int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 6));
#endif
            ts.nodes += 2;
            int next = -1;
#define LB_BVH4_CHILD(LX, LY, LZ, HX, HY, HZ, C)                                                                        \
            {                                                                                                           \
                const float x0 = ((LX) - l1.x) * ix, x1 = ((HX) - l1h.x) * ix, y0 = ((LY) - l1.y) * iy, y1 = ((HY) - l1h.y) * iy; \
                const float z0 = ((LZ) - l1.z) * iz, z1 = ((HZ) - l1h.z) * iz;                                          \
                const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));                 \
                const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));                 \
                if (t0 <= t1 + 2e-6f && (C) != BVH4_EMPTY) {                                                            \
                    if ((C) < 0) tq[nq++] = ~(unsigned)(C);                                                             \
                    else if (next < 0) next = (C);                                                                      \
                    else stack_n[sp++] = (C);                                                                           \
                }                                                                                                       \
            }
            LB_BVH4_CHILD(lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, k.x)
            LB_BVH4_CHILD(lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, k.y)
#if LB_VIS_SKIPEMPTY
            /* a node of two leaf children uses two slots: when every active lane is on such a node the warp skips half the box arithmetic */
            if (k.z != BVH4_EMPTY || k.w != BVH4_EMPTY) {
                LB_BVH4_CHILD(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, k.z)
                LB_BVH4_CHILD(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, k.w)
            }
#else
            LB_BVH4_CHILD(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, k.z)
            LB_BVH4_CHILD(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, k.w)
#endif
#undef LB_BVH4_CHILD
            node = next >= 0 ? next : (sp ? stack_n[--sp] : -1);
            if (nq >= LQ) break;
        }
        while (nq) {
            const unsigned code = tq[--nq];
            const RayTri *tp = tris + (code >> 3);
            for (unsigned t = code & 7u; t; --t, ++tp) {
                RayTri T;
                load_raytri(tp, T);
                ts.tris++;
                if (seg_tri_prepared<true>(l1, d, T) < 1.0f) return true;     /* scene-BVH triangles are "useful" by construction */
            }
        }
        if (node < 0) return false;
    }
    }
}

template <int FLUSH>
__device__ __forceinline__ bool bvh4_anyhit_core(const Bvh4Node *__restrict__ nodes, const RayTri *__restrict__ tris, const V3 l1, const V3 l1h, const V3 d,
                                                 const float ix, const float iy, const float iz, int (&stack_n)[BVH_STACK], int sp, int node, TravStats &ts,
                                                 const Bvh4Node *sm_nodes = nullptr /* LB_VIS_SMEMNODES: shared-memory copies of the entry nodes, addressed as LB_SM_TAG + slot */)
{
    /* Measured and NOT adopted (B200, config 4): picking the near / far plane of every slab by the sign of the direction
     * (bit-identical to min/max, 4 three-way min/max per child instead of 10 two-way) needs a separate address per
     * float4 of the node; the extra pointers cost 16-24 registers and the kernel is more sensitive to occupancy than to
     * those instructions: 265 ms vs 236 ms for this form at 56 registers. */
#if LB_VIS_LEAFQ
    constexpr int LQ = (FLUSH + 1) / 2;
    unsigned tq[LQ + 3];                             /* LQ - 1 pending + the four children of one node */
    return bvh4_anyhit_core_q<FLUSH>(nodes, tris, l1, l1h, d, ix, iy, iz, stack_n, sp, node, tq, 0, ts, sm_nodes);
#else
    constexpr int TQ = FLUSH + 28;                   /* FLUSH - 1 pending + four leaves of up to 7 triangles */
    int tq[TQ];
    int nq = 0;
    for (;;) {
        while (node >= 0) {
            const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
            const float4 lx = __ldg(n4), ly = __ldg(n4 + 1), lz = __ldg(n4 + 2), hx = __ldg(n4 + 3), hy = __ldg(n4 + 4), hz = __ldg(n4 + 5);
            const int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 6));
            ts.nodes += 2;
            int next = -1;
#define LB_BVH4_CHILD(LX, LY, LZ, HX, HY, HZ, C)                                                                        \
            {                                                                                                           \
                const float x0 = ((LX) - l1.x) * ix, x1 = ((HX) - l1h.x) * ix, y0 = ((LY) - l1.y) * iy, y1 = ((HY) - l1h.y) * iy; \
                const float z0 = ((LZ) - l1.z) * iz, z1 = ((HZ) - l1h.z) * iz;                                          \
                const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));                 \
                const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));                 \
                if (t0 <= t1 + 2e-6f && (C) != BVH4_EMPTY) {                                                            \
                    if ((C) < 0) { const unsigned code = ~(C); for (unsigned t = 0, first = code >> 3; t < (code & 7u); ++t) tq[nq++] = (int)(first + t); } \
                    else if (next < 0) next = (C);                                                                      \
                    else stack_n[sp++] = (C);                                                                           \
                }                                                                                                       \
            }
            LB_BVH4_CHILD(lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, k.x)
            LB_BVH4_CHILD(lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, k.y)
            LB_BVH4_CHILD(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, k.z)
            LB_BVH4_CHILD(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, k.w)
#undef LB_BVH4_CHILD
            node = next >= 0 ? next : (sp ? stack_n[--sp] : -1);
            if (nq >= FLUSH) break;
        }
        while (nq) {
            RayTri T;
            load_raytri(tris + tq[--nq], T);
            ts.tris++;
            if (seg_tri_prepared(l1, d, T) < 1.0f) return true;
        }
        if (node < 0) return false;
    }
#endif
}

template <int FLUSH = 10>
__device__ __forceinline__ bool bvh4_anyhit(const Bvh4Node *__restrict__ nodes, const RayTri *__restrict__ tris, V3 l1, V3 l2, TravStats &ts)
{
    int stack_n[BVH_STACK];
    const V3 d = l2 - l1;
    const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
    const V3 l1h = mk3(lb_slab_origin_hi(l1.x, d.x), lb_slab_origin_hi(l1.y, d.y), lb_slab_origin_hi(l1.z, d.z));
    return bvh4_anyhit_core<FLUSH>(nodes, tris, l1, l1h, d, ix, iy, iz, stack_n, 0, 0, ts);
}

/*
 * The same walk started from the entry set of the ray's bundle (bvh_entry.h) instead of the root: the ray tests the
 * entry boxes (one box each; the set lives in shared memory, every lane reads the same words) and walks the sub-trees it
 * touches.  ts.entries counts the entry boxes tested.
 */
template <int FLUSH = 10>
__device__ __forceinline__ bool bvh4_anyhit_entries(const Bvh4Node *__restrict__ nodes, const RayTri *__restrict__ tris, const BvhEntrySet &E,
                                                    V3 l1, V3 l2, TravStats &ts, const Bvh4Node *sm_nodes = nullptr)
{
    int stack_n[BVH_STACK];
    int sp = 0;
    const V3 d = l2 - l1;
    const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
    const V3 l1h = mk3(lb_slab_origin_hi(l1.x, d.x), lb_slab_origin_hi(l1.y, d.y), lb_slab_origin_hi(l1.z, d.z));
    const int n = E.n;
    for (int i = 0; i < n; ++i) {
        const float x0 = (E.lox[i] - l1.x) * ix, x1 = (E.hix[i] - l1h.x) * ix, y0 = (E.loy[i] - l1.y) * iy, y1 = (E.hiy[i] - l1h.y) * iy;
        const float z0 = (E.loz[i] - l1.z) * iz, z1 = (E.hiz[i] - l1h.z) * iz;
        const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
        const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
#if LB_VIS_SMEMNODES
        if (t0 <= t1 + 2e-6f) stack_n[sp++] = LB_SM_TAG + i;
#else
        if (t0 <= t1 + 2e-6f) stack_n[sp++] = E.node[i];
#endif
    }
    ts.entries += (unsigned)n;
    if (sp == 0) return false;
    const int node = stack_n[--sp];
    return bvh4_anyhit_core<FLUSH>(nodes, tris, l1, l1h, d, ix, iy, iz, stack_n, sp, node, ts, sm_nodes);
}

/*
 * The entry-set walk again, with the box tests in FMA form (rad_visibility_kernel; LB_VIS_FMA=0 selects the plain form for A/B).
 *
 *   t = (face - o) * inv  =  fma(face, inv, -(o * inv))          one instruction per face instead of two
 *
 * The product o * inv is rounded once per ray, so t carries an absolute error of ~2^-24 |o * inv| (1e-4 and more for short
 * rays far from the origin): this form can only be used as a CONSERVATIVE pre-filter -- its slack `eps` covers that error,
 * so it never rejects a box the plain test accepts.  What keeps the results bit-identical to the plain walk: an inner child
 * that passes too easily only costs a visit, and a LEAF child that passes is re-tested with the plain, exact-as-before
 * arithmetic (lb_slab_inv / lb_slab_origin_hi, slack 2e-6) before its triangles are queued -- the set of triangles tested,
 * and with it every hit / miss, is that of bvh4_anyhit_entries.  Per ray: 4.4 node visits x 4 children x 6 instructions and
 * 7 entry boxes x 6 instructions fewer, ~1.5 exact re-tests more.
 * The ray's origin and direction are only needed by the re-test and by the triangle tests: they wait in shared memory
 * (6 floats per thread, [component][thread]) instead of occupying six registers across the node loop.
 */
struct SlabF { float ix, iy, iz, nlx, nly, nlz, nhx, nhy, nhz, eps; };

__device__ __forceinline__ SlabF make_slabf(V3 l1, V3 d)
{
    const float BIG = 1099511627776.0f;            /* 2^40 on an axis the segment does not move along: products by a power of two are exact, so
                                                    * fma(face, BIG, -(o * BIG)) has the exact sign of face - o and is 0 on the face (geom.h) */
    SlabF S;
    S.ix = d.x != 0 ? __frcp_rn(d.x) : BIG; S.iy = d.y != 0 ? __frcp_rn(d.y) : BIG; S.iz = d.z != 0 ? __frcp_rn(d.z) : BIG;
    S.nlx = -(l1.x * S.ix); S.nly = -(l1.y * S.iy); S.nlz = -(l1.z * S.iz);
    S.nhx = -(lb_slab_origin_hi(l1.x, d.x) * S.ix); S.nhy = -(lb_slab_origin_hi(l1.y, d.y) * S.iy); S.nhz = -(lb_slab_origin_hi(l1.z, d.z) * S.iz);
    const float m = fmaxf(fmaxf(d.x != 0 ? fabsf(S.nlx) : 0.f, d.y != 0 ? fabsf(S.nly) : 0.f), d.z != 0 ? fabsf(S.nlz) : 0.f);
    S.eps = 2e-6f + 2.4e-7f * m;                   /* 4 x (2^-24 |o * inv|) on top of the plain test's own slack */
    return S;
}

__device__ __forceinline__ bool slabf_hit(const SlabF &S, float lx, float ly, float lz, float hx, float hy, float hz)
{
    const float x0 = __fmaf_rn(lx, S.ix, S.nlx), x1 = __fmaf_rn(hx, S.ix, S.nhx), y0 = __fmaf_rn(ly, S.iy, S.nly), y1 = __fmaf_rn(hy, S.iy, S.nhy);
    const float z0 = __fmaf_rn(lz, S.iz, S.nlz), z1 = __fmaf_rn(hz, S.iz, S.nhz);
    const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
    const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
    return t0 <= t1 + S.eps;
}

/* the plain test (what bvh4_anyhit_core does): the two origins come back from shared memory, the inverse direction is the
 * pre-filter's (__frcp_rn is the correctly rounded 1/d of the plain form; on a zero axis 2^40 instead of 1e30 decides alike:
 * the products are 0 or far outside [0,1] either way) */
__device__ __forceinline__ bool slab_exact_hit(const SlabF &S, const float *s_ray, float lx, float ly, float lz, float hx, float hy, float hz)
{
    const float x0 = (lx - s_ray[0]) * S.ix, x1 = (hx - s_ray[3 * LB_BLOCK]) * S.ix, y0 = (ly - s_ray[LB_BLOCK]) * S.iy, y1 = (hy - s_ray[4 * LB_BLOCK]) * S.iy;
    const float z0 = (lz - s_ray[2 * LB_BLOCK]) * S.iz, z1 = (hz - s_ray[5 * LB_BLOCK]) * S.iz;
    const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
    const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
    return t0 <= t1 + 2e-6f;
}

template <int FLUSH>
__device__ __forceinline__ bool bvh4_anyhit_entries_f(const Bvh4Node *__restrict__ nodes, const RayTri *__restrict__ tris, const BvhEntrySet &E,
                                                      const V3 l1, const V3 l2, float *s_ray /* this thread's column of [9][LB_BLOCK] */, TravStats &ts)
{
    constexpr int TQ = FLUSH + 28;
    int stack_n[BVH_STACK];
    int tq[TQ];
    int sp = 0, nq = 0;
    unsigned visits = 0;                             /* per ray; the long-lived counters in ts are touched once per ray, not once per node */
    {
        const V3 d = l2 - l1;
        s_ray[0] = l1.x; s_ray[LB_BLOCK] = l1.y; s_ray[2 * LB_BLOCK] = l1.z;
        s_ray[3 * LB_BLOCK] = lb_slab_origin_hi(l1.x, d.x); s_ray[4 * LB_BLOCK] = lb_slab_origin_hi(l1.y, d.y); s_ray[5 * LB_BLOCK] = lb_slab_origin_hi(l1.z, d.z);
        s_ray[6 * LB_BLOCK] = d.x; s_ray[7 * LB_BLOCK] = d.y; s_ray[8 * LB_BLOCK] = d.z;
    }
    const SlabF S = make_slabf(l1, l2 - l1);
    const int n = E.n;
    for (int i = 0; i < n; ++i)
        if (slabf_hit(S, E.lox[i], E.loy[i], E.loz[i], E.hix[i], E.hiy[i], E.hiz[i])) stack_n[sp++] = E.node[i];
    ts.entries += (unsigned)n;
    if (sp == 0) return false;
    int node = stack_n[--sp];
    for (;;) {
        while (node >= 0) {
            const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
            const float4 lx = __ldg(n4), ly = __ldg(n4 + 1), lz = __ldg(n4 + 2), hx = __ldg(n4 + 3), hy = __ldg(n4 + 4), hz = __ldg(n4 + 5);
            const int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 6));
            ++visits;
            int next = -1;
#define LB_BVH4_CHILD_F(LX, LY, LZ, HX, HY, HZ, C)                                                                      \
            if (slabf_hit(S, LX, LY, LZ, HX, HY, HZ) && (C) != BVH4_EMPTY) {                                            \
                if ((C) < 0) {                                                                                          \
                    if (slab_exact_hit(S, s_ray, LX, LY, LZ, HX, HY, HZ)) {                                             \
                        const unsigned code = ~(C); for (unsigned t = 0, first = code >> 3; t < (code & 7u); ++t) tq[nq++] = (int)(first + t); \
                    }                                                                                                   \
                } else if (next < 0) next = (C);                                                                        \
                else stack_n[sp++] = (C);                                                                               \
            }
            LB_BVH4_CHILD_F(lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, k.x)
            LB_BVH4_CHILD_F(lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, k.y)
            LB_BVH4_CHILD_F(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, k.z)
            LB_BVH4_CHILD_F(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, k.w)
#undef LB_BVH4_CHILD_F
            node = next >= 0 ? next : (sp ? stack_n[--sp] : -1);
            if (nq >= FLUSH) break;
        }
        if (nq) {
            const V3 o = mk3(s_ray[0], s_ray[LB_BLOCK], s_ray[2 * LB_BLOCK]), d = mk3(s_ray[6 * LB_BLOCK], s_ray[7 * LB_BLOCK], s_ray[8 * LB_BLOCK]);
            ts.tris += (unsigned)nq;
            while (nq) {
                RayTri T;
                load_raytri(tris + tq[--nq], T);
                if (seg_tri_prepared(o, d, T) < 1.0f) { ts.tris -= (unsigned)nq; ts.nodes += 2u * visits; return true; }
            }
        }
        if (node < 0) { ts.nodes += 2u * visits; return false; }
    }
}

/*
 * LB_VIS_FMA == 2: the FMA-form walk WITHOUT the exact re-test at the leaves and without the shared-memory copy of the ray.
 * Why no re-test is needed for identical results: a segment is blocked iff SOME triangle of the scene passes the exact
 * seg_tri_prepared test; box tests only prune.  slabf_hit never rejects a box the plain test accepts, so the triangles
 * tested here are a superset of the plain walk's and a subset of the scene: blocked / free come out the same (only the
 * triangle-test counter differs).  Leaves are queued as codes (see bvh4_anyhit_core, LB_VIS_LEAFQ).
 */
template <int FLUSH>
__device__ __forceinline__ bool bvh4_anyhit_entries_f2(const Bvh4Node *__restrict__ nodes, const RayTri *__restrict__ tris, const BvhEntrySet &E,
                                                       const V3 l1, const V3 l2, TravStats &ts)
{
    constexpr int LQ = (FLUSH + 1) / 2;
    int stack_n[BVH_STACK];
    unsigned tq[LQ + 3];
    int sp = 0, nq = 0;
    const V3 d = l2 - l1;
    const SlabF S = make_slabf(l1, d);
    const int n = E.n;
    for (int i = 0; i < n; ++i)
        if (slabf_hit(S, E.lox[i], E.loy[i], E.loz[i], E.hix[i], E.hiy[i], E.hiz[i])) stack_n[sp++] = E.node[i];
    ts.entries += (unsigned)n;
    if (sp == 0) return false;
    int node = stack_n[--sp];
    for (;;) {
        while (node >= 0) {
            const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
            const float4 lx = __ldg(n4), ly = __ldg(n4 + 1), lz = __ldg(n4 + 2), hx = __ldg(n4 + 3), hy = __ldg(n4 + 4), hz = __ldg(n4 + 5);
            const int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 6));
            ts.nodes += 2;
            int next = -1;
#define LB_BVH4_CHILD_F2(LX, LY, LZ, HX, HY, HZ, C)                                                                     \
            if (slabf_hit(S, LX, LY, LZ, HX, HY, HZ) && (C) != BVH4_EMPTY) {                                            \
                if ((C) < 0) tq[nq++] = ~(unsigned)(C);                                                                 \
                else if (next < 0) next = (C);                                                                          \
                else stack_n[sp++] = (C);                                                                               \
            }
            LB_BVH4_CHILD_F2(lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, k.x)
            LB_BVH4_CHILD_F2(lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, k.y)
            LB_BVH4_CHILD_F2(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, k.z)
            LB_BVH4_CHILD_F2(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, k.w)
#undef LB_BVH4_CHILD_F2
            node = next >= 0 ? next : (sp ? stack_n[--sp] : -1);
            if (nq >= LQ) break;
        }
        while (nq) {
            const unsigned code = tq[--nq];
            const RayTri *tp = tris + (code >> 3);
            for (unsigned t = code & 7u; t; --t, ++tp) {
                RayTri T;
                load_raytri(tp, T);
                ts.tris++;
                if (seg_tri_prepared(l1, d, T) < 1.0f) return true;
            }
        }
        if (node < 0) return false;
    }
}

/*
 * LB_VIS_Q8: the entry-set walk over 64-byte quantised nodes (bvh.h Bvh4QNode): 4 instead of 7 sixteen-byte loads per visit.
 * A child plane is corner + q * step; in the ray's parameter space
 *     t = (corner + q * step - o) * inv = q * (step * inv) + (corner - o) * inv
 * so per node and axis: sx = step * inv (exact: the step is a power of two), b = fma(corner, inv, -(o * inv)) for the lower
 * and the upper origin, and per plane ONE fma.  The byte q becomes a float without a conversion instruction: PRMT puts it
 * under the exponent of 2^23 (0x4B000000 | q = 8388608 + q exactly) and the constant is folded into b:
 *     t = fma(8388608 + q, sx, b - 8388608 * sx).
 * Rounding: b - 8388608 * sx is off by at most half an ulp of 2^23 * sx = sx / 2 -- half a quantisation step in t -- which
 * the builder's one-step widening covers; the remaining terms are those of the FMA form (make_slabf's eps).  Box tests only
 * prune, so as with LB_VIS_FMA == 2 blocked / free are those of the plain walk.
 */
__device__ __forceinline__ float q8_magic(uint32_t word, int byte)
{
    uint32_t r;
    switch (byte) {                                  /* selector nibbles pick: byte k of `word` as the low byte, 00 00 4B above it */
    case 0: asm("prmt.b32 %0, %1, %2, 0x7440;" : "=r"(r) : "r"(word), "r"(0x4B000000u)); break;
    case 1: asm("prmt.b32 %0, %1, %2, 0x7441;" : "=r"(r) : "r"(word), "r"(0x4B000000u)); break;
    case 2: asm("prmt.b32 %0, %1, %2, 0x7442;" : "=r"(r) : "r"(word), "r"(0x4B000000u)); break;
    default: asm("prmt.b32 %0, %1, %2, 0x7443;" : "=r"(r) : "r"(word), "r"(0x4B000000u)); break;
    }
    return __uint_as_float(r);
}

template <int FLUSH>
__device__ __forceinline__ bool bvh4q_anyhit_entries(const Bvh4QNode *__restrict__ nodes, const RayTri *__restrict__ tris, const BvhEntrySet &E,
                                                     const V3 l1, const V3 l2, TravStats &ts)
{
    constexpr int LQ = (FLUSH + 1) / 2;
    int stack_n[BVH_STACK];
    unsigned tq[LQ + 3];
    int sp = 0, nq = 0;
    const V3 d = l2 - l1;
    const SlabF S = make_slabf(l1, d);
    const int n = E.n;
    for (int i = 0; i < n; ++i)
        if (slabf_hit(S, E.lox[i], E.loy[i], E.loz[i], E.hix[i], E.hiy[i], E.hiz[i])) stack_n[sp++] = E.node[i];
    ts.entries += (unsigned)n;
    if (sp == 0) return false;
    int node = stack_n[--sp];
    for (;;) {
        while (node >= 0) {
            const uint4 *n4 = reinterpret_cast<const uint4 *>(nodes + node);
            const uint4 a = __ldg(n4), b = __ldg(n4 + 1), c = __ldg(n4 + 2), k4 = __ldg(n4 + 3);
            ts.nodes += 1;                           /* 64 bytes per visit */
            const float sx = __uint_as_float((a.w & 0xffu) << 23) * S.ix, sy = __uint_as_float((a.w & 0xff00u) << 15) * S.iy, sz = __uint_as_float((a.w & 0xff0000u) << 7) * S.iz;
            const float ox = __uint_as_float(a.x), oy = __uint_as_float(a.y), oz = __uint_as_float(a.z);
            const float blx = __fmaf_rn(-8388608.f, sx, __fmaf_rn(ox, S.ix, S.nlx)), bhx = __fmaf_rn(-8388608.f, sx, __fmaf_rn(ox, S.ix, S.nhx));
            const float bly = __fmaf_rn(-8388608.f, sy, __fmaf_rn(oy, S.iy, S.nly)), bhy = __fmaf_rn(-8388608.f, sy, __fmaf_rn(oy, S.iy, S.nhy));
            const float blz = __fmaf_rn(-8388608.f, sz, __fmaf_rn(oz, S.iz, S.nlz)), bhz = __fmaf_rn(-8388608.f, sz, __fmaf_rn(oz, S.iz, S.nhz));
            int next = -1;
#define LB_BVH4Q_CHILD(B, C)                                                                                            \
            {                                                                                                           \
                const float x0 = __fmaf_rn(q8_magic(b.x, B), sx, blx), x1 = __fmaf_rn(q8_magic(b.w, B), sx, bhx);        \
                const float y0 = __fmaf_rn(q8_magic(b.y, B), sy, bly), y1 = __fmaf_rn(q8_magic(c.x, B), sy, bhy);        \
                const float z0 = __fmaf_rn(q8_magic(b.z, B), sz, blz), z1 = __fmaf_rn(q8_magic(c.y, B), sz, bhz);        \
                const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));                 \
                const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));                 \
                if (t0 <= t1 + S.eps && (C) != BVH4_EMPTY) {                                                            \
                    if ((C) < 0) tq[nq++] = ~(unsigned)(C);                                                             \
                    else if (next < 0) next = (C);                                                                      \
                    else stack_n[sp++] = (C);                                                                           \
                }                                                                                                       \
            }
            LB_BVH4Q_CHILD(0, (int)c.z)
            LB_BVH4Q_CHILD(1, (int)c.w)
            LB_BVH4Q_CHILD(2, (int)k4.x)
            LB_BVH4Q_CHILD(3, (int)k4.y)
#undef LB_BVH4Q_CHILD
            node = next >= 0 ? next : (sp ? stack_n[--sp] : -1);
            if (nq >= LQ) break;
        }
        while (nq) {
            const unsigned code = tq[--nq];
            const RayTri *tp = tris + (code >> 3);
            for (unsigned t = code & 7u; t; --t, ++tp) {
                RayTri T;
                load_raytri(tp, T);
                ts.tris++;
                if (seg_tri_prepared(l1, d, T) < 1.0f) return true;
            }
        }
        if (node < 0) return false;
    }
}

/* lane index of the (k + 1)-th set bit of a mask whose set bits all lie in bits 0..3 (the four children of a node):
 * a handful of integer instructions instead of __fns, whose software loop was 3 % of the visibility kernel's instructions (ncu) */
__device__ __forceinline__ int nth_set_bit4(unsigned m, int k)
{
    const unsigned m1 = m & (m - 1u), m2 = m1 & (m1 - 1u), m3 = m2 & (m2 - 1u);
    const unsigned pick = k == 0 ? m : k == 1 ? m1 : k == 2 ? m2 : m3;
    return __ffs(pick) - 1;
}

/*
 * Warp-cooperative form of bvh_entry_search_t (bvh_entry.h): the same frontier, the same picks and the same slot
 * assignment -- hence the same entry set as the scalar host model -- but entry i lives in the registers of lane i and the
 * children of the picked node are tested one per lane, so an iteration is a few dozen warp instructions instead of a
 * one-lane loop (ncu on the scalar version: 10 % of the visibility kernel's time went into lane 0's search).
 * All 32 lanes must call it; the bundle box must already be padded and identical on every lane.  The result is written
 * to E (shared memory) and is visible to the warp on return.
 */
template <class A>
__device__ __forceinline__ void bvh_entry_search_warp(const typename A::Node *__restrict__ nodes, float qlx, float qly, float qlz, float qhx, float qhy, float qhz,
                                                      BvhEntrySet &E, const unsigned lane)
{
    const unsigned FULL = 0xffffffffu;
    int e_node = 0;
    float e_lx = -INFINITY, e_ly = -INFINITY, e_lz = -INFINITY, e_hx = INFINITY, e_hy = INFINITY, e_hz = INFINITY;
    bool e_fin = false;
    int n = 1;
    for (int it = 0; it < 64; ++it) {
        /* pick: the first non-final entry of largest extent */
        float key = ((int)lane < n && !e_fin) ? (e_hx - e_lx) + (e_hy - e_ly) + (e_hz - e_lz) : -1.f;
        int who = (int)lane;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {                /* entries live in lanes 0..7 */
            const float k2 = __shfl_xor_sync(FULL, key, o);
            const int w2 = __shfl_xor_sync(FULL, who, o);
            if (k2 > key || (k2 == key && w2 < who)) { key = k2; who = w2; }
        }
        key = __shfl_sync(FULL, key, 0); who = __shfl_sync(FULL, who, 0);
        if (!(key >= 0.f)) break;
        const int pick = who;
        const int pnode = __shfl_sync(FULL, e_node, pick);
        /* children of the picked node, one per lane */
        const int c = (int)(lane % (unsigned)A::W);
        const typename A::Node &N = nodes[pnode];
        const int code = A::code(N, c);
        float lx, ly, lz, hx, hy, hz;
        A::box(N, c, lx, ly, lz, hx, hy, hz);
        const bool in = lane < (unsigned)A::W && !A::empty(code) && lx <= qhx && hx >= qlx && ly <= qhy && hy >= qly && lz <= qhz && hz >= qlz;
        const unsigned hm = __ballot_sync(FULL, in);
        const unsigned leafm = __ballot_sync(FULL, in && code < 0);
        const int nh = __popc(hm);
        if (leafm || n - 1 + nh > BVH_ENTRY_MAX) { if ((int)lane == pick) e_fin = true; continue; }
        if (nh == 0) {                                   /* drop the entry: the last one takes its slot */
            const int last = n - 1;
            const int t_node = __shfl_sync(FULL, e_node, last);
            const float t_lx = __shfl_sync(FULL, e_lx, last), t_ly = __shfl_sync(FULL, e_ly, last), t_lz = __shfl_sync(FULL, e_lz, last);
            const float t_hx = __shfl_sync(FULL, e_hx, last), t_hy = __shfl_sync(FULL, e_hy, last), t_hz = __shfl_sync(FULL, e_hz, last);
            const int t_fin = __shfl_sync(FULL, (int)e_fin, last);
            if ((int)lane == pick && pick != last) { e_node = t_node; e_lx = t_lx; e_ly = t_ly; e_lz = t_lz; e_hx = t_hx; e_hy = t_hy; e_hz = t_hz; e_fin = t_fin != 0; }
            if ((int)lane == last) e_fin = false;
            n = last;
            continue;
        }
        /* the k-th child in range goes to slot pick (k = 0) or n + k - 1; seen from the destination lane: */
        int k = -1;
        if ((int)lane == pick) k = 0;
        else if ((int)lane >= n && (int)lane < n + nh - 1) k = (int)lane - n + 1;
        const int src = k >= 0 ? nth_set_bit4(hm, k) : 0;
        const int t_node = __shfl_sync(FULL, code, src);
        const float t_lx = __shfl_sync(FULL, lx, src), t_ly = __shfl_sync(FULL, ly, src), t_lz = __shfl_sync(FULL, lz, src);
        const float t_hx = __shfl_sync(FULL, hx, src), t_hy = __shfl_sync(FULL, hy, src), t_hz = __shfl_sync(FULL, hz, src);
        if (k >= 0) { e_node = t_node; e_lx = t_lx; e_ly = t_ly; e_lz = t_lz; e_hx = t_hx; e_hy = t_hy; e_hz = t_hz; e_fin = false; }
        n += nh - 1;
    }
    __syncwarp();
    if ((int)lane < n) {
        E.node[lane] = e_node;
        E.lox[lane] = e_lx; E.loy[lane] = e_ly; E.loz[lane] = e_lz;
        E.hix[lane] = e_hx; E.hiy[lane] = e_hy; E.hiz[lane] = e_hz;
    }
    if (lane == 0) E.n = n;
    __syncwarp();
}

/*
 * Warp-cooperative form of bvh4_entry_search2 (bvh_entry.h, version 2: shaft planes, leaf entries, <= BVH_ENTRY2_MAX <= 16
 * entries): the same frontier, picks and slot assignment as the scalar model.  Entry i lives in the registers of lane i.
 * A child of the picked node is tested by the eight lanes c, c + 4, ..: every lane holds two of the twelve shaft planes
 * (lane >> 2 and, for lanes < 16, 8 + (lane >> 2)) in registers, so the four boxes meet all planes in one step and a ballot
 * collects the verdicts.  All 32 lanes must call it with the same arguments; E.rc must hold the boxes R and C (written by
 * the caller, warp-synchronised); q is the padded bundle box.  The result is written to E (shared memory).
 */
template <bool SHAFT>
__device__ __forceinline__ void bvh4_entry_search2_warp(const Bvh4Node *__restrict__ nodes, float qlx, float qly, float qlz, float qhx, float qhy, float qhz,
                                                        float maxabs, BvhEntrySet2 &E, const unsigned lane)
{
    const unsigned FULL = 0xffffffffu;
    float p_nu0 = 0.f, p_nv0 = 0.f, p_ru0 = 0.f, p_rv0 = 0.f, p_lim0 = 0.f, p_nu1 = 0.f, p_nv1 = 0.f, p_ru1 = 0.f, p_rv1 = 0.f, p_lim1 = 0.f;
    const int pa0 = (int)(lane >> 4);                       /* axis pair of plane lane >> 2 (planes 0-3: 0, 4-7: 1); the second plane's is 2 */
    if (SHAFT) {
        bvh_shaft_plane(E.rc, E.rc + 6, maxabs, (int)(lane >> 2), p_nu0, p_nv0, p_ru0, p_rv0, p_lim0);
        if (lane < 16u) bvh_shaft_plane(E.rc, E.rc + 6, maxabs, 8 + (int)(lane >> 2), p_nu1, p_nv1, p_ru1, p_rv1, p_lim1);
    }
    int e_code = 0;
    float e_lx = -INFINITY, e_ly = -INFINITY, e_lz = -INFINITY, e_hx = INFINITY, e_hy = INFINITY, e_hz = INFINITY;
    bool e_fin = false;
    int n = 1;
    for (int it = 0; it < 128; ++it) {
        /* pick: the first non-final entry of largest extent */
        float key = ((int)lane < n && !e_fin) ? (e_hx - e_lx) + (e_hy - e_ly) + (e_hz - e_lz) : -1.f;
        int who = (int)lane;
#pragma unroll
        for (int o = BVH_ENTRY2_MAX > 8 ? 8 : 4; o > 0; o >>= 1) {      /* entries live in lanes 0..BVH_ENTRY2_MAX-1 */
            const float k2 = __shfl_xor_sync(FULL, key, o);
            const int w2 = __shfl_xor_sync(FULL, who, o);
            if (k2 > key || (k2 == key && w2 < who)) { key = k2; who = w2; }
        }
        key = __shfl_sync(FULL, key, 0); who = __shfl_sync(FULL, who, 0);
        if (!(key >= 0.f)) break;
        const int pick = who;
        const int pnode = __shfl_sync(FULL, e_code, pick);
        /* child lane & 3 of the picked node on every lane */
        const int c = (int)(lane & 3u);
        const Bvh4Node &N = nodes[pnode];
        const int code = N.c[c];
        const float lx = N.lox[c], ly = N.loy[c], lz = N.loz[c], hx = N.hix[c], hy = N.hiy[c], hz = N.hiz[c];
        unsigned om = 0u;
        if (SHAFT) {
            bool out;
            {
                const float lu = pa0 == 0 ? lx : ly, hu = pa0 == 0 ? hx : hy, lv = pa0 == 0 ? ly : lz, hv = pa0 == 0 ? hy : hz;
                const float bu = p_nu0 > 0.f ? lu : hu, bv = p_nv0 > 0.f ? lv : hv;
                out = p_nu0 * (bu - p_ru0) + p_nv0 * (bv - p_rv0) > p_lim0;
            }
            {                                             /* axis pair 2 = (z, x); lanes >= 16 hold a null plane (0 > 0 is false) */
                const float bu = p_nu1 > 0.f ? lz : hz, bv = p_nv1 > 0.f ? lx : hx;
                out = out || (p_nu1 * (bu - p_ru1) + p_nv1 * (bv - p_rv1) > p_lim1);
            }
            om = __ballot_sync(FULL, out);
        }
        const bool in = lane < 4u && code != BVH4_EMPTY && lx <= qhx && hx >= qlx && ly <= qhy && hy >= qly && lz <= qhz && hz >= qlz &&
                        !(om & (0x11111111u << c));
        const unsigned hm = __ballot_sync(FULL, in);
        const int nh = __popc(hm);
        if (n - 1 + nh > BVH_ENTRY2_MAX) { if ((int)lane == pick) e_fin = true; continue; }
        if (nh == 0) {                                   /* drop the entry: the last one takes its slot */
            const int last = n - 1;
            const int t_code = __shfl_sync(FULL, e_code, last);
            const float t_lx = __shfl_sync(FULL, e_lx, last), t_ly = __shfl_sync(FULL, e_ly, last), t_lz = __shfl_sync(FULL, e_lz, last);
            const float t_hx = __shfl_sync(FULL, e_hx, last), t_hy = __shfl_sync(FULL, e_hy, last), t_hz = __shfl_sync(FULL, e_hz, last);
            const int t_fin = __shfl_sync(FULL, (int)e_fin, last);
            if ((int)lane == pick && pick != last) { e_code = t_code; e_lx = t_lx; e_ly = t_ly; e_lz = t_lz; e_hx = t_hx; e_hy = t_hy; e_hz = t_hz; e_fin = t_fin != 0; }
            if ((int)lane == last) e_fin = false;
            n = last;
            continue;
        }
        /* the k-th child in range goes to slot pick (k = 0) or n + k - 1; seen from the destination lane: */
        int k = -1;
        if ((int)lane == pick) k = 0;
        else if ((int)lane >= n && (int)lane < n + nh - 1) k = (int)lane - n + 1;
        const int src = k >= 0 ? nth_set_bit4(hm, k) : 0;
        const int t_code = __shfl_sync(FULL, code, src);
        const float t_lx = __shfl_sync(FULL, lx, src), t_ly = __shfl_sync(FULL, ly, src), t_lz = __shfl_sync(FULL, lz, src);
        const float t_hx = __shfl_sync(FULL, hx, src), t_hy = __shfl_sync(FULL, hy, src), t_hz = __shfl_sync(FULL, hz, src);
        if (k >= 0) { e_code = t_code; e_lx = t_lx; e_ly = t_ly; e_lz = t_lz; e_hx = t_hx; e_hy = t_hy; e_hz = t_hz; e_fin = t_code < 0; }   /* a leaf is final at once */
        n += nh - 1;
    }
    __syncwarp();
    if ((int)lane < n) {
        *reinterpret_cast<float4 *>(E.lo[lane]) = make_float4(e_lx, e_ly, e_lz, __int_as_float(e_code));
        *reinterpret_cast<float4 *>(E.hi[lane]) = make_float4(e_hx, e_hy, e_hz, 0.f);
    }
    if (lane == 0) E.n = n;
    __syncwarp();
}

/*
 * Any-hit walk from a version-2 entry set: the ray tests the <= 16 entry boxes (two 16-byte shared-memory loads each, the
 * same words on every lane), pushes the inner nodes it touches and queues the leaves it touches for the triangle phase.
 */
template <int FLUSH = 10>
__device__ __forceinline__ bool bvh4_anyhit_entries2(const Bvh4Node *__restrict__ nodes, const RayTri *__restrict__ tris, const BvhEntrySet2 &E,
                                                     V3 l1, V3 l2, TravStats &ts)
{
    int stack_n[BVH_STACK];
    int sp = 0;
    constexpr int LQ = (FLUSH + 1) / 2;
    unsigned tq[(LQ + 3 > BVH_ENTRY2_MAX ? LQ + 3 : BVH_ENTRY2_MAX) + 4];
    int nq = 0;
    const V3 d = l2 - l1;
    const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
    const V3 l1h = mk3(lb_slab_origin_hi(l1.x, d.x), lb_slab_origin_hi(l1.y, d.y), lb_slab_origin_hi(l1.z, d.z));
    const int n = E.n;
    for (int i = 0; i < n; ++i) {
        const float4 lo = *reinterpret_cast<const float4 *>(E.lo[i]), hi = *reinterpret_cast<const float4 *>(E.hi[i]);
        const float x0 = (lo.x - l1.x) * ix, x1 = (hi.x - l1h.x) * ix, y0 = (lo.y - l1.y) * iy, y1 = (hi.y - l1h.y) * iy;
        const float z0 = (lo.z - l1.z) * iz, z1 = (hi.z - l1h.z) * iz;
        const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
        const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
        if (t0 <= t1 + 2e-6f) {
            const int code = __float_as_int(lo.w);
            if (code >= 0) stack_n[sp++] = code; else tq[nq++] = ~(unsigned)code;
        }
    }
    ts.entries += (unsigned)n;
    if ((sp | nq) == 0) return false;
    const int node = sp ? stack_n[--sp] : -1;
    return bvh4_anyhit_core_q<FLUSH>(nodes, tris, l1, l1h, d, ix, iy, iz, stack_n, sp, node, tq, nq, ts);
}

/* warp-aggregated counter add */
__device__ __forceinline__ void count_add(unsigned long long *counters, int which, unsigned v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counters + which, (unsigned long long)v);
}
