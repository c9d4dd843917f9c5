/*
 * lighter_b200.h -- extension C ABI of liblighter_b200.so (everything the reference's lighter.h
 * does not have).  Plain C types only.  None of these calls is needed for a drop-in bake: a caller
 * that only knows lighter.h gets a single-GPU bake on the current CUDA device.
 *
 * Groups:
 *   ltrx_Set* / ltrx_Nccl* / ltrx_ShardRange  multi-GPU: one process per GPU, lumels sharded,
 *                                              BVH replicated, per-bounce radiance all-gather
 *                                              (replaces the reference's shared-memory thread
 *                                              pool, lighter_int.hpp:17-258 / DoWork call sites
 *                                              lighter.cpp:637-653,1133)
 *   ltrx_GetStats / ltrx_GetError             measurement and error surfacing (the reference only
 *                                              has the stage string, lighter.cpp:1159-1164)
 *   ltrx_Prepare / ltrx_BakeResident          bench support: time the GPU stages with the scene
 *                                              already resident in HBM
 *   ltrx_Get*                                 stage dumps for parity tests (what a reference
 *                                              driver reads from ltr_Scene members)
 *   ltrx_test_*                               primitive / kernel-level entry points for parity tests
 */
#ifndef LIGHTER_B200_H
#define LIGHTER_B200_H

#include "lighter.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LTRX_NCCL_ID_BYTES 128

typedef struct ltrx_Stats {
    /* host wall-clock seconds per stage of the last bake (stage names as in lighter.cpp:1052-1138) */
    double t_total, t_prexform, t_accel, t_upload, t_samples, t_direct, t_radiosity, t_ao, t_finalize, t_readback;
    /* device milliseconds (CUDA events on the bake stream) */
    float gpu_ms_samples, gpu_ms_direct, gpu_ms_march, gpu_ms_radiosity, gpu_ms_ao, gpu_ms_finalize, gpu_ms_total;
    float gpu_ms_rad_pairs, gpu_ms_rad_vis, gpu_ms_span;   /* radiosity pair sweep / visibility kernels; first-to-last event span of the bake */
    /* work counters of THIS rank's shard (units of SURVEY.md 8d) */
    uint64_t n_lumels_total, n_lumels_local, n_triangles, n_bvh_nodes;
    uint64_t n_marches, n_distance_queries, n_ao_segments, n_correction_rays;
    uint64_t n_rad_pairs, n_rad_segments, n_rad_links;
    uint64_t n_node_visits, n_tri_tests;      /* distance-query traversal (march): BVH nodes visited, point/triangle tests */
    uint64_t n_ray_node_visits, n_ray_tri_tests; /* segment traversal (AO + radiosity visibility) */
    uint64_t n_rad_tile_loads;                /* 4 KiB column tiles staged by the radiosity pair sweep */
    uint64_t kernel_launches, h2d_bytes, d2h_bytes;
    uint64_t n_rad_batches;                   /* launches of the pair-sweep / visibility kernel pair (candidate buffer refills) */
    uint64_t n_shadow_rays;                   /* sampled-shadow extension: any-hit rays, lumel x light x sample */
    uint64_t n_ray_entry_tests;               /* entry boxes tested by rays that start at a bundle's entry set instead of the root */
    double t_sample_fn;                       /* host seconds the sample_fn thread ran (request chunks + callbacks); overlaps t_direct and t_radiosity */
} ltrx_Stats;

typedef struct ltrx_Lumels {
    u32 count, width, height;
    const float *pos_xyz;      /* count*3, world position after offset/overlap correction */
    const float *nrm_xyz;      /* count*3 */
    const u32   *loc;          /* texel index y*width+x */
    const float *radinfo_xyzw; /* uv1.x, uv1.y, part, world area per texel */
    const float *rgb;          /* per-lumel colour before finalize */
} ltrx_Lumels;

typedef struct ltrx_Links {
    uint64_t rows, count;
    const uint64_t *row_offset;   /* rows+1; FULL symmetric rows (both directions), partners ascending */
    const u32      *other;
    const float    *factor;
} ltrx_Links;

LTRAPI const char *ltrx_Version(void);
/* example native material callback, assignable to ltr_Config::sample_fn (see bake.cpp) */
LTRAPI LTRBOOL ltrx_SampleFnChecker(ltr_Config *config, ltr_SampleRequest *req);

/* multi-GPU -------------------------------------------------------------------------------- */
LTRAPI int  ltrx_SetDevice(ltr_Scene *scene, int cuda_device);
LTRAPI int  ltrx_NcclUniqueId(unsigned char out_id[LTRX_NCCL_ID_BYTES]);
LTRAPI int  ltrx_SetShard(ltr_Scene *scene, int rank, int world, const unsigned char *nccl_id /* NULL iff world==1 */);
/* sharded bake: 1 = only rank 0 reads the finished lightmaps back to the host (the other ranks report lightmap_count 0);
 * 0 (default) = every rank gets every lightmap, as a caller of the plain API expects */
LTRAPI int  ltrx_SetOutputRoot(ltr_Scene *scene, int root_only);
LTRAPI void ltrx_ShardRange(uint64_t n, int rank, int world, uint64_t *begin, uint64_t *end);

/* direct-light shadow term ----------------------------------------------------------------------
 * LTRX_SHADOW_MARCH (default) is the reference's behaviour: the distance-field penumbra march of
 * CalcInvShadowFactor (lighter.cpp:190-207); shadow_sample_count is ignored, as in the reference.
 * LTRX_SHADOW_SAMPLED is an EXTENSION (the reference has only a disabled sketch, lighter.cpp:566-584):
 * shadow_sample_count (1..64) any-hit rays per lumel and light towards a golden-angle disk of
 * light_radius, VisibilityTest semantics (lighter.cpp:138-147); f_vis = 1 - blocked/samples. */
#define LTRX_SHADOW_MARCH   0
#define LTRX_SHADOW_SAMPLED 1
LTRAPI int ltrx_SetShadowMode(ltr_Scene *scene, int mode);
/* after a bake with ltrx_SetDebug(scene,1) in sampled mode: per local lumel, bit s set = sample s blocked */
LTRAPI int ltrx_GetShadowMasks(ltr_Scene *scene, u32 light, const uint64_t **out, uint64_t *count);
/* the segment sample `sample` of `light` casts from a lumel (host evaluation of the kernel's own function) */
LTRAPI int ltrx_ShadowSampleSegment(ltr_Scene *scene, u32 light, u32 sample, const float pos[3], const float nrm[3],
                                    float from_out[3], float to_out[3]);

/* measurement ------------------------------------------------------------------------------ */
LTRAPI int         ltrx_GetStats(ltr_Scene *scene, ltrx_Stats *out);
LTRAPI const char *ltrx_GetError(ltr_Scene *scene);   /* "" when the last bake succeeded */
LTRAPI int         ltrx_Prepare(ltr_Scene *scene);    /* host pre-pass + upload; synchronous */
LTRAPI int         ltrx_BakeResident(ltr_Scene *scene, float *gpu_ms_out); /* GPU stages only, synchronous */
LTRAPI int         ltrx_Finish(ltr_Scene *scene);     /* read back outputs after ltrx_BakeResident */
/* FNV-1a-64 of the float bytes of all lightmaps in output order followed by the probe colours (0 = no outputs yet) */
LTRAPI int         ltrx_OutputHash(ltr_Scene *scene, uint64_t *fnv1a64);

/* stage dumps (valid after a bake run with ltrx_SetDebug(scene,1)) --------------------------- */
LTRAPI int ltrx_SetDebug(ltr_Scene *scene, int keep_stage_arrays);
LTRAPI int ltrx_GetLumels(ltr_Scene *scene, u32 instance, ltrx_Lumels *out);
LTRAPI int ltrx_GetLinks(ltr_Scene *scene, ltrx_Links *out);
LTRAPI int ltrx_GetShadowFactors(ltr_Scene *scene, u32 light, const float **out, uint64_t *count);

/* kernel-level entry points: host arrays in, host arrays out, device round trip inside -------- */
LTRAPI int ltrx_test_point_tri_distance(const float *pts3, const float *tris9, u32 n, float *out);
LTRAPI int ltrx_test_seg_tri(const float *a3, const float *b3, const float *tris9, u32 n, float *out);
LTRAPI int ltrx_test_scene_queries(const float *tris9, u32 ntris,
                                   const float *a3, const float *b3, u32 n,
                                   float *dist_out,      /* min(2, nearest distance) at a */
                                   int   *anyhit_out,    /* segment a-b blocked (end points NOT shortened) */
                                   float *closest_out,   /* closest-hit parameter on a-b, 2 = none */
                                   int   *closest_tri_out);
LTRAPI int ltrx_test_march(const float *tris9, u32 ntris, const float *from3, const float *to3,
                           const float *k, u32 n, float *out, u32 *steps_out);
/* the scene BVH built on the device (csrc/gpu_bvh.cu) against the host builder (csrc/bvh.cpp) on the same triangles: sizes,
 * inner levels, device build time (CUDA events) and the number of records that differ (0 for non-degenerate input) */
LTRAPI int ltrx_test_device_bvh(const float *tris9, u32 ntris, int leaf_max, u32 *n_nodes, u32 *n_nodes4, int *height, float *build_ms,
                                u32 *mismatch_nodes, u32 *mismatch_nodes4, u32 *mismatch_leaves, int *host_height);
LTRAPI int ltrx_test_spiral_dirs(const float *nrm3, const float *randoff, u32 n, int samples, float *out3);
/* host-only builders (no GPU): the reference-order tree (8 words per node: lo3, hi3, ch, ido; item stream
 * "<count> ids...") and the flat scene BVH with a structural self-check (returns 0 when it fails) */
LTRAPI int ltrx_test_reftree(const float *tris9, u32 ntris, void *nodes_out, u32 nodes_cap, int32_t *items_out, u32 items_cap,
                             u32 *n_nodes, u32 *n_items);
/* n draws of the process's libc rand() stream as randf() = rand()/RAND_MAX (the AO pass's per-lumel offsets); returns 1
 * when the lock-free table-level path was used, 0 for the plain rand() loop -- same values, same libc state afterwards */
LTRAPI int ltrx_test_rand_fill(float *out, uint64_t n);
LTRAPI int ltrx_test_bvh(const float *tris9, u32 ntris, int leaf_max, u32 *n_nodes, u32 *depth, u32 *order_out, float *bounds6);
/* host-only: per-segment cost (4-wide node reads, triangle tests) of the walk from the bundle's entry set */
LTRAPI int ltrx_test_bvh_entry_cost(const float *tris9, u32 ntris, int leaf_max, const float *segs6, const u32 *bundle_off, u32 n_bundles,
                                    uint32_t *nodes_out, uint32_t *tris_out);
/* host-only: the conservative culling tests of the radiosity pair sweep (csrc/rad_cull.h) on one row block x one column block */
LTRAPI int ltrx_test_rad_cull(const float *rowP3, const float *rowN3, u32 nrows, const float *colP3, const float *colN3, u32 ncols,
                              int *block_ok, uint8_t *row_ok, uint8_t *pair_fast /* nrows*ncols or NULL */);
/* host-only: the host pre-pass of a scene without a device; FNV-1a fingerprints of the arrays it would upload */
LTRAPI int ltrx_test_host_prepare(ltr_Scene *scene, uint64_t out_hash[12]);
/* host-only: entry sets of segment bundles (csrc/bvh_entry.h) -- reachability self-check, root walk vs entry walk */
LTRAPI int ltrx_test_bvh_entry(const float *tris9, u32 ntris, int leaf_max, const float *segs6, const u32 *bundle_off, u32 n_bundles,
                               u32 *entries_out, uint64_t *visits_root, uint64_t *visits_entry, uint64_t *entry_tests, u32 *mismatches);
/* host-only: version-2 entry sets (shaft-culled search, leaf entries): root walk vs entry walk incl. triangle-test counts */
LTRAPI int ltrx_test_bvh_entry2(const float *tris9, u32 ntris, int leaf_max, const float *segs6, const u32 *bundle_off, u32 n_bundles,
                                int max_entries, int use_shaft, int batch /* > 0: packet form, leaves listed per `batch` segments */,
                                u32 *entries_out, uint64_t *stats4, u32 *mismatches, u32 *test_diffs);

#ifdef __cplusplus
}
#endif
#endif
