/*
 * gpu.h -- the thin C-ABI layer between the host C++ pipeline (bake.cpp) and the CUDA side
 * (gpu_*.cu).  POD descriptors and extern "C" entry points only: host code never sees a kernel,
 * a stream or a device pointer, the CUDA side never sees an ltr_Scene.
 *
 * Every entry point returns 0 on success; on failure the message is available from
 * ltrgpu_last_error().  All work is issued on the context's own stream; entry points that hand
 * data back to the host synchronise that stream, the others only enqueue.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "bvh.h"
#include "geom.h"
#include "reftree.h"
#include "vmath.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ltrgpu_Ctx ltrgpu_Ctx;

typedef struct ltrgpu_Light {          /* 80 bytes; ref: ltr_Light lighter_int.hpp:906-925 */
    V3 pos;    uint32_t type;
    V3 dir;    float range;
    V3 color;  float power;
    float radius, angle_out_rad, angle_diff, curve;
    uint32_t n_samples, sample_off, pad0, pad1;   /* sampled-shadow extension: slice of the sample table */
} ltrgpu_Light;

typedef struct ltrgpu_Inst {           /* 48 bytes */
    uint32_t lm_w, lm_h;
    uint32_t texel_off_lo, texel_off_hi;   /* start of this instance's image in the concatenated texel space */
    uint32_t node_off, item_off, tri_off;  /* slices of the concatenated reference-order trees */
    uint32_t tree_tris;                    /* triangles in that tree */
    uint32_t shadow;                       /* instance occludes (scene BVH membership) */
    uint32_t pad0, pad1, pad2;
} ltrgpu_Inst;

typedef struct ltrgpu_RasterTri {      /* one lightmap-UV triangle in reference raster order */
    uint32_t inst, part, i0, i1, i2;   /* vertex indices into the concatenated world-space arrays */
} ltrgpu_RasterTri;

typedef struct ltrgpu_Params {         /* the slice of ltr_Config the device needs */
    float ambient[3];
    float max_correct_dist, corr_min_dot;
    float ao_distance, ao_multiplier, ao_falloff, ao_effect, ao_color[3];
    int   ao_num_samples;
    float blur_size;
    int   ds2x, normalmap;
    float amb_brightness;
    int   shadow_mode;                 /* 0 = the reference's distance march, 1 = sampled any-hit shadow rays (extension) */
} ltrgpu_Params;

typedef struct ltrgpu_SceneDesc {
    ltrgpu_Params params;
    /* instances (index 0 = probe container) */
    uint32_t n_inst;          const ltrgpu_Inst *inst;
    /* world-space vertex streams, concatenated over instances */
    uint32_t n_verts;         const V3 *wpos; const V3 *wnrm; const float *vtex2; const float *ltex2;
    uint32_t n_rtris;         const ltrgpu_RasterTri *rtris;
    /* reference-order trees, concatenated over instances */
    uint32_t n_rnodes;        const RefNode *rnodes;
    uint32_t n_ritems;        const int32_t *ritems;
    uint32_t n_rtree_tris;    const float *rtree_tris9;
    /* flat scene BVH over shadow-casting triangles.
     *   bvh != NULL: built on the host (bvh.cpp), tris9 already in BVH order, tri_orig = original index per slot;
     *   bvh == NULL: tris9 is in SCENE order and the tree is built on the device (gpu_bvh.cu) with leaves of <= bvh_leaf_max
     *                triangles; tris9 may alias rtree_tris9 (every instance casts shadows), then it is uploaded once. */
    uint32_t n_bvh_nodes;     const BvhNode *bvh;
    uint32_t n_bvh4_nodes;    const Bvh4Node *bvh4;            /* 4-wide collapse of the same tree (any-hit walks) */
    uint32_t n_tris;          const float *tris9; const uint32_t *tri_orig;
    int bvh_leaf_max, bvh_height;                              /* bvh_height: inner levels of a host-built tree */
    /* multi-GPU: the concatenated per-instance arrays hold valid data only in THIS rank's slice; shard_x[r] .. shard_x[r+1]
     * (elements, world + 1 entries, NULL = everything is valid) is rank r's slice of: wpos / wnrm / vtex2 / ltex2 (verts),
     * rtris, rnodes, ritems, rtree_tris9 (tris).  The upload sends the slice and completes the device arrays by an
     * all-gather over NVLink. */
    const uint64_t *shard_verts, *shard_rtris, *shard_rnodes, *shard_ritems, *shard_tris;
    int scene_covers_rtree;                                    /* every instance casts shadows: the scene BVH holds exactly the triangles of the instance trees */
    /* lights + light->instance visibility table [n_lights][n_inst] */
    uint32_t n_lights;        const ltrgpu_Light *lights; const uint8_t *light_inst;
    uint32_t n_light_samples; const float *light_samples4;     /* sampled-shadow extension: float4 per (light, sample) */
    /* probes */
    uint32_t n_probes;        const V3 *probe_pos; const V3 *probe_nrm;
    /* AO hemisphere table: cos_side[s], sin_side[s] for s < ao_num_samples (host libm) */
    const float *ao_cos_side; const float *ao_sin_side;
    /* gaussian kernel taps (host libm), 2*ext+1 entries, ext = ceil(blur_size) */
    int blur_ext;             const float *blur_kernel;
} ltrgpu_SceneDesc;

typedef struct ltrgpu_Counters {
    uint64_t marches, distance_queries, ao_segments, correction_rays;
    uint64_t rad_pairs, rad_segments, rad_links;
    uint64_t node_visits, tri_tests;           /* distance queries (march): 64 B nodes, 160 B prepared triangles */
    uint64_t ray_node_visits, ray_tri_tests;   /* segment queries (AO, radiosity): 64 B nodes, 64 B ray triangles */
    uint64_t rad_tile_loads;                   /* 128-lumel column tiles (4 KiB) staged into shared memory by the pair sweep */
    uint64_t kernel_launches, h2d_bytes, d2h_bytes;
    uint64_t rad_batches;                      /* pair-sweep + visibility launch pairs */
    uint64_t shadow_rays;                      /* sampled-shadow extension: any-hit rays lumel x light x sample */
    uint64_t ray_entry_tests;                  /* entry boxes tested by rays started from a bundle's entry set (bvh_entry.h) */
    float ms_samples, ms_direct, ms_march, ms_radiosity, ms_ao, ms_finalize, ms_rad_pairs, ms_rad_vis, ms_span;
} ltrgpu_Counters;

/* all-gather hook for the multi-GPU radiance exchange: gathers `bytes_per_rank` bytes from
 * `send` (device) of every rank into `recv` (device, world*bytes_per_rank) on `stream`. */
typedef int (*ltrgpu_allgather_fn)(void *user, const void *send, void *recv, size_t bytes_per_rank, void *cuda_stream);
/* in-place all-gather of UNEVEN contiguous slices: rank r owns bytes [offsets[r], offsets[r+1]) of `buf` (device) and every
 * rank ends up with all of them (the direct-light colours: runs of equal march cost, not of equal length) */
typedef int (*ltrgpu_gatherv_fn)(void *user, void *buf, const uint64_t *offsets /* world+1 */, void *cuda_stream);

int  ltrgpu_create(ltrgpu_Ctx **out, int device);
void ltrgpu_destroy(ltrgpu_Ctx *ctx);
const char *ltrgpu_last_error(ltrgpu_Ctx *ctx);
void *ltrgpu_stream(ltrgpu_Ctx *ctx);
void ltrgpu_release_memory(void);           /* return the caching allocator's idle device blocks to the driver */

int ltrgpu_upload_scene(ltrgpu_Ctx *ctx, const ltrgpu_SceneDesc *desc);
/* the scene BVH as resident on the device: binary node count, inner levels, device build time (0 = host-built) */
int ltrgpu_bvh_info(ltrgpu_Ctx *ctx, uint32_t *n_nodes, int *height, float *build_ms);

/* stage: lumel generation (raster -> ordered compaction -> concave-edge offset -> overlap correction).
 * inst_lumel_off receives n_inst+1 prefix offsets into the global lumel array (probes first). */
/* Early start of the device-built scene BVH, so that it runs beside the host's reference-order tree builds instead of after
 * them (bake.cpp host_prepare).  _tris_early: the reference-order triangle array goes up now (sharded: completed by a
 * collective -- call it on the bake thread, at the same point on every rank).  _bvh_early: builds the scene BVH over those
 * triangles on the context's SECOND stream and blocks until the build has been queued to its end; it contains no collective
 * and may run on another host thread.  ltrgpu_upload_scene then skips both steps (it checks the triangle count) and makes
 * the bake stream wait for the build. */
int ltrgpu_upload_tris_early(ltrgpu_Ctx *ctx, const float *rtree_tris9, uint32_t n_rtree_tris, const uint64_t *shard_tris /* NULL or world + 1 prefix offsets */);
int ltrgpu_build_bvh_early(ltrgpu_Ctx *ctx, int bvh_leaf_max);
const char *ltrgpu_early_error(ltrgpu_Ctx *ctx);

int ltrgpu_generate_lumels(ltrgpu_Ctx *ctx, uint64_t *inst_lumel_off);

/* multi-GPU identity of this context and the all-gather hook; call before ltrgpu_generate_lumels */
int ltrgpu_set_world(ltrgpu_Ctx *ctx, int rank, int world, ltrgpu_allgather_fn allgather, void *allgather_user);
int ltrgpu_set_gatherv(ltrgpu_Ctx *ctx, ltrgpu_gatherv_fn gatherv);
/* personalised exchange: bytes [send_off[r], send_off[r+1]) of `send` go to rank r, the bytes from rank r land at
 * [recv_off[r], recv_off[r+1]) of `recv` (offsets: world + 1 entries each; nothing is sent to oneself) */
typedef int (*ltrgpu_alltoallv_fn)(void *user, const void *send, const uint64_t *send_off, void *recv, const uint64_t *recv_off, void *cuda_stream);
int ltrgpu_set_alltoallv(ltrgpu_Ctx *ctx, ltrgpu_alltoallv_fn fn);
/* all-gather of a small HOST table through the device and the all-gather hook: recv = world x bytes, rank-major (synchronous) */
int ltrgpu_host_allgather(ltrgpu_Ctx *ctx, const void *send, void *recv, size_t bytes);            /* same user pointer as the all-gather hook */

/* restrict the per-lumel stages to global lumels [begin,end) (multi-GPU shard); default = all */
int ltrgpu_set_shard(ltrgpu_Ctx *ctx, uint64_t begin, uint64_t end, int rank, int world,
                     ltrgpu_allgather_fn allgather, void *allgather_user);

int ltrgpu_direct_light(ltrgpu_Ctx *ctx);

/* host copies of the lumel arrays (n = global lumel count); any pointer may be NULL */
int ltrgpu_download_lumels(ltrgpu_Ctx *ctx, float *pos3, float *nrm3, uint32_t *loc, float *radinfo4, float *rgb3);

/* stage: radiosity.  diffuse3 / emissive3: per global lumel material from the host callback, or
 * NULL for (1,1,1) / no extra emission on mesh lumels (probes always diffuse 0, area 0). */
int ltrgpu_radiosity(ltrgpu_Ctx *ctx, const float *diffuse3, const float *emissive3, int bounces);
/* The same with the materials asked for as late as possible: `fn` is called once, after link generation and right before
 * the first bounce reads the materials (it may block: bake.cpp waits there for the sample_fn callbacks, which ran on a
 * host thread while the GPU did direct light and link generation).  fn returns 0 and the two host arrays (or NULLs). */
typedef int (*ltrgpu_materials_fn)(void *user, const float **diffuse3, const float **emissive3);
int ltrgpu_radiosity_ex(ltrgpu_Ctx *ctx, ltrgpu_materials_fn fn, void *user, int bounces);

/* sample_fn batching (ref: lighter.cpp:690-715 builds one ltr_SampleRequest per mesh lumel on the host).  The request
 * fields are computed on the device -- normalised normal, tex1 = (texel + 0.5) / lightmap size, with the reference's float
 * operations -- and come back in chunks on a SECOND stream, so a host thread can run the callbacks of chunk k while chunk
 * k+1 is packed and copied and while the bake stream runs direct light and link generation.  Thread contract: _begin on
 * the bake thread after ltrgpu_generate_lumels; _issue / _wait from ONE other thread. */
typedef struct ltrgpu_SampleReq {      /* 48 bytes */
    float pos[3], nrm[3];              /* lumel position, NORMALISED normal */
    float tex0[2], tex1[2];
    uint32_t part_id, inst;
} ltrgpu_SampleReq;
int ltrgpu_sample_requests_begin(ltrgpu_Ctx *ctx);
int ltrgpu_sample_requests_issue(ltrgpu_Ctx *ctx, uint64_t first, uint32_t count, ltrgpu_SampleReq *host_pinned, int slot /* 0 or 1 */);
int ltrgpu_sample_requests_wait(ltrgpu_Ctx *ctx, int slot);
const char *ltrgpu_aux_error(ltrgpu_Ctx *ctx);        /* error text of the three calls above when made from the material thread */

/* stage: ambient occlusion; randoff has one value per lumel OF THIS SHARD (host rand() replay, sliced by the caller) */
int ltrgpu_ambient_occlusion(ltrgpu_Ctx *ctx, const float *randoff);

/* stage: finalize (gather shards, scatter, 3x dilation, blur, ds2x, normal map) */
int ltrgpu_finalize(ltrgpu_Ctx *ctx);
int ltrgpu_output_size(ltrgpu_Ctx *ctx, uint32_t inst, uint32_t *w, uint32_t *h);
int ltrgpu_download_output(ltrgpu_Ctx *ctx, uint32_t inst, float *rgb, float *normals_xyzf /* may be NULL */);
int ltrgpu_download_probe_colors(ltrgpu_Ctx *ctx, float *rgb3);
/* all lightmaps in one copy: rgb_all receives the concatenated output images (instance i at float offset
 * 3*out_off[i], out_off = n_inst+1 prefix offsets in texels); total texel count returned in *texels */
int ltrgpu_output_layout(ltrgpu_Ctx *ctx, uint64_t *out_off /* n_inst+1 */);
int ltrgpu_download_outputs_all(ltrgpu_Ctx *ctx, float *rgb_all);

/* Page-locked host memory from a process-wide cache (a bake moves ~200 MB each way; pinning costs more than
 * the copy, so blocks are kept and reused across bakes).  Falls back to malloc without a CUDA device. */
void *ltrgpu_host_alloc(size_t bytes);
void ltrgpu_host_free(void *p);

int ltrgpu_sync(ltrgpu_Ctx *ctx);
int ltrgpu_get_counters(ltrgpu_Ctx *ctx, ltrgpu_Counters *out);
int ltrgpu_reset_bake(ltrgpu_Ctx *ctx);     /* drop lumels/results, keep the uploaded scene (bench re-runs) */
int ltrgpu_span_begin(ltrgpu_Ctx *ctx);     /* CUDA events on the bake stream bracketing all GPU stages of one bake */
int ltrgpu_span_end(ltrgpu_Ctx *ctx);

/* debug dumps */
int ltrgpu_download_shadow_factors(ltrgpu_Ctx *ctx, uint32_t light, float *out /* local lumels */);
int ltrgpu_download_shadow_masks(ltrgpu_Ctx *ctx, uint32_t light, uint64_t *out /* local lumels; bit s = sample s blocked */);
int ltrgpu_download_links(ltrgpu_Ctx *ctx, uint64_t *row_offset, uint32_t *other, float *factor, uint64_t *rows, uint64_t *count);

#ifdef __cplusplus
}
#endif
