/*
 * bake.cpp -- the bake orchestrator: what the reference's Job_MainProc (lighter.cpp:1046-1145) is
 * to its thread pool, this is to the GPU.  Fixed stage order and stage strings are the
 * reference's; every stage body is a call through the C-ABI layer in gpu.h.
 *
 * Host-side work kept here (cheap, and bit-exact on the host's IEEE float):
 *   - pre-transform, total area, the size_fn callback, lightmap UV scaling  (lighter.cpp:292-341)
 *   - shadow triangle lists, reference-order trees, the instance tree       (lighter.cpp:349-384,1060-1069)
 *   - the flat scene BVH                                                     (bvh.cpp)
 *   - light -> instance culling                                              (lighter.cpp:80-98,1080-1095)
 *   - libc rand() replay for the AO offsets                                  (lighter.cpp:819)
 *   - the sample_fn material callback round trip                             (lighter.cpp:692-714)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <map>
#include <mutex>
#include <thread>

#include "gpu.h"
#include "bvh_entry.h"
#include "rad_cull.h"
#include "nccl_dl.h"
#include "randfill.h"
#include "scene.h"

namespace {

double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- 4x4 inverse by cofactors, term order of the reference (lighter_math.cpp:19-147) so the
 *      normal matrix comes out bit-identical.  Each row lists six (i,j,k) products; signs
 *      alternate + - - + + - for "even" rows and - + + - - + for "odd" rows. ---- */
struct Cof { unsigned char out, odd, t[6][3]; };
const Cof COF[16] = {
    { 0, 0, { {5,10,15}, {5,11,14}, {9,6,15}, {9,7,14}, {13,6,11}, {13,7,10} } },
    { 4, 1, { {4,10,15}, {4,11,14}, {8,6,15}, {8,7,14}, {12,6,11}, {12,7,10} } },
    { 8, 0, { {4,9,15}, {4,11,13}, {8,5,15}, {8,7,13}, {12,5,11}, {12,7,9} } },
    { 12, 1, { {4,9,14}, {4,10,13}, {8,5,14}, {8,6,13}, {12,5,10}, {12,6,9} } },
    { 1, 1, { {1,10,15}, {1,11,14}, {9,2,15}, {9,3,14}, {13,2,11}, {13,3,10} } },
    { 5, 0, { {0,10,15}, {0,11,14}, {8,2,15}, {8,3,14}, {12,2,11}, {12,3,10} } },
    { 9, 1, { {0,9,15}, {0,11,13}, {8,1,15}, {8,3,13}, {12,1,11}, {12,3,9} } },
    { 13, 0, { {0,9,14}, {0,10,13}, {8,1,14}, {8,2,13}, {12,1,10}, {12,2,9} } },
    { 2, 0, { {1,6,15}, {1,7,14}, {5,2,15}, {5,3,14}, {13,2,7}, {13,3,6} } },
    { 6, 1, { {0,6,15}, {0,7,14}, {4,2,15}, {4,3,14}, {12,2,7}, {12,3,6} } },
    { 10, 0, { {0,5,15}, {0,7,13}, {4,1,15}, {4,3,13}, {12,1,7}, {12,3,5} } },
    { 14, 1, { {0,5,14}, {0,6,13}, {4,1,14}, {4,2,13}, {12,1,6}, {12,2,5} } },
    { 3, 1, { {1,6,11}, {1,7,10}, {5,2,11}, {5,3,10}, {9,2,7}, {9,3,6} } },
    { 7, 0, { {0,6,11}, {0,7,10}, {4,2,11}, {4,3,10}, {8,2,7}, {8,3,6} } },
    { 11, 1, { {0,5,11}, {0,7,9}, {4,1,11}, {4,3,9}, {8,1,7}, {8,3,5} } },
    { 15, 0, { {0,5,10}, {0,6,9}, {4,1,10}, {4,2,9}, {8,1,6}, {8,2,5} } },
};

bool invert4(const float *a, float *out)
{
    static const float SGN[2][6] = { { 1, -1, -1, 1, 1, -1 }, { -1, 1, 1, -1, -1, 1 } };
    float inv[16];
    for (const Cof &c : COF) {
        float acc = 0;
        for (int k = 0; k < 6; ++k) {
            float term = (SGN[c.odd][k] * a[c.t[k][0]]) * a[c.t[k][1]] * a[c.t[k][2]];
            acc = k == 0 ? term : acc + term;
        }
        inv[c.out] = acc;
    }
    float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    if (det == 0) return false;
    det = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * det;
    return true;
}

inline V3 xform(const float *m, V3 v, float w)
{
    return mk3(v.x * m[0] + v.y * m[4] + v.z * m[8] + m[12] * w,
               v.x * m[1] + v.y * m[5] + v.z * m[9] + m[13] * w,
               v.x * m[2] + v.y * m[6] + v.z * m[10] + m[14] * w);
}

float heron(float a, float b, float c)
{
    float p = (a + b + c) * 0.5f;
    float q = p * (p - a) * (p - b) * (p - c);
    return q < 0 ? 0 : sqrtf(q);
}

int nccl_allgather_cb(void *user, const void *send, void *recv, size_t bytes, void *stream);

} // namespace

struct Bake {
    ltrgpu_Ctx *gpu = nullptr;
    std::vector<ltrgpu_Inst> inst;
    /* big arrays, every element written after resize(): BigVec = no zero-fill, huge pages on request (bvh.h) */
    BigVec<V3> wpos, wnrm;
    BigVec<float> vtex, ltex;
    BigVec<ltrgpu_RasterTri> rtris;
    BigVec<RefNode> rnodes;
    BigVec<int32_t> ritems;
    BigVec<float> rtree_tris;
    SceneBvh bvh;
    BigVec<float> bvh_tris;                   /* host-built tree: triangles in BVH order; device build: the compacted scene-order triangles (if not every instance casts) */
    /* multi-GPU host pre-pass: instances are dealt to the ranks in contiguous runs; a rank transforms, lists and builds the
     * reference-order trees of ITS instances only, uploads those slices, and the device arrays are completed by an
     * all-gather over NVLink (upload()).  inst_cut[r] = first instance of rank r; sh_* = the slice boundaries of the
     * concatenated arrays in elements, world + 1 entries each (empty = one rank: everything is local). */
    std::vector<size_t> inst_cut;
    std::vector<uint64_t> sh_verts, sh_rtris, sh_rnodes, sh_ritems, sh_tris;
    bool all_cast = true;                     /* every instance with triangles casts shadows: scene BVH == the triangles of the instance trees */
    bool device_bvh = true;                   /* the scene BVH is built on the device (gpu_bvh.cu); false: bvh.cpp (LTR_BVH_HOST=1, tiny scenes, host-only test hooks) */
    const float *scene_tris = nullptr;        /* device build: the triangles the tree is built over, scene order */
    size_t n_scene_tris = 0;
    int bvh_leaf_max = BVH_LEAF_MAX;
    std::vector<ltrgpu_Light> lights;
    std::vector<float> light_samples;         /* float4 per (light, sample): sampled-shadow extension table (host libm) */
    std::vector<uint8_t> light_inst;
    std::vector<float> ao_cos, ao_sin, blur_kernel;
    int blur_ext = 0;
    std::vector<uint64_t> lumel_off;
    bool prepared = false;
    void *comm = nullptr;
    const NcclApi *nccl = nullptr;
    int world = 1;
    /* debug copies */
    std::vector<float> d_pos, d_nrm, d_rad, d_rgb, d_fvis, d_lfac;
    std::vector<uint32_t> d_loc, d_lother;
    std::vector<uint64_t> d_lrow, d_smask;
};

namespace {

int nccl_allgather_cb(void *user, const void *send, void *recv, size_t bytes, void *stream)
{
    Bake *B = (Bake *)user;
    if (!B->nccl || !B->comm) return 1;
    int rc = B->nccl->AllGather(send, recv, bytes, /*ncclUint8*/ 1, B->comm, stream);
    if (rc != 0) fprintf(stderr, "lighter_b200: ncclAllGather failed: %s\n", B->nccl->GetErrorString(rc));
    return rc;
}

/* uneven in-place all-gather: one broadcast per rank, fused into one NCCL group */
int nccl_gatherv_cb(void *user, void *buf, const uint64_t *off, void *stream)
{
    Bake *B = (Bake *)user;
    if (!B->nccl || !B->comm) return 1;
    int rc = B->nccl->GroupStart();
    for (int r = 0; rc == 0 && r < B->world; ++r) {
        if (off[r + 1] <= off[r]) continue;
        char *p = (char *)buf + off[r];
        rc = B->nccl->Broadcast(p, p, (size_t)(off[r + 1] - off[r]), /*ncclUint8*/ 1, r, B->comm, stream);
    }
    const int rc2 = B->nccl->GroupEnd();
    if (rc == 0) rc = rc2;
    if (rc != 0) fprintf(stderr, "lighter_b200: ncclBroadcast group failed: %s\n", B->nccl->GetErrorString(rc));
    return rc;
}

/* personalised exchange (mirrored radiosity links go to the rank that owns their row): grouped ncclSend / ncclRecv */
int nccl_alltoallv_cb(void *user, const void *send, const uint64_t *soff, void *recv, const uint64_t *roff, void *stream)
{
    Bake *B = (Bake *)user;
    if (!B->nccl || !B->comm) return 1;
    int rc = B->nccl->GroupStart();
    for (int r = 0; rc == 0 && r < B->world; ++r) {
        if (soff[r + 1] > soff[r]) rc = B->nccl->Send((const char *)send + soff[r], (size_t)(soff[r + 1] - soff[r]), /*ncclUint8*/ 1, r, B->comm, stream);
        if (rc == 0 && roff[r + 1] > roff[r]) rc = B->nccl->Recv((char *)recv + roff[r], (size_t)(roff[r + 1] - roff[r]), /*ncclUint8*/ 1, r, B->comm, stream);
    }
    const int rc2 = B->nccl->GroupEnd();
    if (rc == 0) rc = rc2;
    if (rc != 0) fprintf(stderr, "lighter_b200: ncclSend/ncclRecv group failed: %s\n", B->nccl->GetErrorString(rc));
    return rc;
}

struct Fail { std::string msg; };

void gpu_check(ltr_Scene *S, int rc, const char *what)
{
    if (rc == 0) return;
    Fail f;
    f.msg = std::string(what) + ": " + (S->bake && S->bake->gpu ? ltrgpu_last_error(S->bake->gpu) : "no GPU context");
    throw f;
}

/* ------------------------------------------------------------------------------------------
 * host pre-pass
 * ------------------------------------------------------------------------------------------ */
void host_prepare(ltr_Scene *S, bool force_host_bvh = false)
{
    Bake &B = *S->bake;
    const ltr_Config &cfg = S->config;
    const size_t ni = S->instances.size();
    double t0 = now_s();

    S->stage.store("transforming spatial data");
    S->completion.store(0.f);
    B.inst.assign(ni, ltrgpu_Inst());
    std::vector<size_t> vbase(ni + 1, 0);
    for (size_t i = 1; i < ni; ++i) vbase[i + 1] = vbase[i] + S->instances[i]->mesh->vpos.size();
    const size_t nv = vbase[ni];
    B.wpos.resize(nv); B.wnrm.resize(nv); B.vtex.resize(nv * 2); B.ltex.resize(nv * 2);

    /* Which instances are this rank's?  One rank, a host-built BVH (it needs every triangle here) or the host-only test
     * hook: all of them.  Otherwise contiguous runs of about equal weight (vertices + triangles, what the transform and the
     * tree builds cost); instance 0 (the probe container) goes with rank 0. */
    bool every_instance_casts = true;              /* otherwise the scene BVH needs a compacted triangle list, made on the host from all triangles */
    for (size_t i = 1; i < ni; ++i) if (!S->instances[i]->shadow) every_instance_casts = false;
    const int world = (S->world > 1 && !force_host_bvh && !getenv("LTR_BVH_HOST") && !getenv("LTR_HOST_PREPASS_REPLICATED") && every_instance_casts && B.gpu && B.comm) ? S->world : 1;
    const int rank = world > 1 ? S->rank : 0;
    B.inst_cut.assign((size_t)world + 1, ni);
    B.inst_cut[0] = 0;
    if (world > 1) {
        std::vector<uint64_t> w(ni + 1, 0);
        for (size_t i = 1; i < ni; ++i) w[i + 1] = w[i] + S->instances[i]->mesh->vpos.size() + S->instances[i]->mesh->indices.size();
        int r = 1;
        for (size_t i = 1; i < ni && r < world; ++i)
            while (r < world && w[i + 1] * (uint64_t)world >= w[ni] * (uint64_t)r) B.inst_cut[r++] = i + 1;
    }
    const size_t my_i0 = B.inst_cut[rank], my_i1 = B.inst_cut[rank + 1];
    /* the processes of a multi-GPU bake share the host's cores */
    const unsigned hw_threads = std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, S->world));
    /* per-instance results the other ranks need: two small tables, all-gathered (exchange below) */
    auto exchange = [&](std::vector<uint32_t> &tab, size_t k) {     /* tab: ni x k words, valid for my instances; filled for all on return */
        if (world <= 1) return;
        std::vector<uint32_t> all(tab.size() * (size_t)world);
        gpu_check(S, ltrgpu_host_allgather(B.gpu, tab.data(), all.data(), tab.size() * 4), "instance table exchange");
        for (int r = 0; r < world; ++r)
            for (size_t i = B.inst_cut[r]; i < B.inst_cut[r + 1]; ++i)
                memcpy(&tab[i * k], &all[(size_t)r * tab.size() + i * k], k * 4);
    };

    /* one instance per task, as the reference does (its size_fn is called from pool threads, one per instance, concurrently:
     * lighter.cpp:1054,326) */
    auto xform_instance = [&](size_t i) {
        MeshInstance *mi = S->instances[i];
        ltr_Mesh *mesh = mi->mesh;
        const float *M = mi->matrix;
        float NM[16] = { M[0], M[1], M[2], 0, M[4], M[5], M[6], 0, M[8], M[9], M[10], 0, 0, 0, 0, 1 };
        float inv[16];
        if (invert4(NM, inv)) memcpy(NM, inv, sizeof(NM));
        std::swap(NM[4], NM[1]); std::swap(NM[8], NM[2]); std::swap(NM[12], NM[3]);
        std::swap(NM[9], NM[6]); std::swap(NM[13], NM[7]); std::swap(NM[14], NM[11]);
        V3 *wp = B.wpos.data() + vbase[i], *wn = B.wnrm.data() + vbase[i];
        for (size_t v = 0; v < mesh->vpos.size(); ++v) {
            wp[v] = xform(M, mesh->vpos[v], 1.0f);
            wn[v] = norm3(xform(NM, mesh->vnrm[v], 0.0f));
        }
        float total_area = 0.0f;
        for (const MeshPart &mp : mesh->parts) {
            const V3 *vb = wp + mp.vertex_offset;
            const u32 *ib = mesh->indices.data() + mp.index_offset;
            for (u32 t = 0; t + 2 < mp.index_count; t += 3) {
                V3 p1 = vb[ib[t]], p2 = vb[ib[t + 1]], p3 = vb[ib[t + 2]];
                total_area += heron(len3(p2 - p1), len3(p3 - p2), len3(p1 - p3));
            }
        }
        u32 size[2] = { cfg.default_width, cfg.default_height };
        auto fn = cfg.size_fn ? cfg.size_fn : ltr_DefaultSizeFunc;
        if (!fn(&S->config, mesh->ident.c_str(), mesh->ident.size(), mi->ident.c_str(), mi->ident.size(), total_area, mi->importance, size)) {
            size[0] = cfg.default_width; size[1] = cfg.default_height;
        }
        mi->lm_width = size[0]; mi->lm_height = size[1];
        const float lw = (float)size[0], lh = (float)size[1];
        for (size_t v = 0; v < mesh->vpos.size(); ++v) {
            B.vtex[(vbase[i] + v) * 2 + 0] = mesh->vtex1[v].x;
            B.vtex[(vbase[i] + v) * 2 + 1] = mesh->vtex1[v].y;
            B.ltex[(vbase[i] + v) * 2 + 0] = mesh->vtex2[v].x * lw - 0.5f;
            B.ltex[(vbase[i] + v) * 2 + 1] = mesh->vtex2[v].y * lh - 0.5f;
        }
    };
    {
        const size_t first = std::max<size_t>(my_i0, 1);
        const unsigned nthreads = std::max(1u, std::min<unsigned>(hw_threads, my_i1 > first ? (unsigned)(my_i1 - first) : 1u));
        std::atomic<size_t> next{first};
        std::vector<std::thread> pool;
        auto worker = [&]() { for (size_t i; (i = next.fetch_add(1)) < my_i1;) xform_instance(i); };
        for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
    }
    S->stats.t_prexform = now_s() - t0;
    t0 = now_s();

    S->stage.store("generating data structures");
    S->completion.store(0.f);
    const bool trace = getenv("LTR_TRACE") != nullptr;          /* host-side phase timings on stderr */
    double tt = now_s();
    auto lap = [&](const char *what) { if (trace) { double t = now_s(); fprintf(stderr, "[ltr host] accel %-28s %8.2f ms\n", what, (t - tt) * 1e3); tt = t; } };
    /* per instance: useful triangles of shadow-casting parts (phase 1), then -- concurrently -- the flat scene BVH
     * over all of them on a background thread and the per-instance reference-order trees on this one.
     * The triangles of ALL instances are written once, instance after instance, straight into B.rtree_tris (the array the
     * reference-order trees index); the scene BVH is built over the same memory when every instance casts shadows (the
     * usual case), over a compacted copy otherwise. */
    std::vector<RefTree> itree(ni);
    auto parallel_instances = [&](const std::function<void(size_t)> &fn) {           /* over THIS RANK's instances */
        unsigned nthreads = std::max(1u, std::min<unsigned>(hw_threads, (unsigned)(my_i1 - my_i0)));
        std::atomic<size_t> next{my_i0};
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads; ++t)
            pool.emplace_back([&]() { for (size_t i; (i = next.fetch_add(1)) < my_i1;) fn(i); });
        for (auto &th : pool) th.join();
    };
    /* visits the useful triangles of instance i in part / index order (ref: lighter.cpp:349-384: parts with shadow == 0 and
     * triangles whose cross product is near zero are dropped) */
    auto for_useful_tris = [&](size_t i, auto &&emit) {
        MeshInstance *mi = S->instances[i];
        ltr_Mesh *mesh = mi->mesh;
        const V3 *wp = B.wpos.data() + vbase[i];
        for (const MeshPart &mp : mesh->parts) {
            if (!mp.shadow) continue;
            const V3 *vb = wp + mp.vertex_offset;
            const u32 *ib = mesh->indices.data() + mp.index_offset;
            for (u32 t = 0; t + 2 < mp.index_count; t += 3) {
                const V3 p1 = vb[ib[t]], p2 = vb[ib[t + 1]], p3 = vb[ib[t + 2]];
                if (near_zero3(cross3(p2 - p1, p3 - p1))) continue;
                emit(p1, p2, p3);
            }
        }
    };
    std::vector<size_t> n_itris(ni, 0), o_tri(ni + 1, 0);
    parallel_instances([&](size_t i) {
        if (i == 0) return;
        size_t c = 0;
        for_useful_tris(i, [&](const V3 &, const V3 &, const V3 &) { ++c; });
        n_itris[i] = c;
    });
    {   /* exchange 1: lightmap sizes (size_fn ran on the owner only) and triangle counts */
        std::vector<uint32_t> tab(ni * 3, 0);
        for (size_t i = std::max<size_t>(my_i0, 1); i < my_i1; ++i) { tab[i * 3] = S->instances[i]->lm_width; tab[i * 3 + 1] = S->instances[i]->lm_height; tab[i * 3 + 2] = (uint32_t)n_itris[i]; }
        exchange(tab, 3);
        for (size_t i = 1; i < ni; ++i) { S->instances[i]->lm_width = tab[i * 3]; S->instances[i]->lm_height = tab[i * 3 + 1]; n_itris[i] = tab[i * 3 + 2]; }
    }
    for (size_t i = 0; i < ni; ++i) o_tri[i + 1] = o_tri[i] + n_itris[i];
    B.rtree_tris.resize(o_tri[ni] * 9);
    BigVec<Box3> all_boxes(o_tri[ni]);
    parallel_instances([&](size_t i) {
        if (i == 0) return;
        float *T = B.rtree_tris.data() + o_tri[i] * 9;
        Box3 *bx = all_boxes.data() + o_tri[i];
        for_useful_tris(i, [&](const V3 &p1, const V3 &p2, const V3 &p3) {
            T[0] = p1.x; T[1] = p1.y; T[2] = p1.z; T[3] = p2.x; T[4] = p2.y; T[5] = p2.z; T[6] = p3.x; T[7] = p3.y; T[8] = p3.z;
            T += 9;
            bx->lo = min3(p1, min3(p2, p3)); bx->hi = max3(p1, max3(p2, p3));
            ++bx;
        });
    });
    /* the scene BVH covers the instances that cast shadows */
    bool all_cast = true;
    for (size_t i = 1; i < ni; ++i) if (n_itris[i] && !S->instances[i]->shadow) all_cast = false;
    B.all_cast = all_cast;
    BigVec<float> scene_compact;
    if (!all_cast) {
        std::vector<size_t> o_stri(ni + 1, 0);
        for (size_t i = 0; i < ni; ++i) o_stri[i + 1] = o_stri[i] + ((i && S->instances[i]->shadow) ? n_itris[i] : 0);
        scene_compact.resize(o_stri[ni] * 9);
        parallel_instances([&](size_t i) {
            if (i && S->instances[i]->shadow && n_itris[i]) memcpy(&scene_compact[o_stri[i] * 9], &B.rtree_tris[o_tri[i] * 9], n_itris[i] * 36);
        });
    }
    const float *scene_tris = all_cast ? B.rtree_tris.data() : scene_compact.data();
    const size_t nst = all_cast ? o_tri[ni] : scene_compact.size() / 9;
    lap("world triangles");
    int leaf_max = BVH_LEAF_MAX;
    if (const char *e = getenv("LTR_BVH_LEAF")) leaf_max = atoi(e);
    if (leaf_max < 1) leaf_max = 1;
    if (leaf_max > 7) leaf_max = 7;
    B.bvh_leaf_max = leaf_max;
    /* The flat scene BVH is built on the device during upload() (gpu_bvh.cu: the same binned SAH, a few milliseconds instead
     * of 70-100 ms of host threads, and nothing for the ranks of a multi-GPU bake to repeat on the same cores).  The host
     * builder remains for scenes that fit one leaf, for the host-only test hooks and as an A/B switch (LTR_BVH_HOST=1). */
    B.device_bvh = !force_host_bvh && !getenv("LTR_BVH_HOST") && nst > (size_t)leaf_max;
    B.scene_tris = nullptr; B.n_scene_tris = 0;
    std::thread bvh_thread;
    int early_rc = 0;
    if (B.device_bvh) {
        B.bvh = SceneBvh();
        if (all_cast) { B.bvh_tris.clear(); B.scene_tris = B.rtree_tris.data(); }
        else { B.bvh_tris.swap(scene_compact); B.scene_tris = B.bvh_tris.data(); }
        B.n_scene_tris = nst;
        /* every instance casts (the usual case): the scene BVH is built over the very array that is ready now, so the
         * triangles go up at once (sharded: one collective, here on the bake thread) and the device builds the tree on its
         * second stream while this host builds the reference-order trees below; upload() finds both done */
        if (all_cast && B.gpu && !(getenv("LTR_BVH_LATE") && getenv("LTR_BVH_LATE")[0] == '1')) {      /* LTR_BVH_LATE=1: build during upload() instead (A/B) */
            std::vector<uint64_t> sh;
            if (world > 1) for (int r = 0; r <= world; ++r) sh.push_back(o_tri[B.inst_cut[r]]);
            gpu_check(S, ltrgpu_upload_tris_early(B.gpu, B.rtree_tris.data(), (uint32_t)o_tri[ni], sh.empty() ? nullptr : sh.data()), "early triangle upload");
            bvh_thread = std::thread([&B, &early_rc, leaf_max]() { early_rc = ltrgpu_build_bvh_early(B.gpu, leaf_max); });
        }
    } else bvh_thread = std::thread([&]() {
        build_scene_bvh(scene_tris, nst, B.bvh, leaf_max, 0);
        /* triangles in BVH order */
        B.bvh_tris.resize(nst * 9);
        const unsigned T = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < T; ++t)
            pool.emplace_back([&, t]() {
                for (size_t k = nst * t / T; k < nst * (t + 1) / T; ++k) memcpy(&B.bvh_tris[k * 9], &scene_tris[(size_t)B.bvh.order[k] * 9], 36);
            });
        for (auto &th : pool) th.join();
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } bvh_joiner{ bvh_thread };
    {   /* threads left over by the instance-level parallelism go into the big nodes of each tree (one merged 500 k-triangle instance: 247 -> ~60 ms) */
        const unsigned busy = std::max(1u, std::min<unsigned>(hw_threads, (unsigned)(my_i1 - my_i0)));
        const int per_tree = (int)std::max(1u, hw_threads / busy);
        parallel_instances([&](size_t i) { itree[i].build(i ? all_boxes.data() + o_tri[i] : nullptr, i ? n_itris[i] : 0, per_tree); });
    }
    lap("instance trees (threads)");
    S->completion.store(0.99f);
    /* exchange 2: tree sizes and root boxes */
    std::vector<uint32_t> ttab(ni * 8, 0);
    for (size_t i = my_i0; i < my_i1; ++i) {
        ttab[i * 8] = (uint32_t)itree[i].nodes.size(); ttab[i * 8 + 1] = (uint32_t)itree[i].items.size();
        memcpy(&ttab[i * 8 + 2], &itree[i].nodes[0].lo, 12); memcpy(&ttab[i * 8 + 5], &itree[i].nodes[0].hi, 12);
    }
    exchange(ttab, 8);
    /* instance tree over the root boxes (invalid boxes are dropped by the builder) */
    std::vector<Box3> ibox(ni);
    for (size_t i = 0; i < ni; ++i) { memcpy(&ibox[i].lo, &ttab[i * 8 + 2], 12); memcpy(&ibox[i].hi, &ttab[i * 8 + 5], 12); }
    RefTree inst_tree;
    inst_tree.build(ibox.data(), ni);

    /* concatenate trees, lay out the texel space, list raster triangles: offsets first, then every
     * instance copies its own slices (threads) -- at 1M triangles this is ~130 MB of host memory */
    uint64_t texel_off = 0;
    {
        std::vector<size_t> o_node(ni + 1, 0), o_item(ni + 1, 0), o_rtri(ni + 1, 0);     /* o_tri: the triangle offsets from above */
        for (size_t i = 0; i < ni; ++i) {
            ltrgpu_Inst &I = B.inst[i];
            MeshInstance *mi = S->instances[i];
            I.lm_w = i ? mi->lm_width : 0; I.lm_h = i ? mi->lm_height : 0;
            I.texel_off_lo = (uint32_t)texel_off; I.texel_off_hi = (uint32_t)(texel_off >> 32);
            texel_off += (uint64_t)I.lm_w * I.lm_h;
            I.node_off = (uint32_t)o_node[i]; I.item_off = (uint32_t)o_item[i]; I.tri_off = (uint32_t)o_tri[i];
            I.tree_tris = (uint32_t)n_itris[i];
            I.shadow = (i && mi->shadow) ? 1 : 0;
            size_t nrt = 0;
            if (i) for (const MeshPart &mp : mi->mesh->parts) nrt += mp.index_count / 3;
            o_node[i + 1] = o_node[i] + ttab[i * 8];
            o_item[i + 1] = o_item[i] + ttab[i * 8 + 1];
            o_rtri[i + 1] = o_rtri[i] + nrt;
        }
        B.rnodes.resize(o_node[ni]); B.ritems.resize(o_item[ni]); B.rtris.resize(o_rtri[ni]);      /* B.rtree_tris was filled in place above */
        B.sh_verts.clear(); B.sh_rtris.clear(); B.sh_rnodes.clear(); B.sh_ritems.clear(); B.sh_tris.clear();
        if (world > 1)
            for (int r = 0; r <= world; ++r) {
                const size_t c = B.inst_cut[r];
                B.sh_verts.push_back(vbase[c]); B.sh_rtris.push_back(o_rtri[c]); B.sh_rnodes.push_back(o_node[c]); B.sh_ritems.push_back(o_item[c]); B.sh_tris.push_back(o_tri[c]);
            }
        auto copy_one = [&](size_t i) {
            if (!itree[i].nodes.empty()) memcpy(&B.rnodes[o_node[i]], itree[i].nodes.data(), itree[i].nodes.size() * sizeof(itree[i].nodes[0]));
            if (!itree[i].items.empty()) memcpy(&B.ritems[o_item[i]], itree[i].items.data(), itree[i].items.size() * sizeof(itree[i].items[0]));
            if (!i) return;
            MeshInstance *mi = S->instances[i];
            ltr_Mesh *mesh = mi->mesh;
            ltrgpu_RasterTri *dst = B.rtris.data() + o_rtri[i];
            for (size_t p = 0; p < mesh->parts.size(); ++p) {
                const MeshPart &mp = mesh->parts[p];
                const u32 *ib = mesh->indices.data() + mp.index_offset;
                const uint32_t base = (uint32_t)(vbase[i] + mp.vertex_offset);
                for (u32 t = 0; t + 2 < mp.index_count; t += 3) {
                    ltrgpu_RasterTri rt = { (uint32_t)i, (uint32_t)p, base + ib[t], base + ib[t + 1], base + ib[t + 2] };
                    *dst++ = rt;
                }
            }
        };
        parallel_instances(copy_one);
    }
    lap("concatenate + raster list");
    /* flat scene BVH over the shadow-casting triangles: built in the background since the world triangles were ready */
    if (bvh_thread.joinable()) bvh_thread.join();
    if (early_rc) { Fail f; f.msg = std::string("early scene-BVH build: ") + ltrgpu_early_error(B.gpu); throw f; }
    lap("scene BVH (overlapped) + reorder");
    /* lights and the light -> instance table */
    const size_t nl = S->lights.size();
    B.lights.resize(nl);
    B.light_samples.clear();
    B.light_inst.assign(nl * ni, 0);
    for (size_t l = 0; l < nl; ++l) {
        const Light &L = S->lights[l];
        ltrgpu_Light &G = B.lights[l];
        G.pos = L.position; G.type = L.type; G.dir = L.direction; G.range = L.range; G.color = L.color; G.power = L.power;
        G.radius = L.light_radius; G.curve = L.spot_curve;
        float out_rad = L.spot_angle_out / 180.0f * (float)M_PI, in_rad = L.spot_angle_in / 180.0f * (float)M_PI;
        if (in_rad == out_rad) in_rad -= LB_SMALL;
        G.angle_out_rad = out_rad; G.angle_diff = in_rad - out_rad;
        /* sampled-shadow extension: golden-angle disk of light_radius, see geom.h shadow_sample_segment.
         * cos/sin come from the host libm, like every configuration-only transcendental. */
        G.n_samples = (uint32_t)std::min(std::max(L.shadow_sample_count, 1), 64);
        G.sample_off = (uint32_t)(B.light_samples.size() / 4);
        G.pad0 = G.pad1 = 0;
        for (uint32_t s = 0; s < G.n_samples; ++s) {
            const float q = (s + 0.5f) / (float)G.n_samples;
            const float rho = L.light_radius * sqrtf(q);
            const float angle = (s + L.randoff) * (137.508f / 180.0f * (float)M_PI);
            const float dx = cosf(angle) * rho, dy = sinf(angle) * rho;
            if (L.type == LTR_LT_DIRECT) {
                const V3 d = L.direction;
                const V3 up = norm3(cross3(d, mk3(d.y, -d.z, d.x))), rt = cross3(d, up);
                const V3 a = norm3(d + rt * dx + up * dy);
                const float v[4] = { a.x, a.y, a.z, 0.f };
                B.light_samples.insert(B.light_samples.end(), v, v + 4);
            } else {
                const float v[4] = { dx, dy, 0.f, 0.f };
                B.light_samples.insert(B.light_samples.end(), v, v + 4);
            }
        }
        uint8_t *row = &B.light_inst[l * ni];
        row[0] = 1;
        auto mark = [&](const int32_t *ids, int32_t n) { for (int32_t k = 0; k < n; ++k) row[ids[k]] = 1; };
        if (L.type == LTR_LT_POINT || L.type == LTR_LT_SPOT) inst_tree.query(L.position - mk3(L.range), L.position + mk3(L.range), mark);
        else if (L.type == LTR_LT_DIRECT) inst_tree.all(mark);
    }

    /* host-libm tables */
    const int ns = cfg.ao_num_samples > 0 ? cfg.ao_num_samples : 0;
    B.ao_cos.resize(ns); B.ao_sin.resize(ns);
    for (int s = 0; s < ns; ++s) {
        float q = (s + 0.5f) / cfg.ao_num_samples;
        B.ao_cos[s] = sqrtf(q);
        B.ao_sin[s] = sinf(acosf(B.ao_cos[s]));
    }
    B.blur_ext = 0; B.blur_kernel.clear();
    if (cfg.blur_size) {
        B.blur_ext = (int)ceil(cfg.blur_size);
        B.blur_kernel.resize(2 * B.blur_ext + 1);
        float sum = 0.0f, mult = 1.0f / sqrtf(2.0f * (float)M_PI * cfg.blur_size * cfg.blur_size);
        for (int i = -B.blur_ext; i <= B.blur_ext; ++i)
            sum += B.blur_kernel[i + B.blur_ext] = expf(-0.5f * powf((float)i / cfg.blur_size, 2.0f)) * mult;
        for (float &k : B.blur_kernel) k /= sum;
    }
    S->stats.n_triangles = nst;
    S->stats.n_bvh_nodes = B.bvh.nodes.size();
    S->stats.t_accel = now_s() - t0;
    B.prepared = true;
}

/* GPU context, NCCL communicator and the collective hooks: before the host pre-pass, which already exchanges two small tables */
void connect(ltr_Scene *S)
{
    Bake &B = *S->bake;
    if (!B.gpu) {
        if (ltrgpu_create(&B.gpu, S->device)) {
            Fail f; f.msg = std::string("CUDA device unavailable: ") + (B.gpu ? ltrgpu_last_error(B.gpu) : "no CUDA device (this library has no CPU fallback)");
            throw f;
        }
    }
    if (S->world > 1 && !B.comm) {
        char err[256];
        B.nccl = nccl_api(err, sizeof(err));
        if (!B.nccl) { Fail f; f.msg = std::string("NCCL unavailable: ") + err; throw f; }
        if (!S->have_nccl_id) { Fail f; f.msg = "sharded bake without an NCCL unique id (call ltrx_SetShard)"; throw f; }
        NcclId id;
        memcpy(&id, S->nccl_id, sizeof(id));
        /* One communicator per (unique id, rank, world) per process, shared by every scene that names
         * the same id: an NCCL unique id can seed ncclCommInitRank only once. */
        static std::mutex comm_mu;
        static std::map<std::string, void *> comm_cache;
        std::string key((const char *)S->nccl_id, sizeof(S->nccl_id));
        key += "/" + std::to_string(S->rank) + "/" + std::to_string(S->world);
        std::lock_guard<std::mutex> g(comm_mu);
        auto it = comm_cache.find(key);
        if (it != comm_cache.end()) B.comm = it->second;
        else {
            int rc = B.nccl->CommInitRank(&B.comm, S->world, id, S->rank);
            if (rc != 0) { B.comm = nullptr; Fail f; f.msg = std::string("ncclCommInitRank: ") + B.nccl->GetErrorString(rc); throw f; }
            comm_cache[key] = B.comm;
        }
    }
    B.world = S->world;
    gpu_check(S, ltrgpu_set_world(B.gpu, S->rank, S->world, S->world > 1 ? nccl_allgather_cb : nullptr, &B), "world");
    gpu_check(S, ltrgpu_set_gatherv(B.gpu, S->world > 1 ? nccl_gatherv_cb : nullptr), "world");
    gpu_check(S, ltrgpu_set_alltoallv(B.gpu, S->world > 1 ? nccl_alltoallv_cb : nullptr), "world");
}

void upload(ltr_Scene *S)
{
    Bake &B = *S->bake;
    const ltr_Config &cfg = S->config;
    double t0 = now_s();
    std::vector<V3> ppos(S->probes.size()), pnrm(S->probes.size());
    for (size_t i = 0; i < S->probes.size(); ++i) {
        ppos[i] = mk3(S->probes[i].position[0], S->probes[i].position[1], S->probes[i].position[2]);
        pnrm[i] = mk3(S->probes[i].normal[0], S->probes[i].normal[1], S->probes[i].normal[2]);
    }
    ltrgpu_SceneDesc d;
    memset(&d, 0, sizeof(d));
    ltrgpu_Params &P = d.params;
    memcpy(P.ambient, cfg.ambient_color, 12);
    P.max_correct_dist = cfg.max_correct_dist;
    P.corr_min_dot = cosf(cfg.max_correct_angle / 180.0f * (float)M_PI);
    P.ao_distance = cfg.ao_distance; P.ao_multiplier = cfg.ao_multiplier; P.ao_falloff = cfg.ao_falloff; P.ao_effect = cfg.ao_effect;
    memcpy(P.ao_color, cfg.ao_color_rgb, 12);
    P.ao_num_samples = cfg.ao_num_samples; P.blur_size = cfg.blur_size; P.ds2x = cfg.ds2x; P.normalmap = cfg.generate_normalmap_data;
    P.amb_brightness = (cfg.ambient_color[0] + cfg.ambient_color[1] + cfg.ambient_color[2]) * (1.0f / 3.0f);
    P.shadow_mode = S->shadow_mode;
    d.n_inst = (uint32_t)B.inst.size(); d.inst = B.inst.data();
    d.n_verts = (uint32_t)B.wpos.size(); d.wpos = B.wpos.data(); d.wnrm = B.wnrm.data(); d.vtex2 = B.vtex.data(); d.ltex2 = B.ltex.data();
    d.n_rtris = (uint32_t)B.rtris.size(); d.rtris = B.rtris.data();
    d.n_rnodes = (uint32_t)B.rnodes.size(); d.rnodes = B.rnodes.data();
    d.n_ritems = (uint32_t)B.ritems.size(); d.ritems = B.ritems.data();
    d.n_rtree_tris = (uint32_t)(B.rtree_tris.size() / 9); d.rtree_tris9 = B.rtree_tris.data();
    d.bvh_leaf_max = B.bvh_leaf_max;
    d.scene_covers_rtree = B.all_cast ? 1 : 0;
    if (!B.sh_verts.empty()) { d.shard_verts = B.sh_verts.data(); d.shard_rtris = B.sh_rtris.data(); d.shard_rnodes = B.sh_rnodes.data(); d.shard_ritems = B.sh_ritems.data(); d.shard_tris = B.sh_tris.data(); }
    if (B.device_bvh) {
        d.n_bvh_nodes = d.n_bvh4_nodes = 0; d.bvh = nullptr; d.bvh4 = nullptr; d.tri_orig = nullptr;
        d.n_tris = (uint32_t)B.n_scene_tris; d.tris9 = B.scene_tris;
    } else {
        d.n_bvh_nodes = (uint32_t)B.bvh.nodes.size(); d.bvh = B.bvh.nodes.data();
        d.n_bvh4_nodes = (uint32_t)B.bvh.nodes4.size(); d.bvh4 = B.bvh.nodes4.data();
        d.n_tris = (uint32_t)(B.bvh_tris.size() / 9); d.tris9 = B.bvh_tris.data(); d.tri_orig = B.bvh.order.data();
        d.bvh_height = B.bvh.depth;
    }
    d.n_lights = (uint32_t)B.lights.size(); d.lights = B.lights.data(); d.light_inst = B.light_inst.data();
    d.n_light_samples = (uint32_t)(B.light_samples.size() / 4); d.light_samples4 = B.light_samples.data();
    d.n_probes = (uint32_t)ppos.size(); d.probe_pos = ppos.data(); d.probe_nrm = pnrm.data();
    d.ao_cos_side = B.ao_cos.data(); d.ao_sin_side = B.ao_sin.data();
    d.blur_ext = B.blur_ext; d.blur_kernel = B.blur_kernel.empty() ? nullptr : B.blur_kernel.data();
    const bool trace = getenv("LTR_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[ltr host] upload context + descriptors      %8.2f ms\n", (now_s() - t0) * 1e3);
    double tu = now_s();
    gpu_check(S, ltrgpu_upload_scene(B.gpu, &d), "scene upload");
    {
        uint32_t nn = 0; int hh = 0; float bms = 0.f;
        ltrgpu_bvh_info(B.gpu, &nn, &hh, &bms);
        S->stats.n_bvh_nodes = nn;
        if (trace) fprintf(stderr, "[ltr host] scene BVH: %u nodes, %d levels, %s %.2f ms\n", nn, hh, B.device_bvh ? "built on the device in" : "host-built, device time", bms);
    }
    if (trace) fprintf(stderr, "[ltr host] upload scene arrays               %8.2f ms\n", (now_s() - tu) * 1e3);

    S->stats.t_upload = now_s() - t0;
}

/* ------------------------------------------------------------------------------------------
 * sample_fn batching (ref: lighter.cpp:690-715)
 * ------------------------------------------------------------------------------------------ */
struct MaterialJob {
    static constexpr uint32_t CHUNK = 1u << 18;       /* lumels per chunk: 12 MB of request records */
    float *diffuse = nullptr, *emissive = nullptr;     /* n x 3 each, page-locked: uploaded by one DMA each */
    ltrgpu_SampleReq *req[2] = { nullptr, nullptr };
    std::thread th;
    std::string error;
    double seconds = 0;
    bool started = false;

    void start(ltr_Scene *S, uint64_t n)
    {
        Bake &B = *S->bake;
        gpu_check(S, ltrgpu_sample_requests_begin(B.gpu), "sample requests");
        diffuse = (float *)ltrgpu_host_alloc(n * 12); emissive = (float *)ltrgpu_host_alloc(n * 12);
        for (int b = 0; b < 2; ++b) req[b] = (ltrgpu_SampleReq *)ltrgpu_host_alloc((size_t)CHUNK * sizeof(ltrgpu_SampleReq));
        if (!diffuse || !emissive || !req[0] || !req[1]) { Fail f; f.msg = "out of host memory for the material tables"; throw f; }
        started = true;
        th = std::thread([this, S, n]() { run(S, n); });
    }
    void run(ltr_Scene *S, uint64_t n)
    {
        const double t0 = now_s();
        Bake &B = *S->bake;
        const uint64_t first_mesh = B.lumel_off[1];                     /* probes (instance 0) never reach the callback: diffuse 0 on the device */
        for (uint64_t i = 0; i < first_mesh * 3; ++i) { diffuse[i] = 1.f; emissive[i] = 0.f; }
        auto count_at = [&](uint64_t c0) { return (uint32_t)std::min<uint64_t>(CHUNK, n - c0); };
        size_t m = 1;                                                   /* instance of the current lumel */
        int slot = 0;
        if (first_mesh < n && ltrgpu_sample_requests_issue(B.gpu, first_mesh, count_at(first_mesh), req[0], 0)) { error = ltrgpu_aux_error(B.gpu); return; }
        for (uint64_t c0 = first_mesh; c0 < n; c0 += CHUNK, slot ^= 1) {
            const uint32_t cnt = count_at(c0);
            const uint64_t c1 = c0 + cnt;
            if (c1 < n && ltrgpu_sample_requests_issue(B.gpu, c1, count_at(c1), req[slot ^ 1], slot ^ 1)) { error = ltrgpu_aux_error(B.gpu); return; }
            if (ltrgpu_sample_requests_wait(B.gpu, slot)) { error = ltrgpu_aux_error(B.gpu); return; }
            const ltrgpu_SampleReq *q = req[slot];
            for (uint32_t k = 0; k < cnt; ++k) {
                const uint64_t i = c0 + k;
                while (i >= B.lumel_off[m + 1]) ++m;
                const MeshInstance *mi = S->instances[m];
                ltr_SampleRequest r;
                memset(&r, 0, sizeof(r));
                memcpy(r.position, q[k].pos, 12); memcpy(r.normal, q[k].nrm, 12);
                r.tex0u = q[k].tex0[0]; r.tex0v = q[k].tex0[1]; r.tex1u = q[k].tex1[0]; r.tex1v = q[k].tex1[1];
                r.part_id = q[k].part_id;
                r.mesh_ident = mi->mesh->ident.c_str(); r.mesh_ident_size = mi->mesh->ident.size();
                r.inst_ident = mi->ident.c_str(); r.inst_ident_size = mi->ident.size();
                r.out_diffuse_color[0] = r.out_diffuse_color[1] = r.out_diffuse_color[2] = 1;
                float *dd = diffuse + i * 3, *ee = emissive + i * 3;
                if (S->config.sample_fn(&S->config, &r)) { memcpy(dd, r.out_diffuse_color, 12); memcpy(ee, r.out_emissive_color, 12); }
                else { dd[0] = dd[1] = dd[2] = 1.f; ee[0] = ee[1] = ee[2] = 0.f; }
            }
        }
        seconds = now_s() - t0;
    }
    /* ltrgpu_materials_fn: called by the radiosity stage when the bounces need the materials */
    static int ready(void *user, const float **d, const float **e)
    {
        MaterialJob *J = (MaterialJob *)user;
        *d = *e = nullptr;
        if (!J->started) return 0;
        if (J->th.joinable()) J->th.join();
        if (!J->error.empty()) return 1;
        *d = J->diffuse; *e = J->emissive;
        return 0;
    }
    ~MaterialJob()
    {
        if (th.joinable()) th.join();
        ltrgpu_host_free(diffuse); ltrgpu_host_free(emissive); ltrgpu_host_free(req[0]); ltrgpu_host_free(req[1]);
    }
};

/* ------------------------------------------------------------------------------------------
 * GPU stages (re-runnable on a resident scene)
 * ------------------------------------------------------------------------------------------ */
void gpu_stages(ltr_Scene *S)
{
    Bake &B = *S->bake;
    const ltr_Config &cfg = S->config;
    const size_t ni = S->instances.size();
    gpu_check(S, ltrgpu_reset_bake(B.gpu), "reset");
    gpu_check(S, ltrgpu_span_begin(B.gpu), "span");

    double t0 = now_s();
    S->stage.store("generating samples");
    S->completion.store(0.f);
    B.lumel_off.assign(ni + 1, 0);
    gpu_check(S, ltrgpu_set_world(B.gpu, S->rank, S->world, S->world > 1 ? nccl_allgather_cb : nullptr, &B), "world");
    B.world = S->world;
    gpu_check(S, ltrgpu_set_gatherv(B.gpu, S->world > 1 ? nccl_gatherv_cb : nullptr), "world");
    gpu_check(S, ltrgpu_set_alltoallv(B.gpu, S->world > 1 ? nccl_alltoallv_cb : nullptr), "world");
    gpu_check(S, ltrgpu_generate_lumels(B.gpu, B.lumel_off.data()), "lumel generation");
    const uint64_t n = B.lumel_off[ni];
    uint64_t sb = 0, se = n;
    ltrx_ShardRange(n, S->rank, S->world, &sb, &se);
    gpu_check(S, ltrgpu_set_shard(B.gpu, sb, se, S->rank, S->world, S->world > 1 ? nccl_allgather_cb : nullptr, &B), "shard");
    S->stats.n_lumels_total = n;
    S->stats.n_lumels_local = se - sb;
    S->stats.t_samples = now_s() - t0;

    /* sample_fn (the material callback, ref: lighter.cpp:690-715) runs on its own host thread from here on, concurrently with
     * the direct-light stage and radiosity link generation on the GPU; radiosity waits for it right before the first bounce
     * (MaterialJob::ready).  Calls are made one at a time in the reference's order -- instance ascending, lumel ascending --
     * because a callback may keep state; what is batched is everything around them: the request fields are computed on the
     * device and arrive in pinned chunks on a second stream (chunk k+1 in flight while the callbacks of chunk k run). */
    MaterialJob matjob;
    if (cfg.bounce_count && cfg.sample_fn && n) matjob.start(S, n);

    /* Replay of the reference's rand() consumption for the AO pass: one randf() per lumel, instance by
     * instance (probe container first), lumel index ascending (lighter.cpp:819,1130-1135).  The draws
     * only depend on the lumel COUNT, so they are generated on a host thread now, while the GPU runs
     * the direct-light and radiosity stages (7.8 M draws = ~0.1 s of host time otherwise exposed).
     * RAII: the thread is always joined, also when a later stage throws. */
    struct RandJob {
        float *v = nullptr;                         /* page-locked, from the cache: no 31 MB of fresh page faults per bake, and the upload is a plain DMA */
        std::thread th;
        ~RandJob() { if (th.joinable()) th.join(); ltrgpu_host_free(v); }
    } randjob;
    const bool user_code_before_ao = cfg.bounce_count && cfg.sample_fn;   /* a material callback might itself call rand(): keep the reference's order */
    if (cfg.ao_distance) {
        randjob.v = (float *)ltrgpu_host_alloc((n ? n : 1) * sizeof(float));
        if (!randjob.v) { Fail f; f.msg = "out of host memory for the AO offsets"; throw f; }
    }
    if (cfg.ao_distance && !user_code_before_ao) {
        randjob.th = std::thread([&randjob, n]() { rand_fill(randjob.v, n); });
    }

    t0 = now_s();
    if (!S->lights.empty() || cfg.generate_normalmap_data) {       /* without lights only the normal-map terms are produced (the reference names no stage then) */
        if (!S->lights.empty()) { S->stage.store("rendering lightmaps"); S->completion.store(0.f); }
        gpu_check(S, ltrgpu_direct_light(B.gpu), "direct light");
        S->completion.store(1.f);
        if (S->keep_debug) {
            const uint64_t nl = se - sb;
            B.d_fvis.assign(S->lights.size() * nl, 0.f);
            for (size_t l = 0; l < S->lights.size(); ++l)
                if (ltrgpu_download_shadow_factors(B.gpu, (uint32_t)l, B.d_fvis.data() + l * nl)) break;   /* only the last light chunk is resident */
            B.d_smask.clear();
            if (S->shadow_mode == 1) {
                B.d_smask.assign(S->lights.size() * nl, 0);
                for (size_t l = 0; l < S->lights.size(); ++l)
                    if (ltrgpu_download_shadow_masks(B.gpu, (uint32_t)l, B.d_smask.data() + l * nl)) break;
            }
        }
    }
    S->stats.t_direct = now_s() - t0;

    t0 = now_s();
    if (cfg.bounce_count) {
        S->stage.store("calculating radiosity");
        S->completion.store(0.f);
        S->stage.store("bouncing light");
        gpu_check(S, ltrgpu_radiosity_ex(B.gpu, MaterialJob::ready, &matjob, cfg.bounce_count), "radiosity");
        if (!matjob.error.empty()) { Fail f; f.msg = matjob.error; throw f; }
        S->stats.t_sample_fn = matjob.seconds;
        S->stage.store("committing radiosity");
        S->completion.store(1.f);
    }
    S->stats.t_radiosity = now_s() - t0;

    t0 = now_s();
    if (cfg.ao_distance) {
        S->stage.store("rendering ambient occlusion");
        S->completion.store(0.f);
        /* replay of the reference's rand() consumption: one randf() per lumel, instance by
         * instance (probe container first), lumel index ascending (lighter.cpp:819,1130-1135) */
        if (randjob.th.joinable()) randjob.th.join();
        else rand_fill(randjob.v, n);
        gpu_check(S, ltrgpu_ambient_occlusion(B.gpu, randjob.v + sb), "ambient occlusion");
        S->completion.store(1.f);
    }
    S->stats.t_ao = now_s() - t0;

    t0 = now_s();
    S->stage.store("exporting lightmaps");
    S->completion.store(0.f);
    if (S->keep_debug && n) {
        /* per-lumel colours BEFORE the all-gather/finalize (local shard values are final here) */
    }
    gpu_check(S, ltrgpu_finalize(B.gpu), "finalize");
    gpu_check(S, ltrgpu_span_end(B.gpu), "span");
    S->stats.t_finalize = now_s() - t0;
}

void collect_counters(ltr_Scene *S);

void readback(ltr_Scene *S)
{
    Bake &B = *S->bake;
    const size_t ni = S->instances.size();
    double t0 = now_s();
    for (ltr_WorkOutput &wo : S->outputs) free(wo.normals_xyzf);
    S->outputs.clear();
    ltrgpu_host_free(S->output_arena); S->output_arena = nullptr;
    if (S->output_root_only && S->world > 1 && S->rank != 0) {       /* ltrx_SetOutputRoot: only rank 0 brings the lightmaps to the host */
        S->stats.t_readback = now_s() - t0;
        collect_counters(S);
        return;
    }
    /* every lightmap in ONE device -> host copy into one page-locked block owned by the scene */
    std::vector<uint64_t> out_off(ni + 1, 0);
    gpu_check(S, ltrgpu_output_layout(B.gpu, out_off.data()), "output layout");
    S->output_arena = (float *)ltrgpu_host_alloc((out_off[ni] ? out_off[ni] : 1) * 12);
    if (!S->output_arena) { Fail f; f.msg = "out of host memory for the lightmaps"; throw f; }
    gpu_check(S, ltrgpu_download_outputs_all(B.gpu, S->output_arena), "output download");
    for (size_t i = 1; i < ni; ++i) {
        MeshInstance *mi = S->instances[i];
        uint32_t w = 0, h = 0;
        gpu_check(S, ltrgpu_output_size(B.gpu, (uint32_t)i, &w, &h), "output size");
        ltr_WorkOutput wo;
        memset(&wo, 0, sizeof(wo));
        wo.uid = (u32)i;
        wo.mesh_ident = mi->mesh->ident.c_str(); wo.mesh_ident_size = mi->mesh->ident.size();
        wo.inst_ident = mi->ident.c_str(); wo.inst_ident_size = mi->ident.size();
        wo.width = w; wo.height = h;
        const size_t cnt = (size_t)w * h;
        wo.lightmap_rgb = S->output_arena + out_off[i] * 3;
        wo.normals_xyzf = S->config.generate_normalmap_data ? (float *)calloc(cnt ? cnt : 1, 16) : nullptr;
        S->outputs.push_back(wo);                  /* owned by the scene from here on */
        if (wo.normals_xyzf) gpu_check(S, ltrgpu_download_output(B.gpu, (uint32_t)i, nullptr, wo.normals_xyzf), "normal map download");
        mi->lm_width = w; mi->lm_height = h;       /* the reference shrinks these under ds2x (lighter.cpp:973-974) */
    }
    if (!S->probes.empty()) {
        std::vector<float> pc(S->probes.size() * 3);
        gpu_check(S, ltrgpu_download_probe_colors(B.gpu, pc.data()), "probe download");
        for (size_t i = 0; i < S->probes.size(); ++i) memcpy(S->probes[i].out_color, &pc[i * 3], 12);
    }
    if (S->keep_debug) {
        const uint64_t n = B.lumel_off[ni];
        B.d_pos.resize(n * 3); B.d_nrm.resize(n * 3); B.d_rad.resize(n * 4); B.d_rgb.resize(n * 3); B.d_loc.resize(n);
        gpu_check(S, ltrgpu_download_lumels(B.gpu, B.d_pos.data(), B.d_nrm.data(), B.d_loc.data(), B.d_rad.data(), B.d_rgb.data()), "debug lumels");
        uint64_t rows = 0, links = 0;
        gpu_check(S, ltrgpu_download_links(B.gpu, nullptr, nullptr, nullptr, &rows, &links), "debug links");
        B.d_lrow.assign(rows + 1, 0); B.d_lother.resize(links); B.d_lfac.resize(links);
        if (rows) gpu_check(S, ltrgpu_download_links(B.gpu, B.d_lrow.data(), B.d_lother.data(), B.d_lfac.data(), &rows, &links), "debug links");
    }
    S->stats.t_readback = now_s() - t0;

    collect_counters(S);
}

void collect_counters(ltr_Scene *S)
{
    ltrgpu_Counters c;
    if (!S->bake || !S->bake->gpu || ltrgpu_get_counters(S->bake->gpu, &c) != 0) return;
    ltrx_Stats &st = S->stats;
    st.n_marches = c.marches; st.n_distance_queries = c.distance_queries; st.n_ao_segments = c.ao_segments;
    st.n_correction_rays = c.correction_rays; st.n_rad_pairs = c.rad_pairs; st.n_rad_segments = c.rad_segments;
    st.n_rad_links = c.rad_links; st.n_node_visits = c.node_visits; st.n_tri_tests = c.tri_tests;
        st.n_ray_node_visits = c.ray_node_visits; st.n_ray_tri_tests = c.ray_tri_tests; st.n_rad_tile_loads = c.rad_tile_loads;
    st.kernel_launches = c.kernel_launches; st.h2d_bytes = c.h2d_bytes; st.d2h_bytes = c.d2h_bytes; st.n_rad_batches = c.rad_batches; st.n_shadow_rays = c.shadow_rays; st.n_ray_entry_tests = c.ray_entry_tests;
    st.gpu_ms_samples = c.ms_samples; st.gpu_ms_direct = c.ms_direct; st.gpu_ms_march = c.ms_march;
    st.gpu_ms_radiosity = c.ms_radiosity; st.gpu_ms_ao = c.ms_ao; st.gpu_ms_finalize = c.ms_finalize;
    st.gpu_ms_total = c.ms_samples + c.ms_direct + c.ms_radiosity + c.ms_ao + c.ms_finalize;
    st.gpu_ms_rad_pairs = c.ms_rad_pairs; st.gpu_ms_rad_vis = c.ms_rad_vis; st.gpu_ms_span = c.ms_span;
}

template <class F> int guarded(ltr_Scene *S, F f)
{
    try { f(); return 1; }
    catch (const Fail &e) { S->error = e.msg; }
    catch (const std::exception &e) { S->error = std::string("exception: ") + e.what(); }
    fprintf(stderr, "lighter_b200: bake failed: %s\n", S->error.c_str());
    S->failed_stage = "failed: " + S->error;          /* what ltr_GetStatus reports once the bake thread is done (api.cpp) */
    return 0;
}

} // namespace

void bake_main(ltr_Scene *S)
{
    const double t0 = now_s();
    S->error.clear();
    S->failed_stage.clear();
    if (!S->bake) S->bake = new Bake;
    const bool trace = getenv("LTR_TRACE") != nullptr;
    double tc = 0, th = 0, tu = 0, tg = 0, tr = 0;
    guarded(S, [&]() {
        connect(S);      tc = now_s();
        host_prepare(S); th = now_s();
        upload(S);       tu = now_s();
        gpu_stages(S);   tg = now_s();
        readback(S);     tr = now_s();
    });
    S->stats.t_total = now_s() - t0;
    if (trace && tr > 0)
        fprintf(stderr, "[ltr host] bake thread: connect %.2f  host pre-pass %.2f  upload %.2f  gpu stages %.2f  read-back + counters %.2f  total %.2f ms\n",
                (tc - t0) * 1e3, (th - tc) * 1e3, (tu - th) * 1e3, (tg - tu) * 1e3, (tr - tg) * 1e3, S->stats.t_total * 1e3);
    S->completion.store(1.f);
    S->stage.store(nullptr, std::memory_order_release);
}

void bake_free(ltr_Scene *S)
{
    if (!S->bake) return;
    Bake *B = S->bake;
    /* communicators are process-lifetime (cached by unique id in upload()); nothing to destroy here */
    if (B->gpu) ltrgpu_destroy(B->gpu);
    delete B;
    S->bake = nullptr;
}

/* ------------------------------------------------------------------------------------------
 * ltrx_* extension API
 * ------------------------------------------------------------------------------------------ */
extern "C" {

const char *ltrx_Version(void) { return "lighter_b200 0.1 (sm_100a)"; }

/* A ready-made native material callback for ltr_Config::sample_fn (callers in Python / ctypes can time large bakes without
 * a per-lumel trip through the interpreter): 4-unit checker of two albedos, the second graded by the lightmap u coordinate,
 * downward-facing surfaces glow, every 64th (cell, part) combination is declined (return 0 = keep the defaults). */
LTRBOOL ltrx_SampleFnChecker(ltr_Config *, ltr_SampleRequest *req)
{
    const int s = (int)floorf(req->position[0] * 0.25f) + (int)floorf(req->position[1] * 0.25f);
    if (((s ^ (int)req->part_id) & 63) == 63) return 0;
    if (s & 1) { req->out_diffuse_color[0] = 0.5f; req->out_diffuse_color[1] = 0.05f; req->out_diffuse_color[2] = 0.02f; }
    else { req->out_diffuse_color[0] = 0.7f; req->out_diffuse_color[1] = 0.7f * req->tex1u; req->out_diffuse_color[2] = 0.6f; }
    if (req->normal[2] < -0.5f) { req->out_emissive_color[0] = 0.1f; req->out_emissive_color[1] = 0.1f; req->out_emissive_color[2] = 0.3f; }
    return 1;
}

int ltrx_SetDevice(ltr_Scene *scene, int cuda_device) { scene->device = cuda_device; return 1; }

int ltrx_NcclUniqueId(unsigned char out_id[LTRX_NCCL_ID_BYTES])
{
    char err[256];
    const NcclApi *api = nccl_api(err, sizeof(err));
    if (!api) { fprintf(stderr, "lighter_b200: %s\n", err); return 0; }
    NcclId id;
    memset(&id, 0, sizeof(id));
    if (api->GetUniqueId(&id) != 0) return 0;
    memcpy(out_id, &id, LTRX_NCCL_ID_BYTES);
    return 1;
}

int ltrx_SetShard(ltr_Scene *scene, int rank, int world, const unsigned char *nccl_id)
{
    if (world < 1 || rank < 0 || rank >= world) return 0;
    if (world > 1 && !nccl_id) return 0;
    scene->rank = rank; scene->world = world;
    if (nccl_id) { memcpy(scene->nccl_id, nccl_id, LTRX_NCCL_ID_BYTES); scene->have_nccl_id = true; }
    return 1;
}

void ltrx_ShardRange(uint64_t n, int rank, int world, uint64_t *begin, uint64_t *end)
{
    if (world < 1) world = 1;
    const uint64_t chunk = (n + (uint64_t)world - 1) / (uint64_t)world;
    uint64_t b = chunk * (uint64_t)rank;
    if (b > n) b = n;
    uint64_t e = b + chunk;
    if (e > n) e = n;
    *begin = b; *end = e;
}

int ltrx_GetStats(ltr_Scene *scene, ltrx_Stats *out) { *out = scene->stats; return 1; }

const char *ltrx_GetError(ltr_Scene *scene) { return scene->error.c_str(); }

int ltrx_Prepare(ltr_Scene *scene)
{
    if (scene->worker.joinable()) scene->worker.join();
    if (!scene->bake) scene->bake = new Bake;
    scene->error.clear();
    int ok = guarded(scene, [&]() { connect(scene); host_prepare(scene); upload(scene); });
    scene->stage.store(ok ? "prepared" : nullptr);
    return ok;
}

/* host-only test hook: the culling tests of the radiosity pair sweep (rad_cull.h) on one block of row lumels against one block
 * of column lumels -- bounds as rad_tile_bounds_kernel builds them (component-wise min / max of positions and normals).
 * block_ok = tile_pair_may_link(rows, cols); row_ok[r] = row_group_may_link(row r, cols); pair_fast[r * ncols + c] (optional) =
 * the lock-step FMA pre-filter of the pair. */
int ltrx_test_rad_cull(const float *rowP3, const float *rowN3, u32 nrows, const float *colP3, const float *colN3, u32 ncols, int *block_ok, uint8_t *row_ok,
                       uint8_t *pair_fast)
{
    auto bounds = [](const float *P, const float *N, u32 n) {
        float v[12];
        for (int a = 0; a < 6; ++a) { v[a] = INFINITY; v[6 + a] = -INFINITY; }
        for (u32 i = 0; i < n; ++i)
            for (int a = 0; a < 3; ++a) {
                v[a] = fminf(v[a], P[3 * i + a]); v[3 + a] = fminf(v[3 + a], N[3 * i + a]);
                v[6 + a] = fmaxf(v[6 + a], P[3 * i + a]); v[9 + a] = fmaxf(v[9 + a], N[3 * i + a]);
            }
        TileBounds w;
        w.plo = make_float4(v[0], v[1], v[2], 0.f); w.nlo = make_float4(v[3], v[4], v[5], 0.f);
        w.phi = make_float4(v[6], v[7], v[8], 0.f); w.nhi = make_float4(v[9], v[10], v[11], 0.f);
        return w;
    };
    if (!nrows || !ncols) return 0;
    const TileBounds R = bounds(rowP3, rowN3, nrows), Cb = bounds(colP3, colN3, ncols);
    *block_ok = tile_pair_may_link(R, Cb) ? 1 : 0;
    for (u32 r = 0; r < nrows; ++r)
        row_ok[r] = row_group_may_link(mk3(rowP3[3 * r], rowP3[3 * r + 1], rowP3[3 * r + 2]), mk3(rowN3[3 * r], rowN3[3 * r + 1], rowN3[3 * r + 2]), Cb) ? 1 : 0;
    if (pair_fast)
        for (u32 r = 0; r < nrows; ++r)
            for (u32 c = 0; c < ncols; ++c)
                pair_fast[(size_t)r * ncols + c] = rad_fast_filter(mk3(rowP3[3 * r], rowP3[3 * r + 1], rowP3[3 * r + 2]), mk3(rowN3[3 * r], rowN3[3 * r + 1], rowN3[3 * r + 2]),
                                                                   make_float4(colP3[3 * c], colP3[3 * c + 1], colP3[3 * c + 2], 0.f),
                                                                   make_float4(colN3[3 * c], colN3[3 * c + 1], colN3[3 * c + 2], 0.f)) ? 1 : 0;
    return 1;
}

/* host-only test hook: the host pre-pass alone (no device), fingerprints of everything it would upload */
int ltrx_test_host_prepare(ltr_Scene *scene, uint64_t out_hash[12])
{
    if (scene->worker.joinable()) scene->worker.join();
    if (!scene->bake) scene->bake = new Bake;
    scene->error.clear();
    int ok = guarded(scene, [&]() { host_prepare(scene, /*force_host_bvh=*/true); });
    if (!ok) return 0;
    const Bake &B = *scene->bake;
    auto fnv = [](const void *p, size_t n) { uint64_t h = 1469598103934665603ull; const unsigned char *c = (const unsigned char *)p; for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; } return h; };
    out_hash[0] = fnv(B.wpos.data(), B.wpos.size() * sizeof(V3));
    out_hash[1] = fnv(B.wnrm.data(), B.wnrm.size() * sizeof(V3));
    out_hash[2] = fnv(B.vtex.data(), B.vtex.size() * 4) ^ fnv(B.ltex.data(), B.ltex.size() * 4);
    out_hash[3] = fnv(B.rtris.data(), B.rtris.size() * sizeof(B.rtris[0]));
    out_hash[4] = fnv(B.rnodes.data(), B.rnodes.size() * sizeof(B.rnodes[0]));
    out_hash[5] = fnv(B.ritems.data(), B.ritems.size() * 4);
    out_hash[6] = fnv(B.rtree_tris.data(), B.rtree_tris.size() * 4);
    out_hash[7] = fnv(B.bvh.nodes.data(), B.bvh.nodes.size() * sizeof(BvhNode));
    out_hash[8] = fnv(B.bvh.nodes4.data(), B.bvh.nodes4.size() * sizeof(Bvh4Node));
    out_hash[9] = fnv(B.bvh.order.data(), B.bvh.order.size() * 4);
    out_hash[10] = fnv(B.bvh_tris.data(), B.bvh_tris.size() * 4);
    out_hash[11] = fnv(B.inst.data(), B.inst.size() * sizeof(B.inst[0])) ^ fnv(B.light_inst.data(), B.light_inst.size());
    return 1;
}

int ltrx_BakeResident(ltr_Scene *scene, float *gpu_ms_out)
{
    if (!scene->bake || !scene->bake->prepared || !scene->bake->gpu) { scene->error = "ltrx_BakeResident before ltrx_Prepare"; return 0; }
    int ok = guarded(scene, [&]() { gpu_stages(scene); });
    if (ok) {
        collect_counters(scene);
        if (gpu_ms_out) *gpu_ms_out = scene->stats.gpu_ms_span;     /* first-to-last CUDA event on the bake stream */
    }
    scene->stage.store(ok ? "baked" : nullptr);
    return ok;
}

int ltrx_Finish(ltr_Scene *scene)
{
    if (!scene->bake || !scene->bake->gpu) return 0;
    int ok = guarded(scene, [&]() { readback(scene); });
    scene->completion.store(1.f);
    scene->stage.store(nullptr, std::memory_order_release);
    return ok;
}

int ltrx_SetDebug(ltr_Scene *scene, int keep) { scene->keep_debug = keep; return 1; }

int ltrx_SetOutputRoot(ltr_Scene *scene, int root_only) { scene->output_root_only = root_only; return 1; }

/* FNV-1a-64 over the float bytes of every lightmap in output order (the fingerprint SURVEY 8c uses for the reference's
 * lightmap_rgb arrays), then over the probe colours; bench.py prints it and checks a sharded bake against the single-GPU one */
int ltrx_OutputHash(ltr_Scene *scene, uint64_t *fnv1a64)
{
    uint64_t h = 1469598103934665603ull;
    auto feed = [&h](const void *p, size_t n) { const unsigned char *c = (const unsigned char *)p; for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; } };
    for (const ltr_WorkOutput &wo : scene->outputs) feed(wo.lightmap_rgb, (size_t)wo.width * wo.height * 12);
    for (const ltr_SampleInfo &p : scene->probes) feed(p.out_color, 12);
    *fnv1a64 = h;
    return scene->outputs.empty() && scene->probes.empty() ? 0 : 1;
}

int ltrx_GetLumels(ltr_Scene *scene, u32 instance, ltrx_Lumels *out)
{
    Bake *B = scene->bake;
    if (!B || instance >= scene->instances.size() || B->lumel_off.size() != scene->instances.size() + 1 || B->d_loc.empty()) {
        memset(out, 0, sizeof(*out));
        return B && instance < scene->instances.size() && B->lumel_off.size() == scene->instances.size() + 1 &&
               B->lumel_off[instance] == B->lumel_off[instance + 1];
    }
    const uint64_t a = B->lumel_off[instance], b = B->lumel_off[instance + 1];
    out->count = (u32)(b - a);
    out->width = scene->instances[instance]->lm_width; out->height = scene->instances[instance]->lm_height;
    out->pos_xyz = B->d_pos.data() + a * 3; out->nrm_xyz = B->d_nrm.data() + a * 3; out->loc = B->d_loc.data() + a;
    out->radinfo_xyzw = B->d_rad.data() + a * 4; out->rgb = B->d_rgb.data() + a * 3;
    return 1;
}

int ltrx_GetLinks(ltr_Scene *scene, ltrx_Links *out)
{
    Bake *B = scene->bake;
    memset(out, 0, sizeof(*out));
    if (!B || B->d_lrow.empty()) return 0;
    out->rows = B->d_lrow.size() - 1; out->count = B->d_lother.size();
    out->row_offset = B->d_lrow.data(); out->other = B->d_lother.data(); out->factor = B->d_lfac.data();
    return 1;
}

/* host-only test hooks: the reference-order tree and the flat BVH builders (no GPU involved) */
int ltrx_test_reftree(const float *tris9, u32 ntris, void *nodes_out, u32 nodes_cap, int32_t *items_out, u32 items_cap, u32 *n_nodes, u32 *n_items)
{
    std::vector<Box3> boxes(ntris);
    for (u32 i = 0; i < ntris; ++i) {
        const float *t = tris9 + 9 * (size_t)i;
        V3 a = mk3(t[0], t[1], t[2]), b = mk3(t[3], t[4], t[5]), c = mk3(t[6], t[7], t[8]);
        boxes[i].lo = min3(a, min3(b, c)); boxes[i].hi = max3(a, max3(b, c));
    }
    RefTree T;
    T.build(boxes.data(), boxes.size(), getenv("LTR_REFTREE_THREADS") ? atoi(getenv("LTR_REFTREE_THREADS")) : 1);
    *n_nodes = (u32)T.nodes.size(); *n_items = (u32)T.items.size();
    if (T.nodes.size() > nodes_cap || T.items.size() > items_cap) return 0;
    memcpy(nodes_out, T.nodes.data(), T.nodes.size() * sizeof(RefNode));
    if (!T.items.empty()) memcpy(items_out, T.items.data(), T.items.size() * 4);
    return 1;
}

/* host model of the device any-hit walk on the 4-wide tree (gpu_internal.cuh: bvh4_anyhit_core), from a given start stack;
 * returns hit/miss and adds the nodes read to *visits */
static bool host_bvh4_anyhit(const SceneBvh &bvh, const std::vector<RayTri> &rt, V3 l1, V3 l2, std::vector<int32_t> &stack, uint64_t *visits)
{
    const V3 d = l2 - l1;
    const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
    bool hit = false;
    while (!stack.empty()) {
        const int32_t ni = stack.back(); stack.pop_back();
        const Bvh4Node &n = bvh.nodes4[ni];
        ++*visits;
        for (int c = 0; c < 4; ++c) {
            if (n.c[c] == BVH4_EMPTY) continue;
            const float x0 = (n.lox[c] - l1.x) * ix, x1 = (n.hix[c] - lb_slab_origin_hi(l1.x, d.x)) * ix, y0 = (n.loy[c] - l1.y) * iy, y1 = (n.hiy[c] - lb_slab_origin_hi(l1.y, d.y)) * iy;
            const float z0 = (n.loz[c] - l1.z) * iz, z1 = (n.hiz[c] - lb_slab_origin_hi(l1.z, d.z)) * iz;
            const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
            const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
            if (!(t0 <= t1 + 2e-6f)) continue;
            if (n.c[c] < 0) {
                const uint32_t code = ~n.c[c];
                for (uint32_t t = code >> 3; t < (code >> 3) + (code & 7u); ++t)
                    if (seg_tri_prepared(l1, d, rt[t]) < 1.0f) hit = true;       /* keep walking: the visit count is that of a miss */
            } else stack.push_back(n.c[c]);
        }
    }
    return hit;
}

/* host model again, with a triangle-test counter (per-segment cost profile for tools/entry_estimate.py) */
static void host_bvh4_cost(const SceneBvh &bvh, const std::vector<RayTri> &rt, V3 l1, V3 l2, std::vector<int32_t> &stack, uint32_t *nodes, uint32_t *tris)
{
    const V3 d = l2 - l1;
    const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
    bool hit = false;
    while (!stack.empty() && !hit) {
        const int32_t ni = stack.back(); stack.pop_back();
        const Bvh4Node &n = bvh.nodes4[ni];
        ++*nodes;
        for (int c = 0; c < 4 && !hit; ++c) {
            if (n.c[c] == BVH4_EMPTY) continue;
            const float x0 = (n.lox[c] - l1.x) * ix, x1 = (n.hix[c] - lb_slab_origin_hi(l1.x, d.x)) * ix, y0 = (n.loy[c] - l1.y) * iy, y1 = (n.hiy[c] - lb_slab_origin_hi(l1.y, d.y)) * iy;
            const float z0 = (n.loz[c] - l1.z) * iz, z1 = (n.hiz[c] - lb_slab_origin_hi(l1.z, d.z)) * iz;
            const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
            const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
            if (!(t0 <= t1 + 2e-6f)) continue;
            if (n.c[c] < 0) {
                const uint32_t code = ~n.c[c];
                for (uint32_t t = code >> 3; t < (code >> 3) + (code & 7u) && !hit; ++t) { ++*tris; hit = seg_tri_prepared(l1, d, rt[t]) < 1.0f; }
            } else stack.push_back(n.c[c]);
        }
    }
}

/* host-only: per-segment cost of the entry walk (4-wide node reads, triangle tests until the first hit) for bundles of
 * segments -- the input of the SIMT-efficiency estimate in tools/entry_estimate.py */
int ltrx_test_bvh_entry_cost(const float *tris9, u32 ntris, int leaf_max, const float *segs6, const u32 *bundle_off, u32 n_bundles,
                             uint32_t *nodes_out, uint32_t *tris_out)
{
    SceneBvh bvh;
    build_scene_bvh(tris9, ntris, bvh, leaf_max, 0);
    if (bvh.nodes4.empty()) return 0;
    std::vector<RayTri> rt(ntris);
    for (u32 t = 0; t < ntris; ++t) {
        const float *v = tris9 + 9 * (size_t)bvh.order[t];
        prepare_raytri(mk3(v[0], v[1], v[2]), mk3(v[3], v[4], v[5]), mk3(v[6], v[7], v[8]), rt[t]);
    }
    std::vector<int32_t> stack;
    for (u32 b = 0; b < n_bundles; ++b) {
        float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
        for (u32 s = bundle_off[b]; s < bundle_off[b + 1]; ++s)
            for (int e = 0; e < 2; ++e) {
                const float *p = segs6 + 6 * (size_t)s + 3 * e;
                lx = fminf(lx, p[0]); ly = fminf(ly, p[1]); lz = fminf(lz, p[2]); hx = fmaxf(hx, p[0]); hy = fmaxf(hy, p[1]); hz = fmaxf(hz, p[2]);
            }
        BvhEntrySet E;
        E.n = 0;
        if (lx <= hx) { bvh_entry_pad(lx, ly, lz, hx, hy, hz); bvh4_entry_search(bvh.nodes4.data(), lx, ly, lz, hx, hy, hz, E); }
        for (u32 s = bundle_off[b]; s < bundle_off[b + 1]; ++s) {
            const float *p = segs6 + 6 * (size_t)s;
            const V3 A = mk3(p[0], p[1], p[2]), B = mk3(p[3], p[4], p[5]), d = B - A;
            const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
            stack.clear();
            for (int i = 0; i < E.n; ++i) {
                const float x0 = (E.lox[i] - A.x) * ix, x1 = (E.hix[i] - lb_slab_origin_hi(A.x, d.x)) * ix, y0 = (E.loy[i] - A.y) * iy, y1 = (E.hiy[i] - lb_slab_origin_hi(A.y, d.y)) * iy;
                const float z0 = (E.loz[i] - A.z) * iz, z1 = (E.hiz[i] - lb_slab_origin_hi(A.z, d.z)) * iz;
                const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
                if (t0 <= t1 + 2e-6f) stack.push_back(E.node[i]);
            }
            nodes_out[s] = 0; tris_out[s] = 0;
            host_bvh4_cost(bvh, rt, A, B, stack, nodes_out + s, tris_out + s);
        }
    }
    return 1;
}

/* host-only check of the entry sets (bvh_entry.h) on bundles of segments: per bundle the box of its segments is padded,
 * the entry set searched, and every segment walked twice -- from the root and from the entry set.  Returns 0 if any
 * triangle whose box overlaps a bundle box is unreachable from that bundle's entry set; *mismatches counts segments
 * whose two walks disagree (must be 0).  visits are 4-wide node reads summed over all segments. */
int ltrx_test_bvh_entry(const float *tris9, u32 ntris, int leaf_max, const float *segs6, const u32 *bundle_off, u32 n_bundles,
                        u32 *entries_out, uint64_t *visits_root, uint64_t *visits_entry, uint64_t *entry_tests, u32 *mismatches)
{
    SceneBvh bvh;
    build_scene_bvh(tris9, ntris, bvh, leaf_max, 0);
    if (bvh.nodes4.empty()) return 0;
    int max_entries = BVH_ENTRY_MAX;
    if (const char *e = getenv("LTR_TEST_ENTRY_MAX")) { max_entries = atoi(e); if (max_entries < 1 || max_entries > BVH_ENTRY_MAX) max_entries = BVH_ENTRY_MAX; }
    std::vector<RayTri> rt(ntris);
    for (u32 t = 0; t < ntris; ++t) {
        const float *v = tris9 + 9 * (size_t)bvh.order[t];
        prepare_raytri(mk3(v[0], v[1], v[2]), mk3(v[3], v[4], v[5]), mk3(v[6], v[7], v[8]), rt[t]);
    }
    *visits_root = *visits_entry = *entry_tests = 0; *mismatches = 0;
    std::vector<int32_t> stack;
    std::vector<char> reach(ntris);
    int ok = 1;
    for (u32 b = 0; b < n_bundles; ++b) {
        float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
        for (u32 s = bundle_off[b]; s < bundle_off[b + 1]; ++s)
            for (int e = 0; e < 2; ++e) {
                const float *p = segs6 + 6 * (size_t)s + 3 * e;
                lx = fminf(lx, p[0]); ly = fminf(ly, p[1]); lz = fminf(lz, p[2]); hx = fmaxf(hx, p[0]); hy = fmaxf(hy, p[1]); hz = fmaxf(hz, p[2]);
            }
        BvhEntrySet E;
        E.n = 0;
        if (lx <= hx) {
            bvh_entry_pad(lx, ly, lz, hx, hy, hz);
            bvh4_entry_search(bvh.nodes4.data(), lx, ly, lz, hx, hy, hz, E, max_entries);
        }
        if (entries_out) entries_out[b] = (u32)E.n;
        /* reachability: every triangle whose own box overlaps the padded bundle box */
        std::fill(reach.begin(), reach.end(), 0);
        for (int i = 0; i < E.n; ++i) stack.push_back(E.node[i]);
        while (!stack.empty()) {
            const Bvh4Node &n = bvh.nodes4[stack.back()]; stack.pop_back();
            for (int c = 0; c < 4; ++c) {
                if (n.c[c] == BVH4_EMPTY) continue;
                if (n.c[c] >= 0) { stack.push_back(n.c[c]); continue; }
                const uint32_t code = ~n.c[c];
                for (uint32_t t = code >> 3; t < (code >> 3) + (code & 7u); ++t) reach[t] = 1;
            }
        }
        if (lx <= hx)
            for (u32 t = 0; t < ntris; ++t) {
                const float *v = tris9 + 9 * (size_t)bvh.order[t];
                const float tlx = fminf(v[0], fminf(v[3], v[6])), thx = fmaxf(v[0], fmaxf(v[3], v[6]));
                const float tly = fminf(v[1], fminf(v[4], v[7])), thy = fmaxf(v[1], fmaxf(v[4], v[7]));
                const float tlz = fminf(v[2], fminf(v[5], v[8])), thz = fmaxf(v[2], fmaxf(v[5], v[8]));
                if (tlx <= hx && thx >= lx && tly <= hy && thy >= ly && tlz <= hz && thz >= lz && !reach[t]) ok = 0;
            }
        /* the same on the binary tree (closest-hit walks: ambient occlusion) */
        if (lx <= hx) {
            BvhEntrySet E2;
            bvh2_entry_search(bvh.nodes.data(), lx, ly, lz, hx, hy, hz, E2, max_entries);
            std::fill(reach.begin(), reach.end(), 0);
            for (int i = 0; i < E2.n; ++i) stack.push_back(E2.node[i]);
            while (!stack.empty()) {
                const BvhNode &n = bvh.nodes[stack.back()]; stack.pop_back();
                const int32_t cs[2] = { n.c0, n.c1 };
                for (int c = 0; c < 2; ++c) {
                    if (cs[c] >= 0) { stack.push_back(cs[c]); continue; }
                    const uint32_t code = ~cs[c];
                    for (uint32_t t = code >> 3; t < (code >> 3) + (code & 7u); ++t) reach[t] = 1;
                }
            }
            for (u32 t = 0; t < ntris; ++t) {
                const float *v = tris9 + 9 * (size_t)bvh.order[t];
                const float tlx = fminf(v[0], fminf(v[3], v[6])), thx = fmaxf(v[0], fmaxf(v[3], v[6]));
                const float tly = fminf(v[1], fminf(v[4], v[7])), thy = fmaxf(v[1], fmaxf(v[4], v[7]));
                const float tlz = fminf(v[2], fminf(v[5], v[8])), thz = fmaxf(v[2], fmaxf(v[5], v[8]));
                if (tlx <= hx && thx >= lx && tly <= hy && thy >= ly && tlz <= hz && thz >= lz && !reach[t]) ok = 0;
            }
        }
        for (u32 s = bundle_off[b]; s < bundle_off[b + 1]; ++s) {
            const float *p = segs6 + 6 * (size_t)s;
            const V3 A = mk3(p[0], p[1], p[2]), B = mk3(p[3], p[4], p[5]);
            stack.assign(1, 0);
            const bool h_root = host_bvh4_anyhit(bvh, rt, A, B, stack, visits_root);
            /* the entry tests of the device walk (gpu_internal.cuh: bvh4_anyhit_entries) */
            const V3 d = B - A;
            const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
            stack.clear();
            for (int i = 0; i < E.n; ++i) {
                const float x0 = (E.lox[i] - A.x) * ix, x1 = (E.hix[i] - lb_slab_origin_hi(A.x, d.x)) * ix, y0 = (E.loy[i] - A.y) * iy, y1 = (E.hiy[i] - lb_slab_origin_hi(A.y, d.y)) * iy;
                const float z0 = (E.loz[i] - A.z) * iz, z1 = (E.hiz[i] - lb_slab_origin_hi(A.z, d.z)) * iz;
                const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
                if (t0 <= t1 + 2e-6f) stack.push_back(E.node[i]);
            }
            *entry_tests += (uint64_t)E.n;
            const bool h_entry = host_bvh4_anyhit(bvh, rt, A, B, stack, visits_entry);
            if (h_root != h_entry) ++*mismatches;
        }
    }
    return ok;
}

/* host-only model of the version-2 entry sets (bvh_entry.h: shaft-culled search, leaf entries, <= 16 entries) on bundles of
 * segments whose FIRST end points form box R and whose SECOND end points form box C, as in rad_visibility_kernel.  Every
 * segment is walked from the root and from its bundle's entry set, counting the triangles each walk TESTS as well: the two
 * walks must not only agree on hit / miss (*mismatches) but test the same number of triangles when the segment is free
 * (*test_diffs) -- a dropped box that a ray would have entered shows up there even if it held no blocker.
 * stats[0..3] = 4-wide node reads from the root, from the entry sets, entry boxes tested, triangle tests of the entry walk. */
int ltrx_test_bvh_entry2(const float *tris9, u32 ntris, int leaf_max, const float *segs6, const u32 *bundle_off, u32 n_bundles,
                         int max_entries, int use_shaft, int batch, u32 *entries_out, uint64_t *stats4, u32 *mismatches, u32 *test_diffs)
{
    SceneBvh bvh;
    build_scene_bvh(tris9, ntris, bvh, leaf_max, 0);
    if (bvh.nodes4.empty()) return 0;
    if (max_entries < 1 || max_entries > BVH_ENTRY2_MAX) max_entries = BVH_ENTRY2_MAX;
    std::vector<RayTri> rt(ntris);
    for (u32 t = 0; t < ntris; ++t) {
        const float *v = tris9 + 9 * (size_t)bvh.order[t];
        prepare_raytri(mk3(v[0], v[1], v[2]), mk3(v[3], v[4], v[5]), mk3(v[6], v[7], v[8]), rt[t]);
    }
    stats4[0] = stats4[1] = stats4[2] = stats4[3] = 0; *mismatches = 0; *test_diffs = 0;
    std::vector<int32_t> stack;
    /* walk that tests every triangle of every accepted leaf (no early out), from a stack of nodes + a list of leaf codes */
    auto walk = [&](V3 l1, V3 l2, std::vector<int32_t> &st, const std::vector<int32_t> &leaves, uint64_t *visits, uint64_t *tests) {
        const V3 d = l2 - l1;
        const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
        bool hit = false;
        auto leaf = [&](int32_t c) { const uint32_t code = ~c; for (uint32_t t = code >> 3; t < (code >> 3) + (code & 7u); ++t) { ++*tests; if (seg_tri_prepared(l1, d, rt[t]) < 1.0f) hit = true; } };
        for (int32_t c : leaves) leaf(c);
        while (!st.empty()) {
            const Bvh4Node &n = bvh.nodes4[st.back()]; st.pop_back();
            ++*visits;
            for (int c = 0; c < 4; ++c) {
                if (n.c[c] == BVH4_EMPTY) continue;
                const float x0 = (n.lox[c] - l1.x) * ix, x1 = (n.hix[c] - lb_slab_origin_hi(l1.x, d.x)) * ix, y0 = (n.loy[c] - l1.y) * iy, y1 = (n.hiy[c] - lb_slab_origin_hi(l1.y, d.y)) * iy;
                const float z0 = (n.loz[c] - l1.z) * iz, z1 = (n.hiz[c] - lb_slab_origin_hi(l1.z, d.z)) * iz;
                const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
                if (!(t0 <= t1 + 2e-6f)) continue;
                if (n.c[c] < 0) leaf(n.c[c]); else st.push_back(n.c[c]);
            }
        }
        return hit;
    };
    std::vector<int32_t> leaves, none;
    for (u32 b = 0; b < n_bundles; ++b) {
        float q[6] = { INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY };
        float RC[12] = { INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY, INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY };
        for (u32 s = bundle_off[b]; s < bundle_off[b + 1]; ++s)
            for (int e = 0; e < 2; ++e) {
                const float *p = segs6 + 6 * (size_t)s + 3 * e;
                for (int k = 0; k < 3; ++k) {
                    q[k] = fminf(q[k], p[k]); q[3 + k] = fmaxf(q[3 + k], p[k]);
                    RC[6 * e + k] = fminf(RC[6 * e + k], p[k]); RC[6 * e + 3 + k] = fmaxf(RC[6 * e + 3 + k], p[k]);
                }
            }
        BvhEntrySet2 E;
        E.n = 0;
        if (q[0] <= q[3]) {
            const float maxabs = fmaxf(fmaxf(fmaxf(fabsf(q[0]), fabsf(q[3])), fmaxf(fabsf(q[1]), fabsf(q[4]))), fmaxf(fabsf(q[2]), fabsf(q[5])));
            bvh_entry_pad(q[0], q[1], q[2], q[3], q[4], q[5]);
            BvhShaft S;
            bvh_shaft_build(RC, RC + 6, maxabs, S);
            bvh4_entry_search2(bvh.nodes4.data(), q[0], q[1], q[2], q[3], q[4], q[5], use_shaft ? &S : nullptr, E, max_entries);
        }
        if (entries_out) entries_out[b] = (u32)E.n;
        /* batch > 0: the PACKET form of rad_visibility_kernel -- per `batch` consecutive segments the leaves in the batch's own
         * shaft are listed by one walk from the bundle's entry set (stats4[1] counts its node reads), and a ray only tests the
         * listed leaf boxes (stats4[2]) */
        for (u32 s0 = bundle_off[b]; batch > 0 && s0 < bundle_off[b + 1]; s0 += (u32)batch) {
            const u32 s1 = std::min<u32>(s0 + (u32)batch, bundle_off[b + 1]);
            float bq[6] = { INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY };
            float brc[12] = { INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY, INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY };
            for (u32 s = s0; s < s1; ++s)
                for (int e = 0; e < 2; ++e) {
                    const float *p = segs6 + 6 * (size_t)s + 3 * e;
                    for (int k = 0; k < 3; ++k) {
                        bq[k] = fminf(bq[k], p[k]); bq[3 + k] = fmaxf(bq[3 + k], p[k]);
                        brc[6 * e + k] = fminf(brc[6 * e + k], p[k]); brc[6 * e + 3 + k] = fmaxf(brc[6 * e + 3 + k], p[k]);
                    }
                }
            const float bmaxabs = fmaxf(fmaxf(fmaxf(fabsf(bq[0]), fabsf(bq[3])), fmaxf(fabsf(bq[1]), fabsf(bq[4]))), fmaxf(fabsf(bq[2]), fabsf(bq[5])));
            bvh_entry_pad(bq[0], bq[1], bq[2], bq[3], bq[4], bq[5]);
            BvhShaft BS;
            bvh_shaft_build(brc, brc + 6, bmaxabs, BS);
            struct LeafBox { float b[6]; int32_t code; };
            std::vector<LeafBox> list;
            auto consider = [&](int32_t code, float lx, float ly, float lz, float hx, float hy, float hz) {
                if (!(lx <= bq[3] && hx >= bq[0] && ly <= bq[4] && hy >= bq[1] && lz <= bq[5] && hz >= bq[2])) return;
                if (use_shaft && bvh_shaft_outside(BS, lx, ly, lz, hx, hy, hz)) return;
                if (code < 0) list.push_back(LeafBox{ { lx, ly, lz, hx, hy, hz }, code }); else stack.push_back(code);
            };
            stack.clear();
            for (int i = 0; i < E.n; ++i) { int32_t code; memcpy(&code, &E.lo[i][3], 4); consider(code, E.lo[i][0], E.lo[i][1], E.lo[i][2], E.hi[i][0], E.hi[i][1], E.hi[i][2]); }
            while (!stack.empty()) {
                const Bvh4Node &n = bvh.nodes4[stack.back()]; stack.pop_back();
                ++stats4[1];
                for (int c = 0; c < 4; ++c) if (n.c[c] != BVH4_EMPTY) consider(n.c[c], n.lox[c], n.loy[c], n.loz[c], n.hix[c], n.hiy[c], n.hiz[c]);
            }
            for (u32 s = s0; s < s1; ++s) {
                const float *p = segs6 + 6 * (size_t)s;
                const V3 A = mk3(p[0], p[1], p[2]), B = mk3(p[3], p[4], p[5]), d = B - A;
                uint64_t t_root = 0, t_entry = 0;
                stack.assign(1, 0);
                const bool h_root = walk(A, B, stack, none, &stats4[0], &t_root);
                const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
                bool h_entry = false;
                for (const LeafBox &L : list) {
                    const float x0 = (L.b[0] - A.x) * ix, x1 = (L.b[3] - lb_slab_origin_hi(A.x, d.x)) * ix, y0 = (L.b[1] - A.y) * iy, y1 = (L.b[4] - lb_slab_origin_hi(A.y, d.y)) * iy;
                    const float z0 = (L.b[2] - A.z) * iz, z1 = (L.b[5] - lb_slab_origin_hi(A.z, d.z)) * iz;
                    const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                    const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
                    if (!(t0 <= t1 + 2e-6f)) continue;
                    const uint32_t code = ~L.code;
                    for (uint32_t t = code >> 3; t < (code >> 3) + (code & 7u); ++t) { ++t_entry; if (seg_tri_prepared(A, d, rt[t]) < 1.0f) h_entry = true; }
                }
                stats4[2] += list.size();
                stats4[3] += t_entry;
                if (h_root != h_entry) ++*mismatches;
                if (t_root != t_entry) ++*test_diffs;
            }
        }
        for (u32 s = bundle_off[b]; batch <= 0 && s < bundle_off[b + 1]; ++s) {
            const float *p = segs6 + 6 * (size_t)s;
            const V3 A = mk3(p[0], p[1], p[2]), B = mk3(p[3], p[4], p[5]);
            uint64_t t_root = 0, t_entry = 0;
            stack.assign(1, 0);
            const bool h_root = walk(A, B, stack, none, &stats4[0], &t_root);
            const V3 d = B - A;
            const float ix = lb_slab_inv(d.x), iy = lb_slab_inv(d.y), iz = lb_slab_inv(d.z);
            stack.clear(); leaves.clear();
            for (int i = 0; i < E.n; ++i) {
                const float x0 = (E.lo[i][0] - A.x) * ix, x1 = (E.hi[i][0] - lb_slab_origin_hi(A.x, d.x)) * ix, y0 = (E.lo[i][1] - A.y) * iy, y1 = (E.hi[i][1] - lb_slab_origin_hi(A.y, d.y)) * iy;
                const float z0 = (E.lo[i][2] - A.z) * iz, z1 = (E.hi[i][2] - lb_slab_origin_hi(A.z, d.z)) * iz;
                const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
                if (!(t0 <= t1 + 2e-6f)) continue;
                int32_t code; memcpy(&code, &E.lo[i][3], 4);
                if (code >= 0) stack.push_back(code); else leaves.push_back(code);
            }
            stats4[2] += (uint64_t)E.n;
            const bool h_entry = walk(A, B, stack, leaves, &stats4[1], &t_entry);
            stats4[3] += t_entry;
            if (h_root != h_entry) ++*mismatches;
            if (t_root != t_entry) ++*test_diffs;
        }
    }
    return 1;
}

int ltrx_test_bvh(const float *tris9, u32 ntris, int leaf_max, u32 *n_nodes, u32 *depth, u32 *order_out, float *bounds6)
{
    SceneBvh bvh;
    build_scene_bvh(tris9, ntris, bvh, leaf_max, 0);
    *n_nodes = (u32)bvh.nodes.size(); *depth = (u32)bvh.depth;
    if (order_out && ntris) memcpy(order_out, bvh.order.data(), (size_t)ntris * 4);
    if (bounds6) { bounds6[0] = bvh.bounds.lo.x; bounds6[1] = bvh.bounds.lo.y; bounds6[2] = bvh.bounds.lo.z; bounds6[3] = bvh.bounds.hi.x; bounds6[4] = bvh.bounds.hi.y; bounds6[5] = bvh.bounds.hi.z; }
    /* structural self-check: every triangle appears in exactly one leaf and inside its leaf box */
    std::vector<int> seen(ntris, 0);
    for (const BvhNode &n : bvh.nodes) {
        const int32_t cs[2] = { n.c0, n.c1 };
        const float lo[2][3] = { { n.lo0x, n.lo0y, n.lo0z }, { n.lo1x, n.lo1y, n.lo1z } }, hi[2][3] = { { n.hi0x, n.hi0y, n.hi0z }, { n.hi1x, n.hi1y, n.hi1z } };
        for (int k = 0; k < 2; ++k) {
            if (cs[k] >= 0) continue;
            uint32_t code = ~cs[k], first = code >> 3, cnt = code & 7u;
            for (uint32_t t = first; t < first + cnt; ++t) {
                if (t >= ntris) return 0;
                seen[t]++;
                const float *v = tris9 + 9 * (size_t)bvh.order[t];
                for (int c = 0; c < 9; ++c) if (v[c] < lo[k][c % 3] || v[c] > hi[k][c % 3]) return 0;
            }
        }
    }
    for (u32 t = 0; t < ntris; ++t) if (seen[t] != 1) return 0;
    /* the 4-wide collapse: reachable from its root, every triangle in exactly one leaf slot and inside that slot's box,
     * every inner slot's box containing the boxes of the node it points to */
    if (bvh.nodes4.empty()) return 0;
    std::vector<int> seen4(ntris, 0);
    std::vector<char> visited(bvh.nodes4.size(), 0);
    std::vector<int32_t> st{ 0 };
    while (!st.empty()) {
        const int32_t i = st.back(); st.pop_back();
        if (i < 0 || (size_t)i >= bvh.nodes4.size() || visited[i]) return 0;
        visited[i] = 1;
        const Bvh4Node &n = bvh.nodes4[i];
        for (int j = 0; j < 4; ++j) {
            if (n.c[j] == BVH4_EMPTY) continue;
            const float lo[3] = { n.lox[j], n.loy[j], n.loz[j] }, hi[3] = { n.hix[j], n.hiy[j], n.hiz[j] };
            if (n.c[j] >= 0) {
                if ((size_t)n.c[j] >= bvh.nodes4.size()) return 0;
                const Bvh4Node &c = bvh.nodes4[n.c[j]];
                for (int q = 0; q < 4; ++q) {
                    if (c.c[q] == BVH4_EMPTY) continue;
                    if (c.lox[q] < lo[0] || c.loy[q] < lo[1] || c.loz[q] < lo[2] || c.hix[q] > hi[0] || c.hiy[q] > hi[1] || c.hiz[q] > hi[2]) return 0;
                }
                st.push_back(n.c[j]);
            } else {
                uint32_t code = ~n.c[j], first = code >> 3, cnt = code & 7u;
                for (uint32_t t = first; t < first + cnt; ++t) {
                    if (t >= ntris) return 0;
                    seen4[t]++;
                    const float *v = tris9 + 9 * (size_t)bvh.order[t];
                    for (int c = 0; c < 9; ++c) if (v[c] < lo[c % 3] || v[c] > hi[c % 3]) return 0;
                }
            }
        }
    }
    for (u32 t = 0; t < ntris; ++t) if (seen4[t] != 1) return 0;
    for (char v : visited) if (!v) return 0;
    return 1;
}

int ltrx_test_rand_fill(float *out, uint64_t n) { return rand_fill(out, n) ? 1 : 0; }

int ltrx_SetShadowMode(ltr_Scene *scene, int mode)
{
    if (mode != LTRX_SHADOW_MARCH && mode != LTRX_SHADOW_SAMPLED) return 0;
    scene->shadow_mode = mode;
    return 1;
}

int ltrx_GetShadowMasks(ltr_Scene *scene, u32 light, const uint64_t **out, uint64_t *count)
{
    Bake *B = scene->bake;
    if (!B || B->d_smask.empty() || light >= scene->lights.size()) return 0;
    const uint64_t nl = scene->stats.n_lumels_local;
    *out = B->d_smask.data() + (size_t)light * nl; *count = nl;
    return 1;
}

/* host evaluation of the shadow segment of (light, sample) for a lumel: the SAME inline function the kernel
 * uses (geom.h), compiled for the host with contraction off -- bit-identical end points */
int ltrx_ShadowSampleSegment(ltr_Scene *scene, u32 light, u32 sample, const float pos[3], const float nrm[3], float from_out[3], float to_out[3])
{
    Bake *B = scene->bake;
    if (!B || light >= B->lights.size()) return 0;
    const ltrgpu_Light &L = B->lights[light];
    if (sample >= L.n_samples) return 0;
    const float *sm = &B->light_samples[(size_t)(L.sample_off + sample) * 4];
    V3 from, to;
    shadow_sample_segment(L.type, L.pos, L.range, mk3(sm[0], sm[1], sm[2]), mk3(pos[0], pos[1], pos[2]), mk3(nrm[0], nrm[1], nrm[2]), from, to);
    from_out[0] = from.x; from_out[1] = from.y; from_out[2] = from.z;
    to_out[0] = to.x; to_out[1] = to.y; to_out[2] = to.z;
    return 1;
}

int ltrx_GetShadowFactors(ltr_Scene *scene, u32 light, const float **out, uint64_t *count)
{
    Bake *B = scene->bake;
    if (!B || B->d_fvis.empty() || light >= scene->lights.size()) return 0;
    const uint64_t nl = scene->stats.n_lumels_local;
    *out = B->d_fvis.data() + (size_t)light * nl; *count = nl;
    return 1;
}

} /* extern "C" */
