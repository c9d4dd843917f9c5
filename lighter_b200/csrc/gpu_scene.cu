/* gpu_scene.cu -- context lifetime, scene upload, per-triangle preparation, counters. */
#include "gpu_internal.cuh"

#include <stdlib.h>
#include <time.h>

#include <map>
#include <mutex>
#include <unordered_map>

/* ---- caching device allocator (see gpu_internal.cuh) ------------------------------------------ */
namespace {
struct Pool {
    std::mutex mu;
    std::unordered_map<void *, std::pair<int, size_t>> live;       /* ptr -> (device, bytes) */
    std::multimap<std::pair<int, size_t>, void *> idle;            /* (device, bytes) -> ptr */
};
Pool &pool() { static Pool p; return p; }
size_t size_class(size_t bytes)
{
    if (bytes < 4096) return 4096;
    /* round up to 1/8 steps of the enclosing power of two: <= 12.5 % slack, few distinct classes */
    size_t p2 = 4096;
    while (p2 < bytes) p2 <<= 1;
    const size_t step = p2 >> 4;
    return (bytes + step - 1) / step * step;
}
} // namespace

cudaError_t lb_malloc(void **p, size_t bytes)
{
    Pool &P = pool();
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t cls = size_class(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> g(P.mu);
        auto it = P.idle.find({ dev, cls });
        if (it != P.idle.end()) {
            *p = it->second;
            P.idle.erase(it);
            P.live[*p] = { dev, cls };
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, cls);
    if (e != cudaSuccess) {                       /* out of memory: give cached blocks back and retry once */
        cudaGetLastError();
        lb_trim();
        e = cudaMalloc(p, cls);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> g(P.mu);
        P.live[*p] = { dev, cls };
    }
    return e;
}

void lb_free(void *p)
{
    if (!p) return;
    Pool &P = pool();
    std::lock_guard<std::mutex> g(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) { cudaFree(p); return; }
    P.idle.insert({ it->second, p });
    P.live.erase(it);
}

void lb_trim(void)
{
    Pool &P = pool();
    std::lock_guard<std::mutex> g(P.mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &kv : P.idle) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
    P.idle.clear();
    cudaSetDevice(cur);
}

/* ---- caching page-locked host allocator --------------------------------------------------------- */
namespace {
struct HostPool {
    std::mutex mu;
    std::unordered_map<void *, std::pair<size_t, bool>> live;      /* ptr -> (bytes, pinned) */
    std::multimap<size_t, void *> idle;                            /* pinned blocks only */
};
HostPool &host_pool() { static HostPool p; return p; }
} // namespace

extern "C" void *ltrgpu_host_alloc(size_t bytes)
{
    HostPool &P = host_pool();
    const size_t cls = size_class(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> g(P.mu);
        auto it = P.idle.find(cls);
        if (it != P.idle.end()) {
            void *p = it->second;
            P.idle.erase(it);
            P.live[p] = { cls, true };
            return p;
        }
    }
    void *p = nullptr;
    bool pinned = cudaHostAlloc(&p, cls, cudaHostAllocPortable) == cudaSuccess;
    if (!pinned) { cudaGetLastError(); p = malloc(cls); }
    if (!p) return nullptr;
    std::lock_guard<std::mutex> g(P.mu);
    P.live[p] = { cls, pinned };
    return p;
}

extern "C" void ltrgpu_host_free(void *p)
{
    if (!p) return;
    HostPool &P = host_pool();
    std::lock_guard<std::mutex> g(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) { free(p); return; }
    if (it->second.second) P.idle.insert({ it->second.first, p }); else free(p);
    P.live.erase(it);
}

/* Host -> device copy of a pageable array through two pinned staging buffers: the CPU copies chunk k+1
 * into one buffer while the DMA engine drains the other.  (cudaMemcpyAsync from pageable memory stages
 * internally too, but synchronously and in small pieces: measured ~1.7 GB/s on the scene upload.) */
int lb_upload_staged(ltrgpu_Ctx *ctx, void *dst, const void *src, size_t bytes)
{
    const size_t CHUNK = (size_t)8 << 20;
    bool pinned = false;                                          /* page-locked source (ltrgpu_host_alloc, or the caller's own): one plain DMA */
    if (bytes >= ((size_t)1 << 20)) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess) pinned = at.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    if (pinned || bytes < ((size_t)1 << 20)) {
        CU_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    }
    if (!ctx->stage_buf[0]) {
        for (int b = 0; b < 2; ++b) {
            ctx->stage_buf[b] = ltrgpu_host_alloc(CHUNK);
            if (!ctx->stage_buf[b]) { snprintf(ctx->err, sizeof(ctx->err), "out of host memory for the upload staging buffers"); return 1; }
            CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[b], cudaEventDisableTiming));
        }
    }
    int b = 0;
    for (size_t off = 0; off < bytes; off += CHUNK, b ^= 1) {
        const size_t n = bytes - off < CHUNK ? bytes - off : CHUNK;
        CU_TRY(ctx, cudaEventSynchronize(ctx->stage_ev[b]));          /* previous DMA out of this buffer finished */
        memcpy(ctx->stage_buf[b], (const char *)src + off, n);
        CU_TRY(ctx, cudaMemcpyAsync((char *)dst + off, ctx->stage_buf[b], n, cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(ctx, cudaEventRecord(ctx->stage_ev[b], ctx->stream));
    }
    return 0;
}

#if LB_VIS_Q8
/* Bvh4Node -> Bvh4QNode (bvh.h), one thread per node.  Per axis: step = the power of two >= extent / 250 (at least 2^-12),
 * corner = lowest child plane minus one step, rounded down; a lower plane is floor((lo - corner) / step) - 1, an upper plane
 * ceil((hi - corner) / step) + 1 (computed in double: exact for these magnitudes), so corner + q * step lies at least one
 * full step outside the float plane.  Unused slots get an inverted box. */
__global__ void quantise_bvh4_kernel(const Bvh4Node *__restrict__ in, uint32_t n, Bvh4QNode *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Bvh4Node N = in[i];
    Bvh4QNode Q;
    const float *lo[3] = { N.lox, N.loy, N.loz }, *hi[3] = { N.hix, N.hiy, N.hiz };
    float corner[3];
    uint32_t ql[3] = { 0, 0, 0 }, qh[3] = { 0, 0, 0 }, exps = 0;
    for (int a = 0; a < 3; ++a) {
        float mn = INFINITY, mx = -INFINITY;
        for (int c = 0; c < 4; ++c) if (N.c[c] != BVH4_EMPTY) { mn = fminf(mn, lo[a][c]); mx = fmaxf(mx, hi[a][c]); }
        if (!(mn <= mx)) { mn = mx = 0.f; }
        int e;
        frexpf(fmaxf((mx - mn) / 250.0f, 1.0f / 4096.0f), &e);          /* x = m * 2^e, m in [0.5, 1): 2^e >= x */
        const float step = ldexpf(1.0f, e);
        const float o = __fsub_rd(mn, step);
        corner[a] = o;
        exps |= (uint32_t)(e + 127) << (8 * a);
        for (int c = 0; c < 4; ++c) {
            uint32_t l = 255u, h = 0u;
            if (N.c[c] != BVH4_EMPTY) {
                const double fl = floor(((double)lo[a][c] - (double)o) / (double)step) - 1.0, ch = ceil(((double)hi[a][c] - (double)o) / (double)step) + 1.0;
                l = (uint32_t)fmin(fmax(fl, 0.0), 255.0); h = (uint32_t)fmin(fmax(ch, 0.0), 255.0);
            }
            ql[a] |= l << (8 * c); qh[a] |= h << (8 * c);
        }
    }
    Q.ox = corner[0]; Q.oy = corner[1]; Q.oz = corner[2]; Q.exps = exps;
    Q.qlox = ql[0]; Q.qloy = ql[1]; Q.qloz = ql[2]; Q.qhix = qh[0]; Q.qhiy = qh[1]; Q.qhiz = qh[2];
    for (int c = 0; c < 4; ++c) Q.c[c] = N.c[c];
    Q.pad[0] = Q.pad[1] = 0;
    out[i] = Q;
}
#endif

/* one thread per BVH-order triangle: expand the 9 floats into the two prepared records
 * (geom.h) with the reference's exact expressions. */
__global__ void prepare_tris_kernel(const float *__restrict__ tris9, const uint32_t *__restrict__ order /* slot -> triangle, or NULL */, uint32_t n,
                                    PreparedTri *__restrict__ pt, RayTri *__restrict__ rt)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *t = tris9 + 9ull * (order ? order[i] : i);
    V3 a = mk3(t[0], t[1], t[2]), b = mk3(t[3], t[4], t[5]), c = mk3(t[6], t[7], t[8]);
    PreparedTri P;
    prepare_tri(a, b, c, P);
    pt[i] = P;
    if (rt) {
        RayTri R;
        prepare_raytri(a, b, c, R);
        rt[i] = R;
    }
}

__global__ void tri_boxes_kernel(const float *__restrict__ tris9, uint32_t n, float4 *__restrict__ boxes)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *t = tris9 + 9ull * i;
    V3 a = mk3(t[0], t[1], t[2]), b = mk3(t[3], t[4], t[5]), c = mk3(t[6], t[7], t[8]);
    V3 lo = min3(a, min3(b, c)), hi = max3(a, max3(b, c));
    boxes[2ull * i] = make_float4(lo.x, lo.y, lo.z, 0.f);
    boxes[2ull * i + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
}

extern "C" int ltrgpu_create(ltrgpu_Ctx **out, int device)
{
    *out = nullptr;
    const bool trace = getenv("LTR_TRACE") != nullptr;
    struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
    auto lap = [&](const char *what) {
        if (!trace) return;
        struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "[ltr gpu] create %-24s %8.2f ms\n", what, (t.tv_sec - t0.tv_sec) * 1e3 + (t.tv_nsec - t0.tv_nsec) * 1e-6);
        t0 = t;
    };
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        fprintf(stderr, "lighter_b200: no CUDA device available (%s); this library has no CPU fallback\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return 1;
    }
    ltrgpu_Ctx *ctx = new ltrgpu_Ctx;
    memset(&ctx->host_counters, 0, sizeof(ctx->host_counters));
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    ctx->device = device;
    *out = ctx;
    CU_TRY(ctx, cudaSetDevice(device));
    lap("device count + set");
    /* the SM count only (cudaGetDeviceProperties fills ~100 fields and costs tens of milliseconds per call) */
    CU_TRY(ctx, cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device));
    lap("attributes");
    CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU_TRY(ctx, cudaEventCreate(&ctx->ev0));
    CU_TRY(ctx, cudaEventCreate(&ctx->ev1));
    CU_TRY(ctx, cudaEventCreate(&ctx->ev_span0));
    CU_TRY(ctx, cudaEventCreate(&ctx->ev_span1));
    CU_TRY(ctx, cudaEventCreate(&ctx->ev_k0));
    CU_TRY(ctx, cudaEventCreate(&ctx->ev_k1));
    lap("stream + events");
    if (dev_alloc(ctx, &ctx->d_counters, CNT_COUNT)) return 1;
    CU_TRY(ctx, cudaMemsetAsync(ctx->d_counters, 0, CNT_COUNT * sizeof(unsigned long long), ctx->stream));
    lap("counters");
    return 0;
}

static void free_bake_state(ltrgpu_Ctx *ctx)
{
    dev_free(&ctx->d_texkey); dev_free(&ctx->d_texidx);
    dev_free(&ctx->d_lpos); dev_free(&ctx->d_lnrm); dev_free(&ctx->d_lrad); dev_free(&ctx->d_lrgb);
    dev_free(&ctx->d_lloc); dev_free(&ctx->d_linst); dev_free(&ctx->d_lnmap);
    dev_free(&ctx->d_fvis); dev_free(&ctx->d_smask); dev_free(&ctx->d_active); dev_free(&ctx->d_active_count);
    dev_free(&ctx->d_rad_rowoff); dev_free(&ctx->d_rad_other); dev_free(&ctx->d_rad_factor); dev_free(&ctx->d_rad_sidx);
    dev_free(&ctx->d_image); dev_free(&ctx->d_image_tmp); dev_free(&ctx->d_mask); dev_free(&ctx->d_mask_tmp);
    dev_free(&ctx->d_normals); dev_free(&ctx->d_out);
    ctx->n_lumels = 0; ctx->rad_rows = ctx->rad_links = 0;
}

extern "C" void ltrgpu_destroy(ltrgpu_Ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_bake_state(ctx);
    dev_free(&ctx->d_inst); dev_free(&ctx->d_wpos); dev_free(&ctx->d_wnrm); dev_free(&ctx->d_vtex); dev_free(&ctx->d_ltex);
    dev_free(&ctx->d_rtris); dev_free(&ctx->d_rnodes); dev_free(&ctx->d_ritems); dev_free(&ctx->d_rtree_tris); dev_free(&ctx->d_rtree_ptris); dev_free(&ctx->d_rtree_boxes);
    if (ctx->lbvh_owned) { lb_free(ctx->d_lbvh); lb_free(ctx->d_lbvh_ptris); }
    dev_free(&ctx->d_bvh); dev_free(&ctx->d_bvh4); dev_free(&ctx->d_bvh4q); dev_free(&ctx->d_ptris); dev_free(&ctx->d_raytris); dev_free(&ctx->d_tri_orig);
    dev_free(&ctx->d_lights); dev_free(&ctx->d_light_inst); dev_free(&ctx->d_light_samples); dev_free(&ctx->d_probe_pos); dev_free(&ctx->d_probe_nrm);
    dev_free(&ctx->d_ao_cos); dev_free(&ctx->d_ao_sin); dev_free(&ctx->d_blur_kernel); dev_free(&ctx->d_counters);
    free(ctx->h_inst); free(ctx->h_lights); free(ctx->h_inst_lumel_off);
    free(ctx->h_out_off); free(ctx->h_out_w); free(ctx->h_out_h);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_span0) cudaEventDestroy(ctx->ev_span0);
    if (ctx->ev_span1) cudaEventDestroy(ctx->ev_span1);
    if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
    if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
    for (int b = 0; b < 2; ++b) { ltrgpu_host_free(ctx->stage_buf[b]); if (ctx->stage_ev[b]) cudaEventDestroy(ctx->stage_ev[b]); }
    for (int b = 0; b < 2; ++b) { if (ctx->d_req[b]) lb_free(ctx->d_req[b]); if (ctx->ev_req[b]) cudaEventDestroy(ctx->ev_req[b]); }
    if (ctx->ev_lumels) cudaEventDestroy(ctx->ev_lumels);
    if (ctx->ev_early) cudaEventDestroy(ctx->ev_early);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    /* cached device blocks stay in the pool for the next bake of this process (ltrgpu_release_memory drops them) */
    delete ctx;
}

extern "C" void ltrgpu_release_memory(void)
{
    lb_trim();
    HostPool &P = host_pool();
    std::lock_guard<std::mutex> g(P.mu);
    for (auto &kv : P.idle) cudaFreeHost(kv.second);
    P.idle.clear();
}
extern "C" const char *ltrgpu_last_error(ltrgpu_Ctx *ctx) { return ctx ? ctx->err : "no GPU context"; }
extern "C" void *ltrgpu_stream(ltrgpu_Ctx *ctx) { return (void *)ctx->stream; }

extern "C" int ltrgpu_sync(ltrgpu_Ctx *ctx)
{
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int ltrgpu_span_begin(ltrgpu_Ctx *ctx)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaEventRecord(ctx->ev_span0, ctx->stream));
    return 0;
}

extern "C" int ltrgpu_span_end(ltrgpu_Ctx *ctx)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaEventRecord(ctx->ev_span1, ctx->stream));
    CU_TRY(ctx, cudaEventSynchronize(ctx->ev_span1));
    float ms = 0;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_span0, ctx->ev_span1));
    ctx->host_counters.ms_span = ms;
    return 0;
}

extern "C" int ltrgpu_reset_bake(ltrgpu_Ctx *ctx)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    free_bake_state(ctx);
    CU_TRY(ctx, cudaMemsetAsync(ctx->d_counters, 0, CNT_COUNT * sizeof(unsigned long long), ctx->stream));
    uint64_t h2d = ctx->host_counters.h2d_bytes;
    memset(&ctx->host_counters, 0, sizeof(ctx->host_counters));
    ctx->host_counters.h2d_bytes = h2d;          /* the scene upload still belongs to this bake */
    return 0;
}

/* one concatenated per-instance array: allocate all of it, send this rank's slice, all-gather the rest */
template <class T> static int upload_sharded(ltrgpu_Ctx *ctx, T **p, const T *host, size_t total, size_t per_elem /* Ts per element */, const uint64_t *shard)
{
    if (!shard || ctx->world <= 1) return dev_upload(ctx, p, host, total * per_elem);
    if (dev_alloc(ctx, p, total * per_elem)) return 1;
    const uint64_t b = shard[ctx->rank] * per_elem, e = shard[ctx->rank + 1] * per_elem;
    if (e > b) {
        if (lb_upload_staged(ctx, *p + b, host + b, (e - b) * sizeof(T))) return 1;
        ctx->host_counters.h2d_bytes += (e - b) * sizeof(T);
    }
    if (!ctx->gatherv) { snprintf(ctx->err, sizeof(ctx->err), "sharded scene upload without a gather hook"); return 1; }
    uint64_t off[65];
    if (ctx->world > 64) { snprintf(ctx->err, sizeof(ctx->err), "more than 64 ranks"); return 1; }
    for (int r = 0; r <= ctx->world; ++r) off[r] = shard[r] * per_elem * sizeof(T);
    if (ctx->gatherv(ctx->allgather_user, *p, off, ctx->stream)) { snprintf(ctx->err, sizeof(ctx->err), "scene upload: all-gather of the instance slices failed"); return 1; }
    return 0;
}

static int ensure_aux_stream(ltrgpu_Ctx *ctx)
{
    if (!ctx->aux_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    if (!ctx->ev_early) CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_early, cudaEventDisableTiming));
    return 0;
}

extern "C" int ltrgpu_upload_tris_early(ltrgpu_Ctx *ctx, const float *rtree_tris9, uint32_t n_rtree_tris, const uint64_t *shard_tris)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->early_tris = 0; ctx->early_bvh = false;
    if (!n_rtree_tris) return 0;
    if (ensure_aux_stream(ctx)) return 1;
    if (upload_sharded(ctx, &ctx->d_rtree_tris, rtree_tris9, n_rtree_tris, 9, shard_tris)) return 1;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));           /* the second stream reads them next */
    ctx->early_tris = n_rtree_tris;
    return 0;
}

extern "C" int ltrgpu_build_bvh_early(ltrgpu_Ctx *ctx, int bvh_leaf_max)
{
    /* errors go to early_err: the bake thread may be writing ctx->err at the same time */
#define EARLY_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(ctx->early_err, sizeof(ctx->early_err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return 1; } } while (0)
    EARLY_TRY(cudaSetDevice(ctx->device));
    ctx->early_err[0] = 0;
    const uint32_t n = ctx->early_tris;
    if (!n) { snprintf(ctx->early_err, sizeof(ctx->early_err), "early BVH build without triangles"); return 1; }
    cudaStream_t st = ctx->aux_stream;
    if (ctx->d_bvh) { lb_free(ctx->d_bvh); ctx->d_bvh = nullptr; }
    if (ctx->d_bvh4) { lb_free(ctx->d_bvh4); ctx->d_bvh4 = nullptr; }
    if (ctx->d_tri_orig) { lb_free(ctx->d_tri_orig); ctx->d_tri_orig = nullptr; }
    if (ctx->d_ptris) { lb_free(ctx->d_ptris); ctx->d_ptris = nullptr; }
    if (ctx->d_raytris) { lb_free(ctx->d_raytris); ctx->d_raytris = nullptr; }
    EARLY_TRY(lb_malloc(&ctx->d_ptris, (size_t)n * sizeof(PreparedTri)));
    EARLY_TRY(lb_malloc(&ctx->d_raytris, (size_t)n * sizeof(RayTri)));
    LbDeviceBvh T;
    cudaEvent_t e0 = nullptr, e1 = nullptr;                    /* own timing events: ev_k0 / ev_k1 belong to the bake thread */
    EARLY_TRY(cudaEventCreate(&e0)); EARLY_TRY(cudaEventCreate(&e1));
    EARLY_TRY(cudaEventRecord(e0, st));
    if (lb_build_bvh_device(st, ctx->d_rtree_tris, n, bvh_leaf_max, ctx->num_sms, &T, ctx->early_err, sizeof(ctx->early_err))) return 1;
    EARLY_TRY(cudaEventRecord(e1, st));
    ctx->d_bvh = T.nodes; ctx->d_bvh4 = T.nodes4; ctx->d_tri_orig = T.order;
    ctx->n_bvh_nodes = T.n_nodes; ctx->bvh_height = T.height; ctx->n_bvh4_nodes = T.n_nodes4;
    ctx->early_launches = T.launches + 1;
    prepare_tris_kernel<<<grid_for(n, 256), 256, 0, st>>>(ctx->d_rtree_tris, ctx->d_tri_orig, n, ctx->d_ptris, ctx->d_raytris);
    EARLY_TRY(cudaGetLastError());
    EARLY_TRY(cudaEventRecord(ctx->ev_early, st));
    EARLY_TRY(cudaEventSynchronize(e1));
    EARLY_TRY(cudaEventElapsedTime(&ctx->bvh_build_ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->early_bvh = true;
    return 0;
#undef EARLY_TRY
}

extern "C" const char *ltrgpu_early_error(ltrgpu_Ctx *ctx) { return ctx->early_err; }

extern "C" int ltrgpu_upload_scene(ltrgpu_Ctx *ctx, const ltrgpu_SceneDesc *d)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const uint32_t early_nodes = ctx->n_bvh_nodes;             /* set by ltrgpu_build_bvh_early, if it ran */
    ctx->params = d->params;
    ctx->n_inst = d->n_inst; ctx->n_verts = d->n_verts; ctx->n_rtris = d->n_rtris;
    ctx->n_rnodes = d->n_rnodes; ctx->n_ritems = d->n_ritems; ctx->n_rtree_tris = d->n_rtree_tris;
    ctx->n_bvh_nodes = d->n_bvh_nodes; ctx->n_tris = d->n_tris; ctx->n_lights = d->n_lights; ctx->n_probes = d->n_probes;
    ctx->blur_ext = d->blur_ext;

    free(ctx->h_inst);
    ctx->h_inst = (ltrgpu_Inst *)malloc(sizeof(ltrgpu_Inst) * (d->n_inst ? d->n_inst : 1));
    memcpy(ctx->h_inst, d->inst, sizeof(ltrgpu_Inst) * d->n_inst);
    free(ctx->h_lights);
    ctx->h_lights = (ltrgpu_Light *)malloc(sizeof(ltrgpu_Light) * (d->n_lights ? d->n_lights : 1));
    if (d->n_lights) memcpy(ctx->h_lights, d->lights, sizeof(ltrgpu_Light) * d->n_lights);
    ctx->n_texels = 0;
    for (uint32_t i = 0; i < d->n_inst; ++i) ctx->n_texels += (uint64_t)d->inst[i].lm_w * d->inst[i].lm_h;

    if (dev_upload(ctx, &ctx->d_inst, d->inst, d->n_inst)) return 1;
    if (upload_sharded(ctx, &ctx->d_wpos, d->wpos, d->n_verts, 1, d->shard_verts)) return 1;
    if (upload_sharded(ctx, &ctx->d_wnrm, d->wnrm, d->n_verts, 1, d->shard_verts)) return 1;
    if (upload_sharded(ctx, (float **)&ctx->d_vtex, d->vtex2, d->n_verts, 2, d->shard_verts)) return 1;
    if (upload_sharded(ctx, (float **)&ctx->d_ltex, d->ltex2, d->n_verts, 2, d->shard_verts)) return 1;
    if (upload_sharded(ctx, &ctx->d_rtris, d->rtris, d->n_rtris, 1, d->shard_rtris)) return 1;
    if (upload_sharded(ctx, &ctx->d_rnodes, d->rnodes, d->n_rnodes, 1, d->shard_rnodes)) return 1;
    if (upload_sharded(ctx, &ctx->d_ritems, d->ritems, d->n_ritems, 1, d->shard_ritems)) return 1;
    const bool tris_early = ctx->early_tris && ctx->early_tris == d->n_rtree_tris;      /* already up (ltrgpu_upload_tris_early) */
    if (!tris_early && upload_sharded(ctx, &ctx->d_rtree_tris, d->rtree_tris9, d->n_rtree_tris, 9, d->shard_tris)) return 1;
    if (dev_upload(ctx, &ctx->d_lights, d->lights, d->n_lights)) return 1;
    if (dev_upload(ctx, &ctx->d_light_inst, d->light_inst, (size_t)d->n_lights * d->n_inst)) return 1;
    if (dev_upload(ctx, &ctx->d_light_samples, d->light_samples4, d->n_light_samples)) return 1;
    if (dev_upload(ctx, &ctx->d_probe_pos, d->probe_pos, d->n_probes)) return 1;
    if (dev_upload(ctx, &ctx->d_probe_nrm, d->probe_nrm, d->n_probes)) return 1;
    int ns = d->params.ao_num_samples > 0 ? d->params.ao_num_samples : 0;
    if (dev_upload(ctx, &ctx->d_ao_cos, d->ao_cos_side, ns)) return 1;
    if (dev_upload(ctx, &ctx->d_ao_sin, d->ao_sin_side, ns)) return 1;
    if (dev_upload(ctx, &ctx->d_blur_kernel, d->blur_kernel, d->blur_kernel ? 2 * d->blur_ext + 1 : 0)) return 1;

    /* scene BVH + its triangles: raw triangles up, expanded on the device into the two prepared records, raw dropped */
    float *d_raw = nullptr;
    const bool raw_is_rtree = d->bvh == nullptr && d->tris9 == d->rtree_tris9 && d->n_tris == d->n_rtree_tris;
    const bool bvh_early = tris_early && ctx->early_bvh && raw_is_rtree;                /* built beside the host pre-pass (ltrgpu_build_bvh_early) */
    ctx->early_tris = 0; ctx->early_bvh = false;
    if (raw_is_rtree) d_raw = ctx->d_rtree_tris;
    else if (dev_upload(ctx, &d_raw, d->tris9, (size_t)d->n_tris * 9)) return 1;
    if (!bvh_early) {
        if (dev_alloc(ctx, &ctx->d_ptris, d->n_tris)) return 1;
        if (dev_alloc(ctx, &ctx->d_raytris, d->n_tris)) return 1;
        ctx->bvh_build_ms = 0.f;
    }
    if (bvh_early) {
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_early, 0));
        ctx->host_counters.kernel_launches += ctx->early_launches;
        ctx->n_bvh_nodes = early_nodes;
    } else if (d->bvh) {
        if (dev_upload(ctx, &ctx->d_bvh, d->bvh, d->n_bvh_nodes)) return 1;
        if (dev_upload(ctx, &ctx->d_bvh4, d->bvh4, d->n_bvh4_nodes)) return 1;
        if (dev_upload(ctx, &ctx->d_tri_orig, d->tri_orig, d->n_tris)) return 1;
        ctx->bvh_height = d->bvh_height; ctx->n_bvh4_nodes = d->n_bvh4_nodes;
        if (d->n_tris) {
            prepare_tris_kernel<<<grid_for(d->n_tris, 256), 256, 0, ctx->stream>>>(d_raw, nullptr, d->n_tris, ctx->d_ptris, ctx->d_raytris);
            CU_LAUNCH_CHECK(ctx);
        }
    } else {
        dev_free(&ctx->d_bvh); dev_free(&ctx->d_bvh4); dev_free(&ctx->d_tri_orig);
        LbDeviceBvh T;
        CU_TRY(ctx, cudaEventRecord(ctx->ev_k0, ctx->stream));
        if (lb_build_bvh_device(ctx->stream, d_raw, d->n_tris, d->bvh_leaf_max, ctx->num_sms, &T, ctx->err, sizeof(ctx->err))) return 1;
        CU_TRY(ctx, cudaEventRecord(ctx->ev_k1, ctx->stream));
        ctx->d_bvh = T.nodes; ctx->d_bvh4 = T.nodes4; ctx->d_tri_orig = T.order;
        ctx->n_bvh_nodes = T.n_nodes; ctx->bvh_height = T.height; ctx->n_bvh4_nodes = T.n_nodes4;
        ctx->host_counters.kernel_launches += T.launches;
        prepare_tris_kernel<<<grid_for(d->n_tris, 256), 256, 0, ctx->stream>>>(d_raw, ctx->d_tri_orig, d->n_tris, ctx->d_ptris, ctx->d_raytris);
        CU_LAUNCH_CHECK(ctx);
        CU_TRY(ctx, cudaEventSynchronize(ctx->ev_k1));
        CU_TRY(ctx, cudaEventElapsedTime(&ctx->bvh_build_ms, ctx->ev_k0, ctx->ev_k1));
    }
#if LB_VIS_Q8
    if (dev_alloc(ctx, &ctx->d_bvh4q, ctx->n_bvh4_nodes)) return 1;
    if (ctx->n_bvh4_nodes) {
        quantise_bvh4_kernel<<<grid_for(ctx->n_bvh4_nodes, 128), 128, 0, ctx->stream>>>(ctx->d_bvh4, ctx->n_bvh4_nodes, ctx->d_bvh4q);
        CU_LAUNCH_CHECK(ctx);
    }
#endif
    /* the traversal stacks are fixed (BVH_STACK): binary walks push one node per level, the 4-wide walk up to three per two levels */
    if (ctx->bvh_height > 40) {
        snprintf(ctx->err, sizeof(ctx->err), "scene BVH is %d levels deep; the traversal stacks hold 40", ctx->bvh_height);
        return 1;
    }
    /* flat BVH over the triangles of every instance tree, for lumel_classify_kernel (gpu_lumels.cu) */
    if (ctx->lbvh_owned) { lb_free(ctx->d_lbvh); lb_free(ctx->d_lbvh_ptris); }
    ctx->d_lbvh = nullptr; ctx->d_lbvh_ptris = nullptr; ctx->lbvh_owned = false;
    if (d->scene_covers_rtree) { ctx->d_lbvh = ctx->d_bvh; ctx->d_lbvh_ptris = ctx->d_ptris; }
    else if (d->n_rtree_tris > (uint32_t)(d->bvh_leaf_max > 0 ? d->bvh_leaf_max : BVH_LEAF_MAX)) {
        LbDeviceBvh T;
        if (lb_build_bvh_device(ctx->stream, ctx->d_rtree_tris, d->n_rtree_tris, d->bvh_leaf_max > 0 ? d->bvh_leaf_max : BVH_LEAF_MAX, ctx->num_sms, &T, ctx->err, sizeof(ctx->err))) return 1;
        ctx->host_counters.kernel_launches += T.launches;
        ctx->d_lbvh = T.nodes; ctx->lbvh_owned = true;
        CU_TRY(ctx, lb_malloc(&ctx->d_lbvh_ptris, (size_t)d->n_rtree_tris * sizeof(PreparedTri)));
        prepare_tris_kernel<<<grid_for(d->n_rtree_tris, 256), 256, 0, ctx->stream>>>(ctx->d_rtree_tris, T.order, d->n_rtree_tris, ctx->d_lbvh_ptris, nullptr);
        CU_LAUNCH_CHECK(ctx);
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        lb_free(T.nodes4); lb_free(T.order);
    }
    /* reference-order triangles (lumel generation): point-query terms precomputed once instead of per query */
    if (dev_alloc(ctx, &ctx->d_rtree_ptris, d->n_rtree_tris)) return 1;
    if (dev_alloc(ctx, &ctx->d_rtree_boxes, (size_t)d->n_rtree_tris * 2)) return 1;
    if (d->n_rtree_tris) {
        tri_boxes_kernel<<<grid_for(d->n_rtree_tris, 256), 256, 0, ctx->stream>>>(ctx->d_rtree_tris, d->n_rtree_tris, ctx->d_rtree_boxes);
        CU_LAUNCH_CHECK(ctx);
        prepare_tris_kernel<<<grid_for(d->n_rtree_tris, 256), 256, 0, ctx->stream>>>(ctx->d_rtree_tris, nullptr, d->n_rtree_tris, ctx->d_rtree_ptris, nullptr);
        CU_LAUNCH_CHECK(ctx);
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!raw_is_rtree) lb_free(d_raw);
    return 0;
}

extern "C" int ltrgpu_host_allgather(ltrgpu_Ctx *ctx, const void *send, void *recv, size_t bytes)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->world <= 1) { memcpy(recv, send, bytes); return 0; }
    if (!ctx->allgather) { snprintf(ctx->err, sizeof(ctx->err), "no all-gather hook"); return 1; }
    const size_t padded = (bytes + 15) & ~(size_t)15;
    char *d = nullptr;
    if (dev_alloc(ctx, &d, padded * ctx->world)) return 1;
    CU_TRY(ctx, cudaMemcpyAsync(d + padded * ctx->rank, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->allgather(ctx->allgather_user, d + padded * ctx->rank, d, padded, ctx->stream)) { lb_free(d); snprintf(ctx->err, sizeof(ctx->err), "host table all-gather failed"); return 1; }
    for (int r = 0; r < ctx->world; ++r) CU_TRY(ctx, cudaMemcpyAsync((char *)recv + bytes * r, d + padded * r, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    lb_free(d);
    return 0;
}

extern "C" int ltrgpu_bvh_info(ltrgpu_Ctx *ctx, uint32_t *n_nodes, int *height, float *build_ms)
{
    if (n_nodes) *n_nodes = ctx->n_bvh_nodes;
    if (height) *height = ctx->bvh_height;
    if (build_ms) *build_ms = ctx->bvh_build_ms;
    return 0;
}

extern "C" int ltrgpu_set_world(ltrgpu_Ctx *ctx, int rank, int world, ltrgpu_allgather_fn allgather, void *allgather_user)
{
    if (world < 1 || rank < 0 || rank >= world) { snprintf(ctx->err, sizeof(ctx->err), "bad rank/world"); return 1; }
    ctx->rank = rank; ctx->world = world; ctx->allgather = allgather; ctx->allgather_user = allgather_user;
    return 0;
}

extern "C" int ltrgpu_set_gatherv(ltrgpu_Ctx *ctx, ltrgpu_gatherv_fn gatherv) { ctx->gatherv = gatherv; return 0; }
extern "C" int ltrgpu_set_alltoallv(ltrgpu_Ctx *ctx, ltrgpu_alltoallv_fn fn) { ctx->alltoallv = fn; return 0; }

extern "C" int ltrgpu_set_shard(ltrgpu_Ctx *ctx, uint64_t begin, uint64_t end, int rank, int world,
                                ltrgpu_allgather_fn allgather, void *allgather_user)
{
    if (end > ctx->n_lumels || begin > end) { snprintf(ctx->err, sizeof(ctx->err), "bad shard range"); return 1; }
    ctx->sh_begin = begin; ctx->sh_end = end; ctx->rank = rank; ctx->world = world;
    ctx->allgather = allgather; ctx->allgather_user = allgather_user;
    return 0;
}

extern "C" int ltrgpu_get_counters(ltrgpu_Ctx *ctx, ltrgpu_Counters *out)
{
    unsigned long long c[CNT_COUNT];
    CU_TRY(ctx, cudaMemcpyAsync(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ltrgpu_Counters &h = ctx->host_counters;
    h.kernel_launches += ctx->aux_launches; h.d2h_bytes += ctx->aux_d2h_bytes;     /* sample_fn batching (second stream; its thread has been joined by now) */
    ctx->aux_launches = ctx->aux_d2h_bytes = 0;
    h.marches = c[CNT_MARCHES]; h.distance_queries = c[CNT_DIST_QUERIES]; h.ao_segments = c[CNT_AO_SEGMENTS];
    h.correction_rays = c[CNT_CORR_RAYS]; h.rad_pairs = c[CNT_RAD_PAIRS]; h.rad_segments = c[CNT_RAD_SEGMENTS];
    h.rad_links = c[CNT_RAD_LINKS]; h.node_visits = c[CNT_NODE_VISITS]; h.tri_tests = c[CNT_TRI_TESTS];
    h.ray_node_visits = c[CNT_RAY_NODE_VISITS]; h.ray_tri_tests = c[CNT_RAY_TRI_TESTS]; h.rad_tile_loads = c[CNT_RAD_TILE_LOADS]; h.shadow_rays = c[CNT_SHADOW_RAYS]; h.ray_entry_tests = c[CNT_RAY_ENTRY_TESTS];
    *out = h;
    return 0;
}
