/*
 * gpu_test.cu -- ltrx_test_* entry points: run ONE device primitive / query over arrays of inputs so
 * that tests can compare the CUDA path with the oracle primitive by primitive.  Host arrays in,
 * host arrays out.  These go through the same device functions the stage kernels use.
 */
#include "gpu_internal.cuh"

#include <vector>

#include "lighter_b200.h"

__device__ __forceinline__ float march_shadow_test(const BvhNode *bvh, const PreparedTri *tris, V3 from, V3 to, float k, unsigned &queries)
{
    TravStats ts = { 0, 0 };
    V3 rd = norm3(to - from);
    float maxt = len3(to - from);
    float res = 1.0f;
    for (float t = 0.001f; t < maxt;) {
        float h = bvh_distance(bvh, tris, from + rd * t, 2.0f, 0.001f, ts);
        ++queries;
        if (h < 0.001f) return 0.0f;
        res = fminr(res, h / fminr(t * k, 2.0f));
        h = fminr(h, 1.0f);
        t += h;
    }
    return res;
}

__global__ void t_ptd_kernel(const float *pts, const float *tris, uint32_t n, float *out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *t = tris + 9ull * i;
    out[i] = point_tri_distance(mk3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), mk3(t[0], t[1], t[2]), mk3(t[3], t[4], t[5]), mk3(t[6], t[7], t[8]));
}

__global__ void t_segtri_kernel(const float *a, const float *b, const float *tris, uint32_t n, float *out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *t = tris + 9ull * i;
    out[i] = seg_tri(mk3(a[3 * i], a[3 * i + 1], a[3 * i + 2]), mk3(b[3 * i], b[3 * i + 1], b[3 * i + 2]), mk3(t[0], t[1], t[2]),
                     mk3(t[3], t[4], t[5]), mk3(t[6], t[7], t[8]));
}

__global__ void t_queries_kernel(const BvhNode *bvh, const Bvh4Node *bvh4, const PreparedTri *pt, const RayTri *rt, const uint32_t *orig, const float *a, const float *b,
                                 uint32_t n, float *dist, int *anyhit, float *closest, int *closest_tri)
{
    __shared__ BvhEntrySet s_entry[4], s_entry1[4];  /* launched with 128 threads */
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const uint32_t q = valid ? i : 0;                /* idle lanes of the last warp stay for the warp-wide reduction below */
    TravStats ts = { 0, 0 };
    V3 A = mk3(a[3 * q], a[3 * q + 1], a[3 * q + 2]), B = mk3(b[3 * q], b[3 * q + 1], b[3 * q + 2]);
    if (dist && valid) dist[i] = bvh_distance(bvh, pt, A, 2.0f, -1.0f, ts);
    if (anyhit) {
        /* the bundle of this warp's 32 segments: box, entry set (bvh_entry.h), walk from the entry set */
        float lx = fminf(A.x, B.x), ly = fminf(A.y, B.y), lz = fminf(A.z, B.z), hx = fmaxf(A.x, B.x), hy = fmaxf(A.y, B.y), hz = fmaxf(A.z, B.z);
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
            hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
        }
        BvhEntrySet &E = s_entry[threadIdx.x >> 5], &E1 = s_entry1[threadIdx.x >> 5];
        bvh_entry_pad(lx, ly, lz, hx, hy, hz);
        if ((threadIdx.x & 31u) == 0) bvh4_entry_search(bvh4, lx, ly, lz, hx, hy, hz, E1);       /* the scalar model (also run on the host) */
        bvh_entry_search_warp<Bvh4Access>(bvh4, lx, ly, lz, hx, hy, hz, E, threadIdx.x & 31u);   /* the form the kernels use */
        bool same = E.n == E1.n;
        for (int e = 0; same && e < E.n; ++e)
            same = E.node[e] == E1.node[e] && E.lox[e] == E1.lox[e] && E.loy[e] == E1.loy[e] && E.loz[e] == E1.loz[e] && E.hix[e] == E1.hix[e] && E.hiy[e] == E1.hiy[e] && E.hiz[e] == E1.hiz[e];
        /* four implementations of the same predicate: ordered segment walk, two-phase binary walk, two-phase 4-wide walk
         * from the root and from the bundle's entry set (the one the radiosity kernel uses); a disagreement is reported as 2 + bits */
        const int h0 = bvh_segment<true>(bvh, rt, orig, A, B, nullptr, ts) < 1.0f ? 1 : 0;
        const int h1 = bvh_anyhit(bvh, rt, A, B, ts) ? 1 : 0, h2 = bvh4_anyhit(bvh4, rt, A, B, ts) ? 1 : 0, h3 = bvh4_anyhit<2>(bvh4, rt, A, B, ts) ? 1 : 0;
        const int h4 = bvh4_anyhit_entries(bvh4, rt, E, A, B, ts) ? 1 : 0;
        if (valid) anyhit[i] = (h0 == h1 && h1 == h2 && h2 == h3 && h3 == h4 && same) ? h0 : 2 + h0 + 2 * h1 + 4 * h2 + 8 * h3 + 16 * h4 + (same ? 0 : 32);
    }
    if (closest && valid) {
        int slot = -1;
        closest[i] = bvh_segment<false>(bvh, rt, orig, A, B, &slot, ts);
        if (closest_tri) closest_tri[i] = slot >= 0 ? (int)orig[slot] : -1;
    }
}

__global__ void t_march_kernel(const BvhNode *bvh, const PreparedTri *pt, const float *from, const float *to, const float *k, uint32_t n,
                               float *out, uint32_t *steps)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned q = 0;
    out[i] = march_shadow_test(bvh, pt, mk3(from[3 * i], from[3 * i + 1], from[3 * i + 2]), mk3(to[3 * i], to[3 * i + 1], to[3 * i + 2]), k[i], q);
    if (steps) steps[i] = q;
}

__global__ void t_spiral_kernel(const float *nrm, const float *randoff, uint32_t n, int samples, const float *cs, const float *sn, float *out)
{
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * (uint32_t)samples) return;
    uint32_t i = e / samples;
    int s = (int)(e % samples);
    const float golden = 137.508f / 180.0f * 3.14159274101257324f;
    V3 dir = mk3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
    float angle = ((float)s + randoff[i]) * golden;
    float ca = ref_cosf(angle), sa = ref_sinf(angle);
    V3 diffvec = mk3(dir.y, -dir.z, dir.x);
    V3 up = norm3(cross3(dir, diffvec));
    V3 rt = cross3(dir, up);
    V3 r = ca * sn[s] * rt + sa * sn[s] * up + cs[s] * dir;
    out[3ull * e] = r.x; out[3ull * e + 1] = r.y; out[3ull * e + 2] = r.z;
}

namespace {

struct Dev {                       /* tiny RAII bag of device allocations for the test entry points */
    std::vector<void *> ptrs;
    bool ok = true;
    template <class T> T *up(const T *h, size_t n)
    {
        T *d = nullptr;
        if (cudaMalloc((void **)&d, (n ? n : 1) * sizeof(T)) != cudaSuccess) { ok = false; return nullptr; }
        ptrs.push_back(d);
        if (h && n && cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
        return d;
    }
    template <class T> T *alloc(size_t n) { return up<T>(nullptr, n); }
    template <class T> void down(T *h, const T *d, size_t n) { if (h && n && cudaMemcpy(h, d, n * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) ok = false; }
    ~Dev() { for (void *p : ptrs) cudaFree(p); }
};

bool have_device()
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess || c == 0) {
        fprintf(stderr, "lighter_b200: ltrx_test_*: no CUDA device (no CPU fallback)\n");
        return false;
    }
    return true;
}

struct SceneOnDevice {
    BvhNode *bvh; Bvh4Node *bvh4; PreparedTri *pt; RayTri *rt; uint32_t *orig;
};

__global__ void t_prepare_kernel(const float *tris9, uint32_t n, PreparedTri *pt, RayTri *rt)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *t = tris9 + 9ull * i;
    V3 a = mk3(t[0], t[1], t[2]), b = mk3(t[3], t[4], t[5]), c = mk3(t[6], t[7], t[8]);
    PreparedTri P; prepare_tri(a, b, c, P); pt[i] = P;
    RayTri R; prepare_raytri(a, b, c, R); rt[i] = R;
}

bool scene_to_device(Dev &D, const float *tris9, uint32_t ntris, SceneOnDevice &S)
{
    SceneBvh bvh;
    int leaf_max = BVH_LEAF_MAX;
    if (const char *e = getenv("LTR_BVH_LEAF")) leaf_max = atoi(e);
    build_scene_bvh(tris9, ntris, bvh, leaf_max, 0);
    std::vector<float> ordered((size_t)ntris * 9);
    for (uint32_t k = 0; k < ntris; ++k) memcpy(&ordered[(size_t)k * 9], tris9 + (size_t)bvh.order[k] * 9, 36);
    float *d_raw = D.up(ordered.data(), ordered.size());
    S.bvh = D.up(bvh.nodes.data(), bvh.nodes.size());
    S.bvh4 = D.up(bvh.nodes4.data(), bvh.nodes4.size());
    S.orig = D.up(bvh.order.data(), bvh.order.size());
    S.pt = D.alloc<PreparedTri>(ntris);
    S.rt = D.alloc<RayTri>(ntris);
    if (!D.ok) return false;
    if (ntris) t_prepare_kernel<<<(ntris + 255) / 256, 256>>>(d_raw, ntris, S.pt, S.rt);
    return cudaDeviceSynchronize() == cudaSuccess;
}

} // namespace

extern "C" {

int ltrx_test_point_tri_distance(const float *pts3, const float *tris9, u32 n, float *out)
{
    if (!have_device()) return 0;
    Dev D;
    float *dp = D.up(pts3, (size_t)n * 3), *dt = D.up(tris9, (size_t)n * 9), *dout = D.alloc<float>(n);
    if (!D.ok) return 0;
    if (n) t_ptd_kernel<<<(n + 255) / 256, 256>>>(dp, dt, n, dout);
    if (cudaDeviceSynchronize() != cudaSuccess) return 0;
    D.down(out, dout, n);
    return D.ok;
}

int ltrx_test_seg_tri(const float *a3, const float *b3, const float *tris9, u32 n, float *out)
{
    if (!have_device()) return 0;
    Dev D;
    float *da = D.up(a3, (size_t)n * 3), *db = D.up(b3, (size_t)n * 3), *dt = D.up(tris9, (size_t)n * 9), *dout = D.alloc<float>(n);
    if (!D.ok) return 0;
    if (n) t_segtri_kernel<<<(n + 255) / 256, 256>>>(da, db, dt, n, dout);
    if (cudaDeviceSynchronize() != cudaSuccess) return 0;
    D.down(out, dout, n);
    return D.ok;
}

int ltrx_test_scene_queries(const float *tris9, u32 ntris, const float *a3, const float *b3, u32 n, float *dist_out, int *anyhit_out,
                            float *closest_out, int *closest_tri_out)
{
    if (!have_device()) return 0;
    Dev D;
    SceneOnDevice S;
    if (!scene_to_device(D, tris9, ntris, S)) return 0;
    float *da = D.up(a3, (size_t)n * 3), *db = D.up(b3, (size_t)n * 3);
    float *dd = dist_out ? D.alloc<float>(n) : nullptr, *dc = closest_out ? D.alloc<float>(n) : nullptr;
    int *dh = anyhit_out ? D.alloc<int>(n) : nullptr, *dct = closest_tri_out ? D.alloc<int>(n) : nullptr;
    if (!D.ok) return 0;
    if (n) t_queries_kernel<<<(n + 127) / 128, 128>>>(S.bvh, S.bvh4, S.pt, S.rt, S.orig, da, db, n, dd, dh, dc, dct);
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "ltrx_test_scene_queries: %s\n", cudaGetErrorString(cudaGetLastError())); return 0; }
    D.down(dist_out, dd, n); D.down(anyhit_out, dh, n); D.down(closest_out, dc, n); D.down(closest_tri_out, dct, n);
    return D.ok;
}

int ltrx_test_march(const float *tris9, u32 ntris, const float *from3, const float *to3, const float *k, u32 n, float *out, u32 *steps_out)
{
    if (!have_device()) return 0;
    Dev D;
    SceneOnDevice S;
    if (!scene_to_device(D, tris9, ntris, S)) return 0;
    float *df = D.up(from3, (size_t)n * 3), *dt = D.up(to3, (size_t)n * 3), *dk = D.up(k, n), *dout = D.alloc<float>(n);
    u32 *ds = steps_out ? D.alloc<u32>(n) : nullptr;
    if (!D.ok) return 0;
    if (n) t_march_kernel<<<(n + 127) / 128, 128>>>(S.bvh, S.pt, df, dt, dk, n, dout, ds);
    if (cudaDeviceSynchronize() != cudaSuccess) return 0;
    D.down(out, dout, n); D.down(steps_out, ds, n);
    return D.ok;
}

int ltrx_test_spiral_dirs(const float *nrm3, const float *randoff, u32 n, int samples, float *out3)
{
    if (!have_device() || samples <= 0) return 0;
    std::vector<float> cs(samples), sn(samples);
    for (int s = 0; s < samples; ++s) { float q = (s + 0.5f) / samples; cs[s] = sqrtf(q); sn[s] = sinf(acosf(cs[s])); }
    Dev D;
    float *dn = D.up(nrm3, (size_t)n * 3), *dr = D.up(randoff, n), *dcs = D.up(cs.data(), cs.size()), *dsn = D.up(sn.data(), sn.size());
    float *dout = D.alloc<float>((size_t)n * samples * 3);
    if (!D.ok) return 0;
    const uint32_t total = n * (uint32_t)samples;
    if (total) t_spiral_kernel<<<(total + 255) / 256, 256>>>(dn, dr, n, samples, dcs, dsn, dout);
    if (cudaDeviceSynchronize() != cudaSuccess) return 0;
    D.down(out3, dout, (size_t)n * samples * 3);
    return D.ok;
}

} /* extern "C" */
