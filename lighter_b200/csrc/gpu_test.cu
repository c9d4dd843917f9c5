/*
 * gpu_test.cu -- ltrx_test_* entry points: run ONE device primitive / query over arrays of inputs so
 * that tests can compare the CUDA path with the oracle primitive by primitive.  Host arrays in,
 * host arrays out.  These go through the same device functions the stage kernels use.
 */
#include "gpu_internal.cuh"

#include <algorithm>
#include <vector>

#include "lighter_b200.h"

/* the production march (gpu_internal.cuh march_shadow) */
__device__ __forceinline__ float march_shadow_test(const BvhNode *bvh, const Bvh4Node *bvh4, const PreparedTri *tris, V3 from, V3 to, float k, unsigned &queries)
{
    TravStats ts = { 0, 0 };
    return march_shadow(bvh, bvh4, tris, from, to, k, queries, ts);
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
    __shared__ BvhEntrySet2 s_entry2[4], s_entry2m[4];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const uint32_t q = valid ? i : 0;                /* idle lanes of the last warp stay for the warp-wide reduction below */
    TravStats ts = { 0, 0 };
    V3 A = mk3(a[3 * q], a[3 * q + 1], a[3 * q + 2]), B = mk3(b[3 * q], b[3 * q + 1], b[3 * q + 2]);
    if (dist && valid) {
        const float db = bvh_distance(bvh, pt, A, 2.0f, -1.0f, ts), d4 = bvh4_distance(bvh4, pt, A, 2.0f, -1.0f, ts);
        dist[i] = db == d4 ? db : -1.0f;          /* the binary and the 4-wide walk must return the same float */
    }
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
        /* version-2 entry sets (rad_visibility_kernel): R = the first end points of the warp's segments, C = the second ones;
         * the scalar model against the warp-cooperative search, and the walk from that set */
        BvhEntrySet2 &F = s_entry2[threadIdx.x >> 5], &F1 = s_entry2m[threadIdx.x >> 5];
        float r[12] = { A.x, A.y, A.z, A.x, A.y, A.z, B.x, B.y, B.z, B.x, B.y, B.z };
        for (int o = 16; o > 0; o >>= 1)
            for (int k = 0; k < 12; ++k) { const float t = __shfl_xor_sync(0xffffffffu, r[k], o); r[k] = (k % 6) < 3 ? fminf(r[k], t) : fmaxf(r[k], t); }
        float qx = fminf(r[0], r[6]), qy = fminf(r[1], r[7]), qz = fminf(r[2], r[8]), Qx = fmaxf(r[3], r[9]), Qy = fmaxf(r[4], r[10]), Qz = fmaxf(r[5], r[11]);
        const float maxabs = fmaxf(fmaxf(fmaxf(fabsf(qx), fabsf(Qx)), fmaxf(fabsf(qy), fabsf(Qy))), fmaxf(fabsf(qz), fabsf(Qz)));
        bvh_entry_pad(qx, qy, qz, Qx, Qy, Qz);
        __syncwarp();
        if ((threadIdx.x & 31u) == 0) {
            for (int k = 0; k < 12; ++k) F.rc[k] = r[k];
            BvhShaft S;
            bvh_shaft_build(r, r + 6, maxabs, S);
            bvh4_entry_search2(bvh4, qx, qy, qz, Qx, Qy, Qz, &S, F1);
        }
        __syncwarp();
        bvh4_entry_search2_warp<true>(bvh4, qx, qy, qz, Qx, Qy, Qz, maxabs, F, threadIdx.x & 31u);
        bool same2 = F.n == F1.n;
        for (int e = 0; same2 && e < F.n; ++e)
            for (int k = 0; k < 4; ++k) same2 = same2 && __float_as_int(F.lo[e][k]) == __float_as_int(F1.lo[e][k]) && (k == 3 || F.hi[e][k] == F1.hi[e][k]);
        const int h5 = bvh4_anyhit_entries2(bvh4, rt, F, A, B, ts) ? 1 : 0;
        if (valid) anyhit[i] = (h0 == h1 && h1 == h2 && h2 == h3 && h3 == h4 && h4 == h5 && same && same2) ? h0
                               : 2 + h0 + 2 * h1 + 4 * h2 + 8 * h3 + 16 * h4 + (same ? 0 : 32) + 64 * h5 + (same2 ? 0 : 128);
    }
    if (closest && valid) {
        int slot = -1;
        closest[i] = bvh_segment<false>(bvh, rt, orig, A, B, &slot, ts);
        if (closest_tri) closest_tri[i] = slot >= 0 ? (int)orig[slot] : -1;
    }
}

__global__ void t_march_kernel(const BvhNode *bvh, const Bvh4Node *bvh4, const PreparedTri *pt, const float *from, const float *to, const float *k, uint32_t n,
                               float *out, uint32_t *steps)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned q = 0;
    out[i] = march_shadow_test(bvh, bvh4, pt, mk3(from[3 * i], from[3 * i + 1], from[3 * i + 2]), mk3(to[3 * i], to[3 * i + 1], to[3 * i + 2]), k[i], q);
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

__global__ void t_prepare_ordered_kernel(const float *tris9, const uint32_t *order, uint32_t n, PreparedTri *pt, RayTri *rt)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *t = tris9 + 9ull * order[i];
    V3 a = mk3(t[0], t[1], t[2]), b = mk3(t[3], t[4], t[5]), c = mk3(t[6], t[7], t[8]);
    PreparedTri P; prepare_tri(a, b, c, P); pt[i] = P;
    RayTri R; prepare_raytri(a, b, c, R); rt[i] = R;
}

struct DeviceTree {                /* a device-built tree kept alive for the duration of a test call */
    LbDeviceBvh T = { nullptr, nullptr, nullptr, 0, 0, 0, 0 };
    ~DeviceTree() { lb_free(T.nodes); lb_free(T.nodes4); lb_free(T.order); }
};

/* The scene the way a bake sets it up: the BVH built on the device (gpu_bvh.cu) unless the scene fits one leaf or
 * LTR_BVH_HOST=1 asks for the host builder (A/B). */
bool scene_to_device(Dev &D, DeviceTree &DT, const float *tris9, uint32_t ntris, SceneOnDevice &S)
{
    int leaf_max = BVH_LEAF_MAX;
    if (const char *e = getenv("LTR_BVH_LEAF")) leaf_max = atoi(e);
    if (leaf_max < 1) leaf_max = 1;
    if (leaf_max > 7) leaf_max = 7;
    S.pt = D.alloc<PreparedTri>(ntris);
    S.rt = D.alloc<RayTri>(ntris);
    if (ntris > (uint32_t)leaf_max && !getenv("LTR_BVH_HOST")) {
        float *d_raw = D.up(tris9, (size_t)ntris * 9);
        if (!D.ok) return false;
        char err[256] = { 0 };
        if (lb_build_bvh_device(nullptr, d_raw, ntris, leaf_max, 148, &DT.T, err, sizeof(err))) { fprintf(stderr, "lighter_b200: %s\n", err); return false; }
        S.bvh = DT.T.nodes; S.bvh4 = DT.T.nodes4; S.orig = DT.T.order;
        t_prepare_ordered_kernel<<<(ntris + 255) / 256, 256>>>(d_raw, DT.T.order, ntris, S.pt, S.rt);
        return cudaDeviceSynchronize() == cudaSuccess;
    }
    SceneBvh bvh;
    build_scene_bvh(tris9, ntris, bvh, leaf_max, 0);
    std::vector<float> ordered((size_t)ntris * 9);
    for (uint32_t k = 0; k < ntris; ++k) memcpy(&ordered[(size_t)k * 9], tris9 + (size_t)bvh.order[k] * 9, 36);
    float *d_raw = D.up(ordered.data(), ordered.size());
    S.bvh = D.up(bvh.nodes.data(), bvh.nodes.size());
    S.bvh4 = D.up(bvh.nodes4.data(), bvh.nodes4.size());
    S.orig = D.up(bvh.order.data(), bvh.order.size());
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
    DeviceTree DT;
    SceneOnDevice S;
    if (!scene_to_device(D, DT, tris9, ntris, S)) return 0;
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
    DeviceTree DT;
    SceneOnDevice S;
    if (!scene_to_device(D, DT, tris9, ntris, S)) return 0;
    float *df = D.up(from3, (size_t)n * 3), *dt = D.up(to3, (size_t)n * 3), *dk = D.up(k, n), *dout = D.alloc<float>(n);
    u32 *ds = steps_out ? D.alloc<u32>(n) : nullptr;
    if (!D.ok) return 0;
    if (n) t_march_kernel<<<(n + 127) / 128, 128>>>(S.bvh, S.bvh4, S.pt, df, dt, dk, n, dout, ds);
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

/* The device builder against the host builder on the same triangles: both run the same binned SAH, so for non-degenerate
 * input the trees must be EQUAL -- binary nodes (boxes numerically, child codes exactly), 4-wide nodes, and the triangle
 * set of every leaf.  Returns 0 on a device failure; the mismatch counters say how many records differ. */
int ltrx_test_device_bvh(const float *tris9, u32 ntris, int leaf_max, u32 *n_nodes, u32 *n_nodes4, int *height, float *build_ms,
                         u32 *mismatch_nodes, u32 *mismatch_nodes4, u32 *mismatch_leaves, int *host_height)
{
    if (!have_device()) return 0;
    Dev D;
    DeviceTree DT;
    float *d_raw = D.up(tris9, (size_t)ntris * 9);
    if (!D.ok) return 0;
    char err[256] = { 0 };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    { LbDeviceBvh W; if (lb_build_bvh_device(nullptr, d_raw, ntris, leaf_max, 148, &W, err, sizeof(err)) == 0) { lb_free(W.nodes); lb_free(W.nodes4); lb_free(W.order); } }   /* warm the allocator */
    cudaEventRecord(e0, nullptr);
    if (lb_build_bvh_device(nullptr, d_raw, ntris, leaf_max, 148, &DT.T, err, sizeof(err))) { fprintf(stderr, "lighter_b200: %s\n", err); return 0; }
    cudaEventRecord(e1, nullptr);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(build_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *n_nodes = DT.T.n_nodes; *n_nodes4 = DT.T.n_nodes4; *height = DT.T.height;
    std::vector<BvhNode> dn(DT.T.n_nodes);
    std::vector<Bvh4Node> dn4(DT.T.n_nodes4);
    std::vector<uint32_t> dord(ntris);
    D.down(dn.data(), DT.T.nodes, dn.size()); D.down(dn4.data(), DT.T.nodes4, dn4.size()); D.down(dord.data(), DT.T.order, dord.size());
    if (!D.ok) return 0;
    SceneBvh H;
    build_scene_bvh(tris9, ntris, H, leaf_max, 0);
    *host_height = H.depth;
    *mismatch_nodes = *mismatch_nodes4 = *mismatch_leaves = 0;
    if (H.nodes.size() != dn.size()) *mismatch_nodes = (u32)(H.nodes.size() > dn.size() ? H.nodes.size() - dn.size() : dn.size() - H.nodes.size()) + 1u;
    if (H.nodes4.size() != dn4.size()) *mismatch_nodes4 = (u32)(H.nodes4.size() > dn4.size() ? H.nodes4.size() - dn4.size() : dn4.size() - H.nodes4.size()) + 1u;
    std::vector<uint32_t> a, b;
    for (size_t i = 0; i < H.nodes.size() && i < dn.size(); ++i) {
        const BvhNode &x = H.nodes[i], &y = dn[i];
        const bool same = x.lo0x == y.lo0x && x.lo0y == y.lo0y && x.lo0z == y.lo0z && x.hi0x == y.hi0x && x.hi0y == y.hi0y && x.hi0z == y.hi0z &&
                          x.lo1x == y.lo1x && x.lo1y == y.lo1y && x.lo1z == y.lo1z && x.hi1x == y.hi1x && x.hi1y == y.hi1y && x.hi1z == y.hi1z &&
                          x.c0 == y.c0 && x.c1 == y.c1;
        if (!same) { ++*mismatch_nodes; continue; }
        const int32_t cs[2] = { x.c0, x.c1 };
        for (int k = 0; k < 2; ++k) {
            if (cs[k] >= 0) continue;
            const uint32_t code = ~cs[k], first = code >> 3, cnt = code & 7u;
            if (first + cnt > ntris) { ++*mismatch_leaves; continue; }
            a.assign(H.order.begin() + first, H.order.begin() + first + cnt); b.assign(dord.begin() + first, dord.begin() + first + cnt);
            std::sort(a.begin(), a.end());
            if (a != b) ++*mismatch_leaves;                            /* the device lists a leaf's triangles in ascending index */
        }
    }
    for (size_t i = 0; i < H.nodes4.size() && i < dn4.size(); ++i) {
        const Bvh4Node &x = H.nodes4[i], &y = dn4[i];
        bool same = true;
        for (int j = 0; j < 4 && same; ++j)
            same = x.c[j] == y.c[j] && (x.c[j] == BVH4_EMPTY || (x.lox[j] == y.lox[j] && x.loy[j] == y.loy[j] && x.loz[j] == y.loz[j] && x.hix[j] == y.hix[j] && x.hiy[j] == y.hiy[j] && x.hiz[j] == y.hiz[j]));
        if (!same) ++*mismatch_nodes4;
    }
    return 1;
}

} /* extern "C" */
