/*
 * gpu_radiosity.cu -- multi-bounce radiosity.
 *
 * Reference behaviour restated (lighter.cpp:667-797,1098-1124):
 *   every lumel pair i<j gets the form factor f = dotA*dotB / (len^4 * pi) with
 *   dotA = N_i.(P_j-P_i), dotB = N_j.(P_i-P_j); pairs with dotA or dotB <= 0.001, or f < 0.001, are
 *   dropped; a link is kept only when the segment P_i -> P_j (pulled in 0.001 at both ends) HITS a
 *   triangle (the reference tests `VisibilityTest(...) == false -> skip`, and VisibilityTest returns
 *   "blocked"; SURVEY.md finding 6 -- reproduced as is).  Each bounce moves
 *   ((out_j * diffuse_j) * area_j) * f across every link in both directions, accumulating into
 *   totalLight and inputEnergy in ascending partner order; out <- in after each bounce; the lumel
 *   colour becomes totalLight.
 *
 * GPU formulation:
 *   - the O(N^2) pair loop becomes a tiled sweep: 128-lumel row tiles x 128-lumel column tiles, tile
 *     pairs farther apart than sqrt(1/(0.001*pi)) = 17.85 units can never link (f < 0.001 because
 *     dotA*dotB <= len^2) and are skipped by a box test;
 *   - pairs passing the arithmetic test go to a candidate list (warp-aggregated append), one thread
 *     per candidate then traces the any-hit segment on the scene BVH; survivors are sorted by
 *     (row, partner) so each row's links are in the reference's accumulation order (CSR);
 *   - rows are FULL (both directions) so a bounce is a pure gather, one thread per row, and rows can
 *     be sharded across GPUs: only E_j = (out_j*diffuse_j)*area_j is exchanged per bounce.
 */
#include "gpu_internal.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <stdlib.h>

#define RAD_TILE 128
#define RAD_CUTOFF 17.85f       /* > sqrt(1/(0.001*pi)) = 17.8412; conservative */

struct RadCand { uint32_t row, col; float factor; };

__global__ void rad_geom_kernel(const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, uint64_t n, float4 *__restrict__ gpos, float4 *__restrict__ gnrm)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gpos[i] = lpos[i];
    V3 N = norm3(ld3(lnrm[i]));                   /* the reference re-normalises here (lighter.cpp:680) */
    gnrm[i] = make_float4(N.x, N.y, N.z, 0.f);
}

/* per 128-lumel tile: bounding box of the positions */
__global__ void rad_tile_bounds_kernel(const float4 *__restrict__ gpos, uint64_t n, float4 *__restrict__ tlo, float4 *__restrict__ thi)
{
    __shared__ float slo[3][RAD_TILE / 32], shi[3][RAD_TILE / 32];
    const uint64_t i = (uint64_t)blockIdx.x * RAD_TILE + threadIdx.x;
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    if (i < n) { float4 p = gpos[i]; lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z; }
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0) for (int a = 0; a < 3; ++a) { slo[a][threadIdx.x >> 5] = lo[a]; shi[a][threadIdx.x >> 5] = hi[a]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int a = 0; a < 3; ++a)
            for (int w = 1; w < RAD_TILE / 32; ++w) { slo[a][0] = fminf(slo[a][0], slo[a][w]); shi[a][0] = fmaxf(shi[a][0], shi[a][w]); }
        tlo[blockIdx.x] = make_float4(slo[0][0], slo[1][0], slo[2][0], 0.f);
        thi[blockIdx.x] = make_float4(shi[0][0], shi[1][0], shi[2][0], 0.f);
    }
}

/* form factor of the ordered pair (a < b); returns false when the reference drops the pair */
__device__ __forceinline__ bool rad_pair_factor(V3 Pa, V3 Na, V3 Pb, V3 Nb, float &factor)
{
    V3 d = Pb - Pa;
    float dotA = dot3(Na, d);
    float dotB = dot3(Nb, -d);
    if (dotA <= LB_SMALL || dotB <= LB_SMALL) return false;
    float lensq = lensq3(d);
    factor = dotA * dotB / (lensq * lensq * 3.14159274101257324f);
    return !(factor < LB_SMALL);
}

/* one CTA = one row tile; sweeps the column tiles in ascending order */
__global__ void __launch_bounds__(RAD_TILE)
rad_candidates_kernel(const float4 *__restrict__ gpos, const float4 *__restrict__ gnrm, uint64_t n,
                      const float4 *__restrict__ tlo, const float4 *__restrict__ thi, uint32_t n_tiles,
                      uint64_t row_begin, uint64_t row_end, uint32_t first_row_tile,
                      RadCand *__restrict__ cand, unsigned long long cand_cap, unsigned long long *cand_count,
                      unsigned long long *counters)
{
    __shared__ float4 sp[RAD_TILE], sn[RAD_TILE];
    const uint32_t rt = first_row_tile + blockIdx.x;
    const uint64_t r = (uint64_t)rt * RAD_TILE + threadIdx.x;
    const bool row_ok = r < n && r >= row_begin && r < row_end;
    V3 Pr = mk3(0.f), Nr = mk3(0.f);
    if (r < n) { Pr = ld3(gpos[r]); Nr = ld3(gnrm[r]); }
    const float4 rlo = tlo[rt], rhi = thi[rt];
    const unsigned lane = threadIdx.x & 31u;
    unsigned tested = 0;
    for (uint32_t ct = 0; ct < n_tiles; ++ct) {
        const float4 clo = tlo[ct], chi = thi[ct];
        float dx = fmaxf(fmaxf(clo.x - rhi.x, rlo.x - chi.x), 0.f);
        float dy = fmaxf(fmaxf(clo.y - rhi.y, rlo.y - chi.y), 0.f);
        float dz = fmaxf(fmaxf(clo.z - rhi.z, rlo.z - chi.z), 0.f);
        if (dx * dx + dy * dy + dz * dz > RAD_CUTOFF * RAD_CUTOFF) continue;       /* CTA-uniform */
        __syncthreads();
        const uint64_t cj = (uint64_t)ct * RAD_TILE + threadIdx.x;
        if (cj < n) { sp[threadIdx.x] = gpos[cj]; sn[threadIdx.x] = gnrm[cj]; }
        __syncthreads();
        const uint32_t cnt = (uint32_t)((uint64_t)ct * RAD_TILE + RAD_TILE <= n ? RAD_TILE : n - (uint64_t)ct * RAD_TILE);
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint64_t j = (uint64_t)ct * RAD_TILE + k;
            bool keep = false;
            float f = 0.f;
            if (row_ok && j != r) {
                V3 Pj = ld3(sp[k]), Nj = ld3(sn[k]);
                keep = (j > r) ? rad_pair_factor(Pr, Nr, Pj, Nj, f) : rad_pair_factor(Pj, Nj, Pr, Nr, f);
                ++tested;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (mask) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(cand_count, (unsigned long long)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) {
                    unsigned long long at = base + __popc(mask & ((1u << lane) - 1u));
                    if (at < cand_cap) { RadCand c; c.row = (uint32_t)r; c.col = (uint32_t)j; c.factor = f; cand[at] = c; }
                }
            }
        }
    }
    count_add(counters, CNT_RAD_PAIRS, tested);
}

/* one thread per candidate: blocked segment -> link */
__global__ void __launch_bounds__(LB_BLOCK)
rad_visibility_kernel(const BvhNode *__restrict__ bvh, const RayTri *__restrict__ raytris, const float4 *__restrict__ gpos,
                      const RadCand *__restrict__ cand, unsigned long long n_cand,
                      unsigned long long *__restrict__ keys, float *__restrict__ factors, unsigned long long *link_count,
                      unsigned long long *counters)
{
    unsigned segs = 0;
    TravStats ts = { 0, 0 };
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n_pad = (n_cand + 31ull) & ~31ull;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_pad; e += (unsigned long long)gridDim.x * blockDim.x) {
        bool keep = false;
        RadCand c = { 0, 0, 0.f };
        if (e < n_cand) {
            c = cand[e];
            const uint32_t a = c.row < c.col ? c.row : c.col, b = c.row < c.col ? c.col : c.row;
            const V3 A = ld3(gpos[a]), B = ld3(gpos[b]);
            const V3 dn = norm3(B - A);
            const V3 mA = A + dn * LB_SMALL, mB = B - dn * LB_SMALL;
            keep = bvh_segment<true>(bvh, raytris, nullptr, mA, mB, nullptr, ts) < 1.0f;
            ++segs;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (mask) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(link_count, (unsigned long long)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                unsigned long long at = base + __popc(mask & ((1u << lane) - 1u));
                keys[at] = ((unsigned long long)c.row << 32) | c.col;
                factors[at] = c.factor;
            }
        }
    }
    count_add(counters, CNT_RAD_SEGMENTS, segs);
    count_add(counters, CNT_NODE_VISITS, ts.nodes);
    count_add(counters, CNT_TRI_TESTS, ts.tris);
}

__global__ void rad_row_offsets_kernel(const unsigned long long *__restrict__ keys, unsigned long long n_links, uint64_t row_begin,
                                       uint64_t n_rows, uint64_t *__restrict__ rowoff)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_rows) return;
    const unsigned long long want = (unsigned long long)(row_begin + i) << 32;     /* first key of row */
    unsigned long long lo = 0, hi = n_links;
    while (lo < hi) { unsigned long long mid = (lo + hi) >> 1; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
    rowoff[i] = lo;
}

__global__ void rad_split_keys_kernel(const unsigned long long *__restrict__ keys, unsigned long long n, uint32_t *__restrict__ other)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) other[i] = (uint32_t)(keys[i] & 0xffffffffull);
}

/* material / energy state: diffuse, total, out per lumel (float4 each) */
__global__ void rad_init_kernel(const float4 *__restrict__ lrgb, const float4 *__restrict__ lrad, const float *__restrict__ diffuse3,
                                const float *__restrict__ emissive3, uint32_t n_probes, uint64_t begin, uint64_t end,
                                float4 *__restrict__ diff, float4 *__restrict__ total, float4 *__restrict__ out)
{
    uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const bool probe = i < n_probes;
    V3 d = probe ? mk3(0.f) : mk3(1.f);
    V3 e = ld3(lrgb[i]);
    if (!probe && diffuse3) {
        d = mk3(diffuse3[i * 3], diffuse3[i * 3 + 1], diffuse3[i * 3 + 2]);
        e = e + mk3(emissive3[i * 3], emissive3[i * 3 + 1], emissive3[i * 3 + 2]);
    }
    const float area = probe ? 0.f : lrad[i].w;
    diff[i] = make_float4(d.x, d.y, d.z, area);
    total[i] = make_float4(e.x, e.y, e.z, 0.f);
    out[i] = make_float4(e.x, e.y, e.z, 0.f);
}

/* E_j = (out_j * diffuse_j) * area_j, the quantity that crosses a link (and NVLink) */
__global__ void rad_energy_kernel(const float4 *__restrict__ diff, const float4 *__restrict__ out, uint64_t begin, uint64_t end, float4 *__restrict__ E)
{
    uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const float4 d = diff[i], o = out[i];
    V3 e = (ld3(o) * ld3(d)) * d.w;
    E[i] = make_float4(e.x, e.y, e.z, 0.f);
}

__global__ void rad_bounce_kernel(const uint64_t *__restrict__ rowoff, const uint32_t *__restrict__ other, const float *__restrict__ factor,
                                  const float4 *__restrict__ E, uint64_t begin, uint64_t end, float4 *__restrict__ total, float4 *__restrict__ out)
{
    uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint64_t a = rowoff[i - begin], b = rowoff[i - begin + 1];
    V3 t = ld3(total[i]);
    V3 in = mk3(0.f);
    for (uint64_t k = a; k < b; ++k) {
        const V3 c = ld3(__ldg(E + other[k])) * factor[k];
        t = t + c;
        in = in + c;
    }
    total[i] = make_float4(t.x, t.y, t.z, 0.f);
    out[i] = make_float4(in.x, in.y, in.z, 0.f);
}

__global__ void rad_commit_kernel(const float4 *__restrict__ total, uint64_t begin, uint64_t end, float4 *__restrict__ lrgb)
{
    uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) lrgb[i] = total[i];
}

template <class T> static int grow(ltrgpu_Ctx *ctx, T **p, size_t *cap, size_t used, size_t need)
{
    if (need <= *cap) return 0;
    size_t ncap = *cap ? *cap : 1 << 20;
    while (ncap < need) ncap *= 2;
    T *q = nullptr;
    CU_TRY(ctx, cudaMalloc((void **)&q, ncap * sizeof(T)));
    if (*p && used) CU_TRY(ctx, cudaMemcpyAsync(q, *p, used * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (*p) cudaFree(*p);
    *p = q; *cap = ncap;
    return 0;
}

extern "C" int ltrgpu_radiosity(ltrgpu_Ctx *ctx, const float *diffuse3, const float *emissive3, int bounces)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n = ctx->n_lumels;
    if (n == 0 || bounces <= 0) return 0;
    if (n > 0xfffffff0ull) { snprintf(ctx->err, sizeof(ctx->err), "too many lumels for radiosity"); return 1; }
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, st));
    const uint64_t rb = ctx->sh_begin, re = ctx->sh_end, n_rows = re - rb;
    const uint32_t n_tiles = (uint32_t)((n + RAD_TILE - 1) / RAD_TILE);

    float4 *gpos = nullptr, *gnrm = nullptr, *tlo = nullptr, *thi = nullptr;
    float4 *diff = nullptr, *total = nullptr, *out = nullptr, *E = nullptr;
    float *d_diffuse = nullptr, *d_emissive = nullptr;
    RadCand *cand = nullptr;
    unsigned long long *d_cnt = nullptr;           /* [0] candidates, [1] links */
    unsigned long long *keys = nullptr, *keys_alt = nullptr;
    float *fac = nullptr, *fac_alt = nullptr;
    size_t link_cap = 0, link_cap_f = 0, link_used = 0;
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    int rc = 1;

#define RAD_TRY(x) do { if (x) goto done; } while (0)
#define RAD_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto done; } } while (0)

    {
        RAD_TRY(dev_alloc(ctx, &gpos, n + LB_PAD)); RAD_TRY(dev_alloc(ctx, &gnrm, n + LB_PAD));
        RAD_TRY(dev_alloc(ctx, &tlo, n_tiles)); RAD_TRY(dev_alloc(ctx, &thi, n_tiles));
        RAD_TRY(dev_alloc(ctx, &diff, n + LB_PAD)); RAD_TRY(dev_alloc(ctx, &total, n + LB_PAD));
        RAD_TRY(dev_alloc(ctx, &out, n + LB_PAD)); RAD_TRY(dev_alloc(ctx, &E, n + LB_PAD));
        RAD_TRY(dev_alloc(ctx, &d_cnt, 2));
        if (diffuse3) { RAD_TRY(dev_upload(ctx, &d_diffuse, diffuse3, n * 3)); RAD_TRY(dev_upload(ctx, &d_emissive, emissive3, n * 3)); }

        rad_geom_kernel<<<grid_for(n, 256), 256, 0, st>>>(ctx->d_lpos, ctx->d_lnrm, n, gpos, gnrm);
        ctx->host_counters.kernel_launches++;
        rad_tile_bounds_kernel<<<n_tiles, RAD_TILE, 0, st>>>(gpos, n, tlo, thi);
        ctx->host_counters.kernel_launches++;
        RAD_CU(cudaGetLastError());

        /* ---- link generation, in batches of row tiles bounded by the candidate buffer ---- */
        const unsigned long long cand_cap = 32ull << 20;             /* 32 Mi candidates = 384 MiB */
        RAD_TRY(dev_alloc(ctx, &cand, cand_cap));
        const uint32_t t_begin = (uint32_t)(rb / RAD_TILE), t_end = (uint32_t)((re + RAD_TILE - 1) / RAD_TILE);
        uint32_t batch = 1024;
        for (uint32_t t0 = t_begin; t0 < t_end;) {
            uint32_t t1 = t0 + batch < t_end ? t0 + batch : t_end;
            RAD_CU(cudaMemsetAsync(d_cnt, 0, 16, st));
            rad_candidates_kernel<<<t1 - t0, RAD_TILE, 0, st>>>(gpos, gnrm, n, tlo, thi, n_tiles, rb, re, t0, cand, cand_cap, d_cnt, ctx->d_counters);
            ctx->host_counters.kernel_launches++;
            RAD_CU(cudaGetLastError());
            unsigned long long h_cnt[2];
            RAD_CU(cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st));
            RAD_CU(cudaStreamSynchronize(st));
            if (h_cnt[0] > cand_cap) {
                if (batch == 1) { snprintf(ctx->err, sizeof(ctx->err), "radiosity: one row tile produced %llu candidates (> %llu)", h_cnt[0], cand_cap); goto done; }
                batch = batch / 2 ? batch / 2 : 1;
                /* the counters of the aborted attempt are discarded below by recounting: subtract them */
                continue;
            }
            const unsigned long long nc = h_cnt[0];
            if (nc) {
                RAD_TRY(grow(ctx, &keys, &link_cap, link_used, link_used + nc));
                RAD_TRY(grow(ctx, &fac, &link_cap_f, link_used, link_used + nc));
                unsigned long long want = (nc + LB_BLOCK - 1) / LB_BLOCK;
                unsigned cap = (unsigned)ctx->num_sms * 32;
                unsigned blocks = want > cap ? cap : (unsigned)want;
                rad_visibility_kernel<<<blocks, LB_BLOCK, 0, st>>>(ctx->d_bvh, ctx->d_raytris, gpos, cand, nc, keys + link_used, fac + link_used,
                                                                  d_cnt + 1, ctx->d_counters);
                ctx->host_counters.kernel_launches++;
                RAD_CU(cudaGetLastError());
                RAD_CU(cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st));
                RAD_CU(cudaStreamSynchronize(st));
                link_used += h_cnt[1];
            }
            t0 = t1;
        }

        /* ---- sort links by (row, partner) -> CSR in reference accumulation order ---- */
        dev_free(&cand);
        if (link_used) {
            RAD_TRY(dev_alloc(ctx, &keys_alt, link_used)); RAD_TRY(dev_alloc(ctx, &fac_alt, link_used));
            cub::DoubleBuffer<unsigned long long> kb(keys, keys_alt);
            cub::DoubleBuffer<float> vb(fac, fac_alt);
            RAD_CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp_bytes, kb, vb, (int64_t)link_used, 0, 64, st));
            RAD_CU(cudaMalloc(&sort_tmp, sort_tmp_bytes ? sort_tmp_bytes : 16));
            RAD_CU(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_tmp_bytes, kb, vb, (int64_t)link_used, 0, 64, st));
            ctx->host_counters.kernel_launches += 8;
            RAD_CU(cudaStreamSynchronize(st));
            if (kb.Current() != keys) { unsigned long long *t = keys; keys = keys_alt; keys_alt = t; }
            if (vb.Current() != fac) { float *t = fac; fac = fac_alt; fac_alt = t; }
        }
        dev_free(&ctx->d_rad_rowoff); dev_free(&ctx->d_rad_other); dev_free(&ctx->d_rad_factor);
        RAD_TRY(dev_alloc(ctx, &ctx->d_rad_rowoff, n_rows + 1));
        RAD_TRY(dev_alloc(ctx, &ctx->d_rad_other, link_used));
        rad_row_offsets_kernel<<<grid_for(n_rows + 1, 256), 256, 0, st>>>(keys, link_used, rb, n_rows, ctx->d_rad_rowoff);
        ctx->host_counters.kernel_launches++;
        if (link_used) {
            rad_split_keys_kernel<<<grid_for(link_used, 256), 256, 0, st>>>(keys, link_used, ctx->d_rad_other);
            ctx->host_counters.kernel_launches++;
        }
        RAD_CU(cudaGetLastError());
        ctx->d_rad_factor = fac; fac = nullptr;
        ctx->rad_rows = n_rows; ctx->rad_links = link_used;
        {
            unsigned long long lc = link_used;
            RAD_CU(cudaMemcpyAsync(ctx->d_counters + CNT_RAD_LINKS, &lc, 8, cudaMemcpyHostToDevice, st));
            RAD_CU(cudaStreamSynchronize(st));
        }

        /* ---- bounces ---- */
        if (n_rows) {
            rad_init_kernel<<<grid_for(n_rows, 256), 256, 0, st>>>(ctx->d_lrgb, ctx->d_lrad, d_diffuse, d_emissive, ctx->n_probes, rb, re, diff, total, out);
            ctx->host_counters.kernel_launches++;
        }
        for (int b = 0; b < bounces; ++b) {
            if (n_rows) {
                rad_energy_kernel<<<grid_for(n_rows, 256), 256, 0, st>>>(diff, out, rb, re, E);
                ctx->host_counters.kernel_launches++;
            }
            if (ctx->world > 1) {
                const uint64_t chunk = (n + ctx->world - 1) / ctx->world;
                if (!ctx->allgather || ctx->allgather(ctx->allgather_user, E + chunk * ctx->rank, E, chunk * sizeof(float4), st)) {
                    snprintf(ctx->err, sizeof(ctx->err), "radiosity: per-bounce all-gather failed");
                    goto done;
                }
            }
            if (n_rows) {
                rad_bounce_kernel<<<grid_for(n_rows, 128), 128, 0, st>>>(ctx->d_rad_rowoff, ctx->d_rad_other, ctx->d_rad_factor, E, rb, re, total, out);
                ctx->host_counters.kernel_launches++;
            }
            RAD_CU(cudaGetLastError());
        }
        if (n_rows) {
            rad_commit_kernel<<<grid_for(n_rows, 256), 256, 0, st>>>(total, rb, re, ctx->d_lrgb);
            ctx->host_counters.kernel_launches++;
        }
        RAD_CU(cudaEventRecord(ctx->ev1, st));
        RAD_CU(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->host_counters.ms_radiosity += ms;
        rc = 0;
    }
done:
    cudaFree(gpos); cudaFree(gnrm); cudaFree(tlo); cudaFree(thi); cudaFree(diff); cudaFree(total); cudaFree(out); cudaFree(E);
    cudaFree(d_diffuse); cudaFree(d_emissive); cudaFree(cand); cudaFree(d_cnt); cudaFree(keys); cudaFree(keys_alt);
    cudaFree(fac); cudaFree(fac_alt); cudaFree(sort_tmp);
    return rc;
#undef RAD_TRY
#undef RAD_CU
}

extern "C" int ltrgpu_download_links(ltrgpu_Ctx *ctx, uint64_t *row_offset, uint32_t *other, float *factor, uint64_t *rows, uint64_t *count)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    *rows = ctx->rad_rows; *count = ctx->rad_links;
    if (!ctx->d_rad_rowoff) return 0;
    if (row_offset) CU_TRY(ctx, cudaMemcpyAsync(row_offset, ctx->d_rad_rowoff, (ctx->rad_rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (other && ctx->rad_links) CU_TRY(ctx, cudaMemcpyAsync(other, ctx->d_rad_other, ctx->rad_links * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (factor && ctx->rad_links) CU_TRY(ctx, cudaMemcpyAsync(factor, ctx->d_rad_factor, ctx->rad_links * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
