/*
 * gpu_ao.cu -- ambient occlusion.
 *
 * Reference behaviour restated: lighter.cpp:799-841 (per lumel: ao_num_samples closest-hit segments
 * of length ao_distance on a golden-angle cosine spiral around the normal, rotated by one random
 * offset per lumel; blend into the lumel colour), lighter_int.hpp:456-471 (spiral direction),
 * lighter.cpp:254-289 (closest hit over shadow-casting instances, end points pulled in by 0.001).
 *
 * GPU formulation: one warp per 32 neighbouring lumels traces their samples on the flat BVH, starting
 * at the group's entry set, and stores the hit parameters; a second kernel, one thread per lumel, sums the samples in index order (the
 * reference's float summation order) and applies the blend.  The per-lumel random offset is the
 * host's libc rand() stream replayed in the reference's order (bake.cpp), uploaded per lumel.
 */
#include "gpu_internal.cuh"

/* ref: lighter_int.hpp:456-471; cos_side/sin_side come from the host table (exact libm values) */
__device__ __forceinline__ V3 spiral_dir(V3 dir, float randoff, int i, float cos_side, float sin_side)
{
    const float golden = 137.508f / 180.0f * 3.14159274101257324f;     /* DEG2RAD(137.508f) in float */
    float angle = ((float)i + randoff) * golden;
    float cos_around = ref_cosf(angle), sin_around = ref_sinf(angle);
    V3 diffvec = mk3(dir.y, -dir.z, dir.x);
    V3 up = norm3(cross3(dir, diffvec));
    V3 rt = cross3(dir, up);
    return cos_around * sin_side * rt + sin_around * sin_side * up + cos_side * dir;
}

/*
 * One warp per group of 32 consecutive lumels (neighbouring texels of one lightmap row), one lumel per lane, the samples
 * in a loop: a warp traces the same sample index for 32 neighbouring lumels, i.e. nearly parallel segments from nearby
 * origins.  Every segment of the group lies within ao_distance of the group's origins, so the scene tree is descended
 * ONCE per group (bvh_entry.h; lane 0, entry set in shared memory) and the 32 x num_samples segments start there
 * instead of at the root -- for 2-unit segments in a 400-unit scene most of a root walk is that descent.
 */
template <bool ENTRY>
__global__ void __launch_bounds__(LB_BLOCK)
ao_trace_kernel(const BvhNode *__restrict__ bvh, const RayTri *__restrict__ raytris, const uint32_t *__restrict__ tri_orig,
                const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, const float *__restrict__ randoff,
                const float *__restrict__ cos_side, const float *__restrict__ sin_side,
                uint64_t sh_begin, uint32_t n_local, int num_samples, float ao_distance,
                float *__restrict__ hits, unsigned long long *counters)
{
    __shared__ BvhEntrySet s_entry[LB_BLOCK / 32];
    BvhEntrySet &E = s_entry[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t n_groups = (n_local + 31u) / 32u;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned segs = 0;
    TravStats ts = { 0, 0 };
    for (uint32_t grp = warp0; grp < n_groups; grp += n_warps) {
        const uint32_t li = grp * 32u + lane;
        const bool valid = li < n_local;
        const uint64_t g = sh_begin + (valid ? li : n_local - 1u);
        const V3 SP = ld3(lpos[g]), SN = ld3(lnrm[g]);
        const V3 origin = SP + SN * (LB_SMALL * 2);
        const float ro = valid ? randoff[li] : 0.f;
        if (ENTRY) {
            /* |spiral_dir| <= max(|SN|, 1) up to rounding (lumel normals are unit, probe normals are whatever the caller gave:
             * the reference does not normalise them, lighter_int.hpp:456-471), so every end point is within that many
             * |ao_distance| of its origin */
            const float reach = fabsf(ao_distance) * fmaxf(len3(SN), 1.0f) * 1.001f;
            float lx = origin.x - reach, ly = origin.y - reach, lz = origin.z - reach, hx = origin.x + reach, hy = origin.y + reach, hz = origin.z + reach;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
                hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
            }
            bvh_entry_pad(lx, ly, lz, hx, hy, hz);
            bvh_entry_search_warp<Bvh2Access>(bvh, lx, ly, lz, hx, hy, hz, E, lane);   /* syncs the warp before it overwrites the previous group's set */
        }
        if (!valid) continue;                         /* no warp-wide operation below */
        for (int s = 0; s < num_samples; ++s) {
            const V3 ray = spiral_dir(SN, ro, s, cos_side[s], sin_side[s]) * ao_distance;
            const V3 B = origin + ray;
            const V3 dn = norm3(B - origin);
            const V3 mA = origin + dn * LB_SMALL, mB = B - dn * LB_SMALL;
            const float hit = ENTRY ? bvh_segment_entries<false>(bvh, raytris, tri_orig, E, mA, mB, nullptr, ts)
                                    : bvh_segment<false>(bvh, raytris, tri_orig, mA, mB, nullptr, ts);
            hits[(uint64_t)li * num_samples + s] = hit;
            ++segs;
        }
    }
    count_add(counters, CNT_AO_SEGMENTS, segs);
    count_add(counters, CNT_RAY_NODE_VISITS, ts.nodes);
    count_add(counters, CNT_RAY_TRI_TESTS, ts.tris);
    count_add(counters, CNT_RAY_ENTRY_TESTS, ts.entries);
}

__global__ void ao_apply_kernel(const float *__restrict__ hits, uint64_t sh_begin, uint32_t n_local, ltrgpu_Params P, float4 *__restrict__ lrgb)
{
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_local) return;
    const int n = P.ao_num_samples;
    float ao = 0;
    for (int s = 0; s < n; ++s) {
        float hit = hits[(uint64_t)li * n + s];
        if (hit < 1.0f) ao += 1.0f - hit;
    }
    ao /= n;
    ao = fminr(ao * P.ao_multiplier, 1.0f);
    if (P.ao_falloff) ao = ref_powf(ao, P.ao_falloff);
    const uint64_t g = sh_begin + li;
    float4 c4 = lrgb[g];
    V3 c = mk3(c4.x, c4.y, c4.z);
    const V3 aoc = mk3(P.ao_color[0], P.ao_color[1], P.ao_color[2]);
    if (P.ao_effect >= 0) {
        c = c * (1 - ao * (1 - P.ao_effect)) + aoc * ao;
    } else {
        V3 inner = lerp3(aoc, aoc * c, -P.ao_effect);
        c = lerp3(c, inner, ao);
    }
    lrgb[g] = make_float4(c.x, c.y, c.z, 0.f);
}

extern "C" int ltrgpu_ambient_occlusion(ltrgpu_Ctx *ctx, const float *randoff_host)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n_local64 = ctx->sh_end - ctx->sh_begin;
    const int ns = ctx->params.ao_num_samples;
    if (n_local64 == 0) return 0;
    const uint32_t n_local = (uint32_t)n_local64;
    float *d_rand = nullptr, *d_hits = nullptr;
    if (dev_upload(ctx, &d_rand, randoff_host, n_local)) return 1;       /* this shard's offsets only */
    if (dev_alloc(ctx, &d_hits, (size_t)n_local * (ns > 0 ? ns : 1))) return 1;
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, st));
    if (ns > 0) {
        const uint64_t n_groups = (n_local + 31u) / 32u;                 /* one warp per group of 32 lumels at a time */
        uint64_t want = (n_groups + LB_BLOCK / 32 - 1) / (LB_BLOCK / 32);
        unsigned cap = (unsigned)ctx->num_sms * 64;
        unsigned blocks = want > cap ? cap : (unsigned)want;
        const char *e = getenv("LTR_AO_ENTRY");                           /* A/B switch: 0 = every segment starts at the root */
        if (!e || atoi(e) != 0)
            ao_trace_kernel<true><<<blocks, LB_BLOCK, 0, st>>>(ctx->d_bvh, ctx->d_raytris, ctx->d_tri_orig, ctx->d_lpos, ctx->d_lnrm, d_rand, ctx->d_ao_cos,
                                                               ctx->d_ao_sin, ctx->sh_begin, n_local, ns, ctx->params.ao_distance, d_hits, ctx->d_counters);
        else
            ao_trace_kernel<false><<<blocks, LB_BLOCK, 0, st>>>(ctx->d_bvh, ctx->d_raytris, ctx->d_tri_orig, ctx->d_lpos, ctx->d_lnrm, d_rand, ctx->d_ao_cos,
                                                                ctx->d_ao_sin, ctx->sh_begin, n_local, ns, ctx->params.ao_distance, d_hits, ctx->d_counters);
        CU_LAUNCH_CHECK(ctx);
    }
    ao_apply_kernel<<<grid_for(n_local, 256), 256, 0, st>>>(d_hits, ctx->sh_begin, n_local, ctx->params, ctx->d_lrgb);
    CU_LAUNCH_CHECK(ctx);
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->host_counters.ms_ao += ms;
    lb_free(d_rand); lb_free(d_hits);
    return 0;
}
