/*
 * gpu_direct.cu -- the direct-light pass, the dominant cost of a bake.
 *
 * Reference behaviour restated (file:line into /root/reference):
 *   shading   lighter.cpp:485-600   point / spot / directional terms, lights accumulate in add order
 *   march     lighter.cpp:190-207   CalcInvShadowFactor: sphere-traced penumbra estimate
 *   distance  lighter.cpp:150-188   min(2, nearest triangle distance) over shadow-casting instances
 *
 * GPU formulation (three kernels):
 *   1. direct_classify  one thread per (light, lumel): evaluate the cheap terms; pairs whose product
 *                       is > 0 are appended (warp-aggregated) to a work list, so that the expensive
 *                       kernel only sees real marches and neighbouring list entries are neighbouring
 *                       lumels of the same light (coherent BVH walks);
 *   2. direct_march     one thread per work-list entry: the march itself, each step a nearest-distance
 *                       query on the flat BVH; writes the shadow factor f_vis[light][lumel];
 *   3. direct_accumulate one thread per lumel: re-evaluates the cheap terms (bit-identical, same code)
 *                       and adds the lights in their original order -- float summation order is the
 *                       reference's.
 */
#include "gpu_internal.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>

#include <vector>

struct ShadeTerms {
    float pre;         /* product of the non-shadow factors the reference tests against <= 0 */
    float f_dist, f_ndotl, f_dir;
    V3 to;             /* march target */
    V3 s2l;            /* unit vector sample -> light (contribution direction for the normal map) */
};

/* ref: lighter.cpp:493-502 (point), :528-540 (spot), :563-587 (directional) */
__device__ __forceinline__ ShadeTerms shade_terms(const ltrgpu_Light &L, V3 SP, V3 SN)
{
    ShadeTerms o;
    if (L.type == 3u) {
        o.f_dist = 1.f; o.f_dir = 1.f;
        o.f_ndotl = fmaxr(0.0f, dot3(L.dir, SN));
        o.pre = 1.f;                                   /* the reference never early-outs a directional light */
        o.to = SP + L.dir * L.range;
        o.s2l = L.dir;
        return o;
    }
    V3 s2l = L.pos - SP;
    float dist = len3(s2l);
    if (dist) s2l = s2l / dist;
    o.f_dist = ref_powf(1 - fminr(1.0f, dist / L.range), L.power);
    o.f_ndotl = fmaxr(0.0f, dot3(s2l, SN));
    o.f_dir = 1.f;
    o.pre = o.f_dist * o.f_ndotl;
    if (L.type == 2u) {
        float angle = ref_acosf(fminr(1.0f, dot3(s2l, -L.dir)));
        float f = fmaxr(0.0f, fminr(1.0f, (angle - L.angle_out_rad) / L.angle_diff));
        o.f_dir = ref_powf(f, L.curve);
        o.pre = o.f_dist * o.f_ndotl * o.f_dir;
    }
    o.to = L.pos;
    o.s2l = s2l;
    return o;
}

/* Work-list order of the march (LTR_MARCH_ORDER, default on).  Lumels are stored texel row by texel row, so 32 consecutive
 * ones are a LINE (3 units long on config 4, eight BVH leaves wide); the 32 marches of a warp walk the tree side by side,
 * and what they share of it decides how many lanes an instruction serves.  The listing pass therefore visits the rank's
 * lumels along a Morton curve over 0.25-unit cells (coordinates wrap every 256 units: no scene bounds needed, a wrap only
 * lets two far cells share a key), so a warp's lumels form a compact patch.  Only the ORDER of the work list changes;
 * every (lumel, light) march and its slot in the factor table are the same. */
__global__ void march_order_kernel(const float4 *__restrict__ lpos, uint64_t first, uint32_t n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = lpos[first + i];
    auto cell = [](float v) { return (uint32_t)((int)floorf(v * 4.0f)) & 1023u; };
    auto spread = [](uint32_t x) { x &= 0x3ffu; x = (x | (x << 16)) & 0x30000ffu; x = (x | (x << 8)) & 0x300f00fu; x = (x | (x << 4)) & 0x30c30c3u; x = (x | (x << 2)) & 0x9249249u; return x; };
    keys[i] = spread(cell(p.x)) | (spread(cell(p.y)) << 1) | (spread(cell(p.z)) << 2);
    vals[i] = i;
}

/* cut the blocks into `world` contiguous runs of (nearly) equal weight; cuts[r] = first block of rank r, cuts[world] = nb */
__global__ void direct_cuts_kernel(const uint32_t *__restrict__ block_w, uint32_t nb, uint32_t world, uint32_t *__restrict__ cuts)
{
    if (blockIdx.x || threadIdx.x) return;
    unsigned long long total = 0;
    for (uint32_t b = 0; b < nb; ++b) total += block_w[b];
    unsigned long long acc = 0;
    uint32_t r = 1;
    cuts[0] = 0;
    for (uint32_t b = 0; b < nb && r < world; ++b) {
        acc += block_w[b];
        while (r < world && acc * world >= total * r) cuts[r++] = b + 1;
    }
    while (r <= world) cuts[r++] = nb;
    cuts[world] = nb;
}

__device__ __forceinline__ bool light_is_supported(unsigned type) { return type == 1u || type == 2u || type == 3u; }

__global__ void direct_classify_kernel(const ltrgpu_Light *__restrict__ lights, uint32_t l0, uint32_t l1,
                                       const uint8_t *__restrict__ light_inst, uint32_t n_inst,
                                       const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, const uint32_t *__restrict__ linst,
                                       uint64_t sh_begin, uint32_t n_local,
                                       uint32_t *__restrict__ block_w /* weight pass: per 1024-lumel block, or NULL */,
                                       uint2 *__restrict__ active, uint32_t *active_count,
                                       const uint32_t *__restrict__ order = nullptr /* listing pass: thread k looks at local lumel order[k] (march_order_kernel) */)
{
    /* (sh_begin, n_local) is the lumel range this launch looks at.  With several GPUs a first pass over ALL lumels and ALL
     * lights (block_w != NULL) sums an integer cost estimate of the marches per 1024-lumel block -- identical on every rank --
     * the blocks are cut into `world` contiguous runs of equal cost (direct_cuts_kernel), and each rank lists, marches and
     * shades the pairs of its own run.  Contiguous = each rank's marches stay in one part of the BVH (an interleaved deal
     * was measured 30 % slower per march: every rank then streams the whole 200 MB scene through its L2); equal cost,
     * because march work is concentrated around the lights (equal lumel counts: 17 ms on the slowest of 8 ranks against a
     * 6 ms mean). */
    const uint32_t k_thread = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t li = (order && k_thread < n_local) ? order[k_thread] : k_thread;
    const uint32_t l = l0 + blockIdx.y;
    bool want = false;
    const uint32_t blk = (uint32_t)((sh_begin + li) >> 10);
    if (k_thread < n_local && l < l1) {
        const ltrgpu_Light L = lights[l];
        const uint64_t g = sh_begin + li;
        if (light_is_supported(L.type) && light_inst[(size_t)l * n_inst + linst[g]]) {
            ShadeTerms t = shade_terms(L, ld3(lpos[g]), ld3(lnrm[g]));
            /* a directional light facing away adds colour * 0: the march result cannot matter, skip it */
            want = (L.type == 3u) ? (t.f_ndotl > 0) : (t.pre > 0);
            if (want && block_w) {
                /* cost ~ distance queries of the march ~ its length (steps are <= 1 and ~1 in open space); integer, so the sum
                 * does not depend on the order of the atomics and every rank computes the same cuts */
                const float len = len3(t.to - (ld3(lpos[g]) + ld3(lnrm[g]) * 0.005f));
                atomicAdd(block_w + blk, 4u + (uint32_t)fminf(len * 4.0f, 4000.0f));
            }
        }
    }
    if (block_w) return;
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask) {
        const unsigned lane = threadIdx.x & 31u;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(active_count, (unsigned)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (want) {
            active[base + __popc(mask & ((1u << lane) - 1u))] = make_uint2(li, l);
        }
    }
}

#ifndef LB_MARCH_MINBLOCKS
#define LB_MARCH_MINBLOCKS 8      /* 64 registers: measured on B200 (config 4): 41.3 ms vs 42.3 uncapped (72 regs), 52.8 at 10 blocks (spills) */
#endif
__global__ void __launch_bounds__(LB_BLOCK, LB_MARCH_MINBLOCKS)
direct_march_kernel(const ltrgpu_Light *__restrict__ lights, const BvhNode *__restrict__ bvh, const Bvh4Node *__restrict__ bvh4, const PreparedTri *__restrict__ tris,
                    const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, uint64_t sh_begin, uint32_t n_local,
                    const uint2 *__restrict__ active, const uint32_t *__restrict__ active_count, uint32_t *cursor, uint32_t l0,
                    float *__restrict__ fvis, unsigned long long *counters)
{
    /* One march per thread per trip; a warp holds 32 consecutive work-list entries = neighbouring
     * lumels marching towards the same light, whose marches have similar lengths and walk the same
     * part of the BVH.  (A persistent-lane variant that refilled finished lanes from a global cursor
     * was measured 40-60 % SLOWER on configs 3 and 4: it trades this coherence for lane occupancy,
     * and the divergence that matters is inside the BVH walk, not in the march length.)
     * Warps pull 32 entries at a time from `cursor`: with 1024-entry CTA chunks a rank of an 8-GPU bake
     * (1.1 M pairs) had one chunk per CTA and ran at a third of the single-GPU throughput. */
    const uint32_t n = *active_count;
    const unsigned lane = threadIdx.x & 31u;
    unsigned queries = 0, marches = 0;
    TravStats ts = { 0, 0 };
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t e = base + lane;
        if (e < n) {
            const uint2 a = active[e];
            const ltrgpu_Light L = lights[a.y];
            const uint64_t g = sh_begin + a.x;
            const V3 SP = ld3(lpos[g]), SN = ld3(lnrm[g]);
            const V3 to = (L.type == 3u) ? SP + L.dir * L.range : L.pos;
            float f = march_shadow(bvh, bvh4, tris, SP + SN * 0.005f, to, L.radius, queries, ts);
            fvis[(size_t)(a.y - l0) * n_local + a.x] = f;
            ++marches;
        }
    }
    count_add(counters, CNT_MARCHES, marches);
    count_add(counters, CNT_DIST_QUERIES, queries);
    count_add(counters, CNT_NODE_VISITS, ts.nodes);
    count_add(counters, CNT_TRI_TESTS, ts.tris);
}

/*
 * Sampled soft shadows (extension mode, see geom.h shadow_sample_segment): one thread per
 * (work-list entry, sample) = lumel x light x soft-shadow sample.  Consecutive lanes are the samples of
 * one lumel/light pair (same origin, targets on one small disk) and neighbouring pairs are neighbouring
 * lumels, so a warp's rays walk the same BVH nodes; node and triangle records are fetched as float4
 * through the read-only path (bvh_anyhit).  A blocked sample sets its bit in the pair's 64-bit mask:
 * bits are OR-reduced over the lanes of the pair first (match_any), one atomic per pair and warp.
 */
__global__ void __launch_bounds__(LB_BLOCK)
direct_sampled_kernel(const ltrgpu_Light *__restrict__ lights, const float4 *__restrict__ samples, const Bvh4Node *__restrict__ bvh,
                      const RayTri *__restrict__ raytris, const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm,
                      uint64_t sh_begin, uint32_t n_local, const uint2 *__restrict__ active, const uint32_t *__restrict__ active_count,
                      uint32_t spp /* samples per pair slot = max over the lights of this chunk */, uint32_t l0,
                      unsigned long long *__restrict__ smask, unsigned long long *counters)
{
    const bool spp_pow2 = (spp & (spp - 1u)) == 0u;
    const unsigned long long total = (unsigned long long)(*active_count) * spp;
    const unsigned long long total_pad = (total + 31ull) & ~31ull;
    unsigned rays = 0;
    TravStats ts = { 0, 0 };
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total_pad; t += (unsigned long long)gridDim.x * blockDim.x) {
        bool blocked = false, live = false;
        uint32_t e = 0xffffffffu, s = 0;
        uint2 a = make_uint2(0, 0);
        if (t < total) {
            e = (uint32_t)(t / spp); s = (uint32_t)(t % spp);
            a = active[e];
            const ltrgpu_Light L = lights[a.y];
            if (s < L.n_samples) {
                live = true;
                const uint64_t g = sh_begin + a.x;
                const float4 sm = samples[L.sample_off + s];
                V3 from, to;
                shadow_sample_segment(L.type, L.pos, L.range, mk3(sm.x, sm.y, sm.z), ld3(lpos[g]), ld3(lnrm[g]), from, to);
                const V3 dn = norm3(to - from);                       /* VisibilityTest: both ends pulled in (lighter.cpp:138-147) */
                blocked = bvh4_anyhit<2>(bvh, raytris, from + dn * LB_SMALL, to - dn * LB_SMALL, ts);   /* shadow rays are often blocked: test early */
                ++rays;
            }
        }
        unsigned lo = (blocked && s < 32u) ? 1u << s : 0u, hi = (blocked && s >= 32u) ? 1u << (s - 32u) : 0u;
        bool leader;
        if (spp_pow2) {
            /* the lanes of a pair are an aligned group of min(spp,32) lanes (t is lane-aligned): butterfly OR */
            const unsigned gw = spp < 32u ? spp : 32u;
            for (unsigned o = 1; o < gw; o <<= 1) { lo |= __shfl_xor_sync(0xffffffffu, lo, o); hi |= __shfl_xor_sync(0xffffffffu, hi, o); }
            leader = ((threadIdx.x & 31u) & (gw - 1u)) == 0u;
        } else {
            const unsigned peers = __match_any_sync(0xffffffffu, e);
            lo = __reduce_or_sync(peers, lo); hi = __reduce_or_sync(peers, hi);
            leader = (threadIdx.x & 31u) == (unsigned)__ffs(peers) - 1u;
        }
        if (live && (lo | hi) && leader) {
            unsigned long long *dst = smask + (size_t)(a.y - l0) * n_local + a.x;
            const unsigned long long bits = ((unsigned long long)hi << 32) | lo;
            if (spp <= 32u) *dst = bits;                                  /* one warp owns the whole pair */
            else atomicOr(dst, bits);
        }
    }
    count_add(counters, CNT_SHADOW_RAYS, rays);
    count_add(counters, CNT_RAY_NODE_VISITS, ts.nodes);
    count_add(counters, CNT_RAY_TRI_TESTS, ts.tris);
}

/* blocked-sample masks -> shadow factor of every work-list pair: f_vis = 1 - blocked / n */
__global__ void sampled_resolve_kernel(const ltrgpu_Light *__restrict__ lights, const uint2 *__restrict__ active, const uint32_t *__restrict__ active_count,
                                       uint32_t n_local, uint32_t l0, const unsigned long long *__restrict__ smask, float *__restrict__ fvis)
{
    const uint32_t n_active = *active_count;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_active; e += gridDim.x * blockDim.x) {
        const uint2 a = active[e];
        const size_t at = (size_t)(a.y - l0) * n_local + a.x;
        const float n = (float)lights[a.y].n_samples;
        fvis[at] = 1.0f - (float)__popcll(smask[at]) / n;
    }
}

__global__ void direct_accumulate_kernel(const ltrgpu_Light *__restrict__ lights, uint32_t l0, uint32_t l1,
                                         const uint8_t *__restrict__ light_inst, uint32_t n_inst,
                                         const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, const uint32_t *__restrict__ linst,
                                         uint64_t sh_begin, uint32_t n_local, const float *__restrict__ fvis, uint64_t tab_base, uint32_t tab_n,
                                         float4 *__restrict__ lrgb)
{
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_local) return;
    const uint64_t g = sh_begin + li;
    const uint32_t ti = (uint32_t)(g - tab_base);                  /* column of this lumel in the factor table */
    const V3 SP = ld3(lpos[g]), SN = ld3(lnrm[g]);
    const uint32_t inst = linst[g];
    float4 c4 = lrgb[g];
    V3 c = mk3(c4.x, c4.y, c4.z);
    for (uint32_t l = l0; l < l1; ++l) {
        const ltrgpu_Light L = lights[l];
        if (!light_is_supported(L.type) || !light_inst[(size_t)l * n_inst + inst]) continue;
        ShadeTerms t = shade_terms(L, SP, SN);
        float fv = fvis[(size_t)(l - l0) * tab_n + ti];
        float f;
        if (L.type == 3u) f = t.f_ndotl * fv;
        else {
            if (!(t.pre > 0)) continue;
            f = (L.type == 2u) ? t.f_dist * t.f_ndotl * t.f_dir * fv : t.f_dist * t.f_ndotl * fv;
        }
        c = c + L.color * f;
    }
    lrgb[g] = make_float4(c.x, c.y, c.z, 0.f);
}

/*
 * Normal / focus map terms per lumel (ref: contributions lighter.cpp:505-513,543-551,590-598,
 * commit :656-664, reduction :977-1012).  Per lumel: average of N*ambient_brightness and every
 * light's (direction * factor) with factor > 0, then "focus" = product over the same contributions
 * of lerp(1, max(dot(avg^, c^),0), min(|c|/|avg|,1)).  Needs all shadow factors resident.
 */
__device__ __forceinline__ bool light_contrib(const ltrgpu_Light &L, V3 SP, V3 SN, float fv, V3 &out)
{
    ShadeTerms t = shade_terms(L, SP, SN);
    const float bright = (L.color.x + L.color.y + L.color.z) * (1.0f / 3.0f);
    float factor;
    if (L.type == 3u) factor = fv * t.f_ndotl * bright;
    else {
        if (!(t.pre > 0)) return false;
        factor = (L.type == 2u) ? t.f_dist * t.f_dir * fv * t.f_ndotl * bright : t.f_dist * fv * t.f_ndotl * bright;
    }
    if (!(factor > 0)) return false;
    out = t.s2l * factor;
    return true;
}

__global__ void normalmap_kernel(const ltrgpu_Light *__restrict__ lights, uint32_t n_lights,
                                 const uint8_t *__restrict__ light_inst, uint32_t n_inst,
                                 const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, const uint32_t *__restrict__ linst,
                                 uint64_t sh_begin, uint32_t n_local, const float *__restrict__ fvis, uint64_t tab_base, uint32_t tab_n,
                                 float amb_brightness, float4 *__restrict__ lnmap)
{
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_local) return;
    const uint64_t g = sh_begin + li;
    const uint32_t ti = (uint32_t)(g - tab_base);
    const V3 SP = ld3(lpos[g]), SN = ld3(lnrm[g]);
    const uint32_t inst = linst[g];
    V3 sum = SN * amb_brightness;
    int count = 1;
    for (uint32_t l = 0; l < n_lights; ++l) {
        const ltrgpu_Light L = lights[l];
        if (!light_is_supported(L.type) || !light_inst[(size_t)l * n_inst + inst]) continue;
        V3 c;
        if (light_contrib(L, SP, SN, fvis[(size_t)l * tab_n + ti], c)) { sum = sum + c; ++count; }
    }
    sum = sum / (float)count;
    float mindot = 1.0f;
    const V3 sn = norm3(sum);
    const float slen = len3(sum);
    for (uint32_t l = 0; l < n_lights; ++l) {
        const ltrgpu_Light L = lights[l];
        if (!light_is_supported(L.type) || !light_inst[(size_t)l * n_inst + inst]) continue;
        V3 c;
        if (!light_contrib(L, SP, SN, fvis[(size_t)l * tab_n + ti], c)) continue;
        float d = fmaxr(dot3(sn, norm3(c)), 0.0f);
        d = lerpf(1.0f, d, fminr(len3(c) / slen, 1.0f));
        mindot *= d;
    }
    lnmap[g] = make_float4(sn.x, sn.y, sn.z, mindot);
}

extern "C" int ltrgpu_direct_light(ltrgpu_Ctx *ctx)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    /* every early return below depends on values that are identical on every rank (collectives follow) */
    if (ctx->n_lumels == 0) return 0;
    if (ctx->n_lumels > 0x7fffffffull) { snprintf(ctx->err, sizeof(ctx->err), "too many lumels"); return 1; }
    if (ctx->n_lights == 0 && !ctx->params.normalmap) return 0;
    const uint32_t world = (uint32_t)(ctx->world > 0 ? ctx->world : 1);
    const uint32_t n_all = (uint32_t)ctx->n_lumels;
    const uint32_t n_blocks = (n_all + 1023u) >> 10;
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, st));

    /* This rank's run of 1024-lumel blocks.  One GPU: everything.  Several: contiguous runs of equal estimated march cost;
     * a rank marches AND shades the lumels of its run -- all lights of a lumel are on one rank, so the factor table never
     * leaves the GPU (lights x lumels x 4 B: it was all-reduced before, 279 MB on config 4 and growing with the light count)
     * and only the shaded colours are exchanged (16 B per lumel, independent of the light count). */
    std::vector<uint32_t> cuts(world + 1, 0);
    cuts[world] = n_blocks;
    if (world > 1) {
        if (!ctx->gatherv) { snprintf(ctx->err, sizeof(ctx->err), "sharded bake without a gather hook"); return 1; }
        uint32_t *d_block_w = nullptr, *d_cuts = nullptr;
        if (dev_alloc(ctx, &d_block_w, n_blocks)) return 1;
        if (dev_alloc(ctx, &d_cuts, (size_t)world + 1)) return 1;
        CU_TRY(ctx, cudaMemsetAsync(d_block_w, 0, (size_t)n_blocks * 4, st));
        for (uint32_t l0 = 0; l0 < ctx->n_lights; l0 += 65535u) {
            const uint32_t l1 = l0 + 65535u < ctx->n_lights ? l0 + 65535u : ctx->n_lights;
            dim3 grid(grid_for(n_all, 256), l1 - l0);
            direct_classify_kernel<<<grid, 256, 0, st>>>(ctx->d_lights, l0, l1, ctx->d_light_inst, ctx->n_inst, ctx->d_lpos, ctx->d_lnrm, ctx->d_linst, 0, n_all,
                                                         d_block_w, nullptr, nullptr);
            CU_LAUNCH_CHECK(ctx);
        }
        direct_cuts_kernel<<<1, 32, 0, st>>>(d_block_w, n_blocks, world, d_cuts);
        CU_LAUNCH_CHECK(ctx);
        CU_TRY(ctx, cudaMemcpyAsync(cuts.data(), d_cuts, (size_t)(world + 1) * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));
        lb_free(d_block_w); lb_free(d_cuts);
    }
    auto lumel_of = [&](uint32_t blk) { const uint64_t v = (uint64_t)blk << 10; return v < n_all ? v : (uint64_t)n_all; };
    const uint64_t tab_base = lumel_of(cuts[ctx->rank]);
    const uint32_t tab_n = (uint32_t)(lumel_of(cuts[ctx->rank + 1]) - tab_base);        /* may be 0: this rank only takes part in the exchange */

    /* lights are processed in chunks so that the factor table stays within a fixed budget;
     * accumulation order over chunks is still the light order */
    const bool sampled = ctx->params.shadow_mode == 1;
    const size_t budget = (size_t)8 << 30;
    uint32_t chunk = (uint32_t)(budget / ((size_t)(tab_n ? tab_n : 1) * (sampled ? 20 : 12)));
    if (chunk < 1) chunk = 1;
    if (chunk > ctx->n_lights) chunk = ctx->n_lights;
    if (ctx->params.normalmap) chunk = ctx->n_lights;           /* the normal map needs every factor resident */
    if (chunk > 65535u) chunk = 65535u;
    if (dev_alloc(ctx, &ctx->d_fvis, (size_t)chunk * tab_n)) return 1;
    if (dev_alloc(ctx, &ctx->d_active, (size_t)chunk * tab_n)) return 1;
    if (dev_alloc(ctx, &ctx->d_active_count, 2)) return 1;
    if (sampled && dev_alloc(ctx, &ctx->d_smask, (size_t)chunk * tab_n)) return 1;
    ctx->fvis_tab_base = tab_base; ctx->fvis_tab_n = tab_n;

    /* the march's work-list order (march_order_kernel) */
    uint32_t *d_order = nullptr;
    {
        const char *e = getenv("LTR_MARCH_ORDER");
        if (tab_n >= (1u << 18) && ctx->n_lights && !(e && e[0] == '0')) {
            uint32_t *k0 = nullptr, *k1 = nullptr, *v0 = nullptr, *v1 = nullptr;
            void *tmp = nullptr;
            size_t tmp_bytes = 0;
            if (dev_alloc(ctx, &k0, tab_n) || dev_alloc(ctx, &k1, tab_n) || dev_alloc(ctx, &v0, tab_n) || dev_alloc(ctx, &v1, tab_n)) return 1;
            march_order_kernel<<<grid_for(tab_n, 256), 256, 0, st>>>(ctx->d_lpos, tab_base, tab_n, k0, v0);
            CU_LAUNCH_CHECK(ctx);
            cub::DoubleBuffer<uint32_t> kb(k0, k1), vb(v0, v1);
            CU_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)tab_n, 0, 30, st));
            CU_TRY(ctx, lb_malloc(&tmp, tmp_bytes ? tmp_bytes : 16));
            CU_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int)tab_n, 0, 30, st));
            ctx->host_counters.kernel_launches += 4;
            d_order = vb.Current();
            CU_TRY(ctx, cudaStreamSynchronize(st));
            lb_free(tmp); lb_free(k0); lb_free(k1); lb_free(d_order == v0 ? v1 : v0);
        }
    }

    std::vector<cudaEvent_t> mev;
    for (uint32_t l0 = 0; tab_n && l0 < ctx->n_lights; l0 += chunk) {
        uint32_t l1 = l0 + chunk < ctx->n_lights ? l0 + chunk : ctx->n_lights;
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_active_count, 0, 8, st));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_fvis, 0, (size_t)(l1 - l0) * tab_n * 4, st));
        dim3 grid(grid_for(tab_n, 256), l1 - l0);
        direct_classify_kernel<<<grid, 256, 0, st>>>(ctx->d_lights, l0, l1, ctx->d_light_inst, ctx->n_inst, ctx->d_lpos, ctx->d_lnrm,
                                                     ctx->d_linst, tab_base, tab_n, nullptr, ctx->d_active, ctx->d_active_count, d_order);
        CU_LAUNCH_CHECK(ctx);
        cudaEvent_t m0, m1;
        CU_TRY(ctx, cudaEventCreate(&m0)); mev.push_back(m0);
        CU_TRY(ctx, cudaEventCreate(&m1)); mev.push_back(m1);
        CU_TRY(ctx, cudaEventRecord(m0, st));
        unsigned blocks = (unsigned)ctx->num_sms * 16;
        if (!sampled) {
            direct_march_kernel<<<blocks, LB_BLOCK, 0, st>>>(ctx->d_lights, ctx->d_bvh, ctx->d_bvh4, ctx->d_ptris, ctx->d_lpos, ctx->d_lnrm, tab_base,
                                                             tab_n, ctx->d_active, ctx->d_active_count, ctx->d_active_count + 1, l0, ctx->d_fvis, ctx->d_counters);
            CU_LAUNCH_CHECK(ctx);
        } else {
            uint32_t spp = 1;
            for (uint32_t l = l0; l < l1; ++l) if (ctx->h_lights[l].n_samples > spp) spp = ctx->h_lights[l].n_samples;
            CU_TRY(ctx, cudaMemsetAsync(ctx->d_smask, 0, (size_t)(l1 - l0) * tab_n * 8, st));     /* pairs that are not listed keep mask 0 */
            direct_sampled_kernel<<<(unsigned)ctx->num_sms * 32, LB_BLOCK, 0, st>>>(ctx->d_lights, ctx->d_light_samples, ctx->d_bvh4, ctx->d_raytris, ctx->d_lpos,
                                                                                   ctx->d_lnrm, tab_base, tab_n, ctx->d_active, ctx->d_active_count, spp, l0,
                                                                                   ctx->d_smask, ctx->d_counters);
            CU_LAUNCH_CHECK(ctx);
            sampled_resolve_kernel<<<(unsigned)ctx->num_sms * 16, 256, 0, st>>>(ctx->d_lights, ctx->d_active, ctx->d_active_count, tab_n, l0,
                                                                                               ctx->d_smask, ctx->d_fvis);
            CU_LAUNCH_CHECK(ctx);
        }
        CU_TRY(ctx, cudaEventRecord(m1, st));
        direct_accumulate_kernel<<<grid_for(tab_n, 256), 256, 0, st>>>(ctx->d_lights, l0, l1, ctx->d_light_inst, ctx->n_inst, ctx->d_lpos,
                                                                      ctx->d_lnrm, ctx->d_linst, tab_base, tab_n, ctx->d_fvis, tab_base, tab_n, ctx->d_lrgb);
        CU_LAUNCH_CHECK(ctx);
    }
    if (d_order) { CU_TRY(ctx, cudaStreamSynchronize(st)); lb_free(d_order); d_order = nullptr; }      /* lb_free recycles at once: the listing passes must be done with it */
    if (ctx->params.normalmap) {
        /* also without any light: the reference then writes normalize(N * ambient brightness), focus 1 (lighter.cpp:977-1012) */
        if (dev_alloc(ctx, &ctx->d_lnmap, ctx->n_lumels + LB_PAD)) return 1;
        if (tab_n) {
            normalmap_kernel<<<grid_for(tab_n, 256), 256, 0, st>>>(ctx->d_lights, ctx->n_lights, ctx->d_light_inst, ctx->n_inst, ctx->d_lpos,
                                                                  ctx->d_lnrm, ctx->d_linst, tab_base, tab_n, ctx->d_fvis, tab_base, tab_n,
                                                                  ctx->params.amb_brightness, ctx->d_lnmap);
            CU_LAUNCH_CHECK(ctx);
        }
    }
    if (world > 1) {
        /* every rank gets the shaded colours (and normal-map terms) of every lumel: rank r's run is broadcast from r */
        std::vector<uint64_t> off(world + 1);
        for (uint32_t r = 0; r <= world; ++r) off[r] = lumel_of(cuts[r]) * sizeof(float4);
        if (ctx->n_lights && ctx->gatherv(ctx->allgather_user, ctx->d_lrgb, off.data(), st)) { snprintf(ctx->err, sizeof(ctx->err), "direct light: exchange of the lumel colours failed"); return 1; }
        if (ctx->params.normalmap && ctx->gatherv(ctx->allgather_user, ctx->d_lnmap, off.data(), st)) { snprintf(ctx->err, sizeof(ctx->err), "direct light: exchange of the normal-map terms failed"); return 1; }
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0, march_ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    for (size_t k = 0; k + 1 < mev.size(); k += 2) { float m = 0; cudaEventElapsedTime(&m, mev[k], mev[k + 1]); march_ms += m; }
    for (cudaEvent_t e : mev) cudaEventDestroy(e);
    ctx->host_counters.ms_direct += ms;
    ctx->host_counters.ms_march += march_ms;
    return 0;
}

extern "C" int ltrgpu_download_shadow_factors(ltrgpu_Ctx *ctx, uint32_t light, float *out)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t n_local = ctx->sh_end - ctx->sh_begin;
    if (!ctx->d_fvis || light >= ctx->n_lights) { snprintf(ctx->err, sizeof(ctx->err), "no shadow factors"); return 1; }
    if (ctx->fvis_tab_base != ctx->sh_begin || ctx->fvis_tab_n != n_local) { snprintf(ctx->err, sizeof(ctx->err), "shadow-factor dumps are single-GPU only"); return 1; }
    CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_fvis + (size_t)light * ctx->fvis_tab_n + (ctx->sh_begin - ctx->fvis_tab_base), n_local * 4,
                                cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int ltrgpu_download_shadow_masks(ltrgpu_Ctx *ctx, uint32_t light, uint64_t *out)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t n_local = ctx->sh_end - ctx->sh_begin;
    if (!ctx->d_smask || light >= ctx->n_lights) { snprintf(ctx->err, sizeof(ctx->err), "no shadow masks (sampled mode only)"); return 1; }
    if (ctx->fvis_tab_base != ctx->sh_begin || ctx->fvis_tab_n != n_local) { snprintf(ctx->err, sizeof(ctx->err), "shadow-mask dumps are single-GPU only"); return 1; }
    CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_smask + (size_t)light * ctx->fvis_tab_n + (ctx->sh_begin - ctx->fvis_tab_base), n_local * 8,
                                cudaMemcpyDeviceToHost, ctx->stream));      /* several GPUs: only the pairs this rank marched are set */
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
