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
 * GPU formulation (the reference's loop is a serial O(N^2) sweep with a stored link list):
 *   1. lumels are sorted along a Morton curve so that 128 consecutive lumels form a spatially compact
 *      TILE with a tight position box and a tight box of normal components;
 *   2. a tile pair is skipped when NO pair in it can link: box distance > 17.85 (f < 0.001 because
 *      dotA*dotB <= len^2), or interval bounds give max dotA <= 0.001, max dotB <= 0.001, or
 *      max dotA * max dotB / (min len^4 * pi) < 0.001 -- exact pruning, no approximation;
 *   3. surviving tile pairs are swept once per UNORDERED pair (the factor is symmetric bit for bit:
 *      P_j-P_i == -(P_i-P_j) exactly in IEEE arithmetic); pairs passing the arithmetic test go
 *      through a per-CTA shared-memory queue to a candidate list;
 *   4. one thread per candidate traces the any-hit segment (from the lower lumel index to the
 *      higher, as the reference does) on the scene BVH; blocked pairs become two directed links;
 *   5. links are radix-sorted by (row, partner index) -> CSR rows in the reference's accumulation
 *      order; a bounce is a pure gather, one thread per row.
 * Multi-GPU: rows are sharded by Morton position (contiguous tile ranges = compact regions); per
 * bounce only E_j = (out_j*diffuse_j)*area_j (16 B/lumel) is all-gathered over NCCL.
 */
#include "gpu_internal.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <time.h>

#include <vector>

#include "rad_cull.h"

#define RAD_TILE 128
#define RAD_QUEUE 1024          /* per-CTA shared-memory candidate queue */
#define RAD_PAD 0xffffffffu     /* storage slot without a lumel */
#define RAD_GROUP_DEFAULT 8     /* lumels per column group of the warp-level culling */

struct RadCand { uint32_t a, b; float factor; };       /* sorted positions of the two lumels */


__global__ void rad_morton_kernel(const float4 *__restrict__ lpos, uint64_t n, float3 lo, float3 inv_ext, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = lpos[i];
    auto q = [](float v) { int x = (int)(v * 1023.0f); x = x < 0 ? 0 : (x > 1023 ? 1023 : x); return (uint32_t)x; };
    auto spread = [](uint32_t x) { x &= 0x3ffu; x = (x | (x << 16)) & 0x30000ffu; x = (x | (x << 8)) & 0x300f00fu; x = (x | (x << 4)) & 0x30c30c3u; x = (x | (x << 2)) & 0x9249249u; return x; };
    uint32_t k = spread(q((p.x - lo.x) * inv_ext.x)) | (spread(q((p.y - lo.y) * inv_ext.y)) << 1) | (spread(q((p.z - lo.z) * inv_ext.z)) << 2);
    keys[i] = k;
    vals[i] = (uint32_t)i;
}

__global__ void rad_bounds_reduce_kernel(const float4 *__restrict__ lpos, uint64_t n, float *__restrict__ out6)
{
    /* scene bounds of the lumel positions: min/max via float atomics on the ordered-int trick */
    __shared__ float slo[3][8], shi[3][8];
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float4 p = lpos[i];
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0) for (int a = 0; a < 3; ++a) { slo[a][threadIdx.x >> 5] = lo[a]; shi[a][threadIdx.x >> 5] = hi[a]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        int a = threadIdx.x;
        float l = slo[a][0], h = shi[a][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { l = fminf(l, slo[a][w]); h = fmaxf(h, shi[a][w]); }
        /* float atomic min/max through the sign-aware integer ordering */
        int *pl = (int *)(out6 + a), *ph = (int *)(out6 + 3 + a);
        if (l >= 0) atomicMin(pl, __float_as_int(l)); else atomicMax((unsigned *)pl, __float_as_uint(l));
        if (h >= 0) atomicMax(ph, __float_as_int(h)); else atomicMin((unsigned *)ph, __float_as_uint(h));
    }
}

/* Morton order -> storage order.  Tiles of 128 consecutive Morton positions are dealt round-robin
 * to the ranks, each rank's tiles stored contiguously: every rank gets a statistically identical
 * sample of the scene (vertical surfaces generate far more candidates than floors; a contiguous
 * split of the curve measured 4x imbalance), and its rows stay one contiguous all-gather chunk. */
__global__ void rad_deal_kernel(const uint32_t *__restrict__ smort, uint64_t n, uint64_t n_pad, uint32_t world, uint32_t tiles_per_rank,
                                uint32_t *__restrict__ sidx)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pad) return;
    const uint32_t t = (uint32_t)(k / RAD_TILE), w = (uint32_t)(k % RAD_TILE);
    const uint32_t st = (t % world) * tiles_per_rank + t / world;
    sidx[(uint64_t)st * RAD_TILE + w] = k < n ? smort[k] : RAD_PAD;
}

/* gather the sorted geometry; the reference re-normalises the normal here (lighter.cpp:680) */
__global__ void rad_gather_kernel(const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, const uint32_t *__restrict__ sidx, uint64_t n,
                                  uint64_t n_pad, float4 *__restrict__ spos, float4 *__restrict__ snrm)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pad) return;
    const uint32_t i = sidx[k];
    if (i != RAD_PAD) {
        spos[k] = lpos[i];
        V3 N = norm3(ld3(lnrm[i]));
        snrm[k] = make_float4(N.x, N.y, N.z, 0.f);
    } else {                                    /* padding lumels can never link: zero normal */
        spos[k] = make_float4(3.0e30f, 3.0e30f, 3.0e30f, 0.f);
        snrm[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

/* Bounds of every 128-lumel tile (tb) and of every G-lumel group inside it (tbg, 128/G per tile).
 * Groups are runs of G consecutive Morton positions: a few texels of one surface, so their position
 * and normal boxes are tight -- that is what makes the warp-level culling in the pair sweep bite. */
template <int G>
__global__ void rad_tile_bounds_kernel(const float4 *__restrict__ spos, const float4 *__restrict__ snrm, const uint32_t *__restrict__ sidx, TileBounds *__restrict__ tb,
                                       TileBounds *__restrict__ tb32 /* per warp of row lumels */, TileBounds *__restrict__ tbg /* per G-lumel column group */)
{
    __shared__ float s[12][RAD_TILE / 32];
    const uint64_t k = (uint64_t)blockIdx.x * RAD_TILE + threadIdx.x;
    float v[12];
    for (int a = 0; a < 6; ++a) { v[a] = INFINITY; v[6 + a] = -INFINITY; }          /* [0..2] plo, [3..5] nlo, [6..8] phi, [9..11] nhi */
    if (sidx[k] != RAD_PAD) {
        float4 p = spos[k], q = snrm[k];
        v[0] = v[6] = p.x; v[1] = v[7] = p.y; v[2] = v[8] = p.z;
        v[3] = v[9] = q.x; v[4] = v[10] = q.y; v[5] = v[11] = q.z;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int a = 0; a < 12; ++a) {
            float t = __shfl_xor_sync(0xffffffffu, v[a], o);
            v[a] = a < 6 ? fminf(v[a], t) : fmaxf(v[a], t);
        }
        if (G < 32 && o * 2 == G && (threadIdx.x & (G - 1)) == 0) {                            /* the xor butterfly has just covered a G-lane group */
            TileBounds w;
            w.plo = make_float4(v[0], v[1], v[2], (v[3] == v[9] && v[4] == v[10] && v[5] == v[11]) ? 1.f : 0.f);      /* .w: the group has one normal (rad_cull.h LB_RAD_FLATN) */
            w.nlo = make_float4(v[3], v[4], v[5], 0.f);
            w.phi = make_float4(v[6], v[7], v[8], 0.f); w.nhi = make_float4(v[9], v[10], v[11], 0.f);
            tbg[(size_t)blockIdx.x * (RAD_TILE / G) + threadIdx.x / G] = w;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 12; ++a) s[a][threadIdx.x >> 5] = v[a];
        TileBounds w;
        w.plo = make_float4(v[0], v[1], v[2], 0.f); w.nlo = make_float4(v[3], v[4], v[5], 0.f);
        w.phi = make_float4(v[6], v[7], v[8], 0.f); w.nhi = make_float4(v[9], v[10], v[11], 0.f);
        tb32[(size_t)blockIdx.x * (RAD_TILE / 32) + (threadIdx.x >> 5)] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int a = 0; a < 12; ++a)
            for (int w = 1; w < RAD_TILE / 32; ++w) s[a][0] = a < 6 ? fminf(s[a][0], s[a][w]) : fmaxf(s[a][0], s[a][w]);
        TileBounds t;
        t.plo = make_float4(s[0][0], s[1][0], s[2][0], 0.f); t.nlo = make_float4(s[3][0], s[4][0], s[5][0], 0.f);
        t.phi = make_float4(s[6][0], s[7][0], s[8][0], 0.f); t.nhi = make_float4(s[9][0], s[10][0], s[11][0], 0.f);
        tb[blockIdx.x] = t;
    }
}

/* ---- mbarrier / bulk-copy (TMA 1-D) primitives: one lane stages a column tile, the warp waits on the barrier ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

#define RAD_STAGE 288            /* per-warp stage: 31 left over + 8 pairs x 32 lanes, (row lane << 7 | k) each */
#ifndef RAD_CHUNK
#define RAD_CHUNK 1024u          /* candidate slots a warp reserves from the global counter at a time (512 / 2048 measured: profiles/r02_ab_runs.md) */
#endif
#define RAD_WARPS 4              /* warps per CTA; the warps do not interact */
/* Work items are taken from the cursor in RUNS of consecutive items.  Items are row-warp-major (a row warp has ~38 of them on
 * config 4), so a run is mostly ONE row warp against neighbouring super-tiles, and the 1024-slot candidate chunks a sweep warp
 * fills -- one entry set each in the visibility pass -- hold one row warp instead of two or three that lie units apart:
 * visibility 208.6 -> 195.6 ms (run 1 / 2 / 4 / 8 / 16 / 32 / 64 / 128: 208.6 / 205.7 / 204.6 / 199.9 / 197.8 / 196.5 / 195.6 / 195.3).
 * GUIDED: towards the end of a launch the runs shrink to a quarter of a warp's fair share of what is left, so that the last warps
 * finish together (fixed runs of 16 cost the sweep itself 73.5 -> 76.8 ms; guided 73.6). */
#ifndef RAD_ITEM_GUIDED
#define RAD_ITEM_GUIDED 1
#endif
#ifndef RAD_ITEM_RUN
#define RAD_ITEM_RUN 64
#endif

template <int G> struct __align__(128) RadColBuf {          /* one staged column tile */
    float4 sp[RAD_TILE], sn[RAD_TILE];
    TileBounds gb[RAD_TILE / G];
};

/* Output cursor of one warp into the global candidate array: slots are reserved RAD_CHUNK at a time (one
 * global atomic per 1024 candidates), unused slots of a chunk are marked RAD_PAD for the visibility pass. */
struct RadOut { unsigned long long pos, end; bool dead; };

__device__ __forceinline__ void rad_out_pad(RadOut &o, RadCand *__restrict__ cand, unsigned lane)
{
    for (unsigned long long e = o.pos + lane; e < o.end; e += 32) cand[e].a = RAD_PAD;
    o.pos = o.end;
}

/* Exact evaluation of one staged pair per lane (the reference's arithmetic, lighter.cpp:735-750) and
 * warp-aggregated append of the linking ones.  `e` = (row lane << 7) | k. */
__device__ __forceinline__ void rad_exact_pair(unsigned e, const V3 &Pr, const V3 &Nr, const float4 *sp, const float4 *sn, uint32_t row_base, uint32_t col_base,
                                               bool valid, bool diag, unsigned lane, unsigned lt_mask, RadOut &o,
                                               RadCand *__restrict__ cand, unsigned long long cand_cap, unsigned long long *cand_count)
{
    const unsigned rl = e >> 7, k = e & 0x7fu;
    if (diag && k <= (row_base & (RAD_TILE - 1)) + rl) valid = false;     /* diagonal tile: only k > my own position */
    /* the row lumel lives in the registers of lane rl of this warp */
    const V3 Pi = mk3(__shfl_sync(0xffffffffu, Pr.x, rl), __shfl_sync(0xffffffffu, Pr.y, rl), __shfl_sync(0xffffffffu, Pr.z, rl));
    const V3 Ni = mk3(__shfl_sync(0xffffffffu, Nr.x, rl), __shfl_sync(0xffffffffu, Nr.y, rl), __shfl_sync(0xffffffffu, Nr.z, rl));
    bool ok = false;
    RadCand c = { 0, 0, 0.f };
    if (valid) {
        const V3 d = ld3(sp[k]) - Pi;
        const float dr = dot3(Ni, d);
        const float dj = -dot3(ld3(sn[k]), d);
        if (!(dr <= LB_SMALL || dj <= LB_SMALL)) {
            const float lensq = lensq3(d);
            const float f = dr * dj / (lensq * lensq * 3.14159274101257324f);
            if (!(f < LB_SMALL)) { ok = true; c.a = row_base + rl; c.b = col_base + k; c.factor = f; }
        }
    }
    const unsigned okm = __ballot_sync(0xffffffffu, ok);
    if (!okm) return;
    const unsigned m = (unsigned)__popc(okm);
    if (o.pos + m > o.end) {                                              /* warp-uniform: next chunk */
        if (!o.dead) rad_out_pad(o, cand, lane);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cand_count, (unsigned long long)RAD_CHUNK);
        base = __shfl_sync(0xffffffffu, base, 0);
        o.dead = base + RAD_CHUNK > cand_cap;                             /* the host sees count > capacity and redoes the batch smaller */
        o.pos = base; o.end = base + RAD_CHUNK;
    }
    if (ok && !o.dead) cand[o.pos + __popc(okm & lt_mask)] = c;
    o.pos += m;
}

/* Sweep one staged column tile against the warp's 32 row lumels. */
template <int G>
__device__ __forceinline__ void rad_sweep_tile(const RadColBuf<G> &B, const TileBounds &Rw, uint16_t *stage, const V3 &Pr, const V3 &Nr,
                                               uint32_t row_base, uint32_t ct, bool diag, unsigned lane, unsigned lt_mask, RadOut &o, unsigned &tested,
                                               RadCand *__restrict__ cand, unsigned long long cand_cap, unsigned long long *cand_count)
{
    constexpr int NG = RAD_TILE / G;
    constexpr int CH = G < 8 ? G : 8;                      /* pairs per lane between two drain checks */
    /* warp x group culling, one group per lane */
    unsigned gm = __ballot_sync(0xffffffffu, lane < (unsigned)NG && tile_pair_may_link(Rw, B.gb[lane < (unsigned)NG ? lane : 0]));
    unsigned wcount = 0;                                   /* staged survivors (warp-uniform) */
    while (gm) {
        const unsigned g = (unsigned)__ffs(gm) - 1u;
        gm &= gm - 1u;
        /* lumel x group: every lane tests ITS row lumel (exact position and normal) against the group's bounds; the group
         * is swept only if some lane can link.  The warp-level interval test above loses the correlation between a row's
         * position and its normal; this one keeps it and rejects about half of the groups that pass it (config 4). */
#if LB_RAD_TWOSTAGE
        {
            RowGroupA ta;
            const bool a_ok = row_group_step_a(Pr, Nr, B.gb[g], ta);
            if (!__any_sync(0xffffffffu, a_ok)) continue;
            if (!__any_sync(0xffffffffu, a_ok && row_group_step_b(B.gb[g], ta))) continue;
        }
#else
        if (!__any_sync(0xffffffffu, row_group_may_link(Pr, Nr, B.gb[g]))) continue;
#endif
        tested += G;                                       /* per lane: pairs this lane goes on to test */
#pragma unroll 1
        for (unsigned q0 = 0; q0 < (unsigned)G; q0 += CH) {
            const unsigned kb = g * G + q0;
            unsigned bits = 0;
#pragma unroll
            for (unsigned q = 0; q < (unsigned)CH; ++q) {
                const float4 pj = B.sp[kb + q], nj = B.sn[kb + q];
                if (rad_fast_filter(Pr, Nr, pj, nj)) bits |= 1u << q;
            }
            if (__any_sync(0xffffffffu, bits != 0)) {
                /* compaction: exclusive scan of the per-lane survivor counts, then every lane appends its own */
                const unsigned cnt = (unsigned)__popc(bits);
                unsigned incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (unsigned)d) incl += t; }
                unsigned at = wcount + incl - cnt;
                wcount += __shfl_sync(0xffffffffu, incl, 31);
                while (bits) {
                    const unsigned q = (unsigned)__ffs(bits) - 1u;
                    bits &= bits - 1u;
                    stage[at++] = (uint16_t)((lane << 7) | (kb + q));
                }
                while (wcount >= 32u) {
                    __syncwarp();
                    wcount -= 32u;
                    rad_exact_pair(stage[wcount + lane], Pr, Nr, B.sp, B.sn, row_base, ct * RAD_TILE, true, diag, lane, lt_mask, o, cand, cand_cap, cand_count);
                }
                __syncwarp();
            }
        }
    }
    if (wcount) {                                          /* partial drain before the buffer is reused */
        __syncwarp();
        rad_exact_pair(lane < wcount ? stage[lane] : (uint16_t)0, Pr, Nr, B.sp, B.sn, row_base, ct * RAD_TILE, lane < wcount, diag, lane, lt_mask, o, cand, cand_cap, cand_count);
    }
    __syncwarp();
}

/*
 * Pair sweep.  The unit of work is a ROW WARP: 32 consecutive (Morton-sorted) lumels, one per lane, held
 * in registers.  Warps are persistent and independent -- each pulls row warps from a global cursor, so a
 * warp next to a wall (thousands of column tiles) never holds up its neighbours over open floor, and
 * there is not a single CTA barrier in the kernel (the previous CTA-per-row-tile version spent most of
 * its time in barriers with one busy warp out of four).
 *
 * Per row warp, four levels of exact culling (interval bounds: no linking pair is ever rejected), then
 * a two-phase pair test:
 *   warp x super-tile   32 super-tiles (32 Morton-consecutive tiles = 4096 lumels each) per step, one per lane;
 *   warp x tile         the 32 tiles of a surviving super-tile, one per lane; a tile pair is swept once in
 *                       the whole job: by the owner of the tile that comes first on the Morton curve;
 *   staging             surviving column tiles (positions, normals, group bounds: 5 KB) are brought into
 *                       this warp's shared-memory double buffer by 1-D bulk copies (TMA) that complete on an
 *                       mbarrier; the copy of tile i+1 is in flight while tile i is swept;
 *   warp x group        each lane tests the warp's rows against one group of G column lumels;
 *   lumel x lumel       FAST filter, all 32 lanes in lock step, no branches: FMA dots and the factor
 *                       inequality without the division, thresholds lowered by 10 % so that rounding
 *                       differences to the exact arithmetic can only let extra pairs through.  Survivors
 *                       (~3 % of the pairs) are compacted into a per-warp stage;
 *   drain               every 32 staged survivors are evaluated one per lane with the reference's exact
 *                       operation order (mul/add dots, IEEE division) -- full SIMT efficiency on the
 *                       expensive path -- and the linking ones appended to the global candidate array.
 */
template <int G>
__global__ void __launch_bounds__(RAD_WARPS * 32)
rad_candidates_kernel(const float4 *__restrict__ spos, const float4 *__restrict__ snrm, const TileBounds *__restrict__ tb,
                      const TileBounds *__restrict__ tb32, const TileBounds *__restrict__ tbg, const TileBounds *__restrict__ tbs,
                      uint32_t n_tiles, uint32_t world, uint32_t tiles_per_rank, const uint2 *__restrict__ items, uint32_t n_items,
                      uint32_t *item_cursor, RadCand *__restrict__ cand, unsigned long long cand_cap, unsigned long long *cand_count, unsigned long long *counters)
{
    constexpr int NG = RAD_TILE / G;
    constexpr uint32_t TILE_BYTES = RAD_TILE * 16u, GB_BYTES = NG * (uint32_t)sizeof(TileBounds);
    extern __shared__ __align__(128) unsigned char rad_smem[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    RadColBuf<G> *buf = reinterpret_cast<RadColBuf<G> *>(rad_smem) + warp * 2;
    unsigned char *tail = rad_smem + sizeof(RadColBuf<G>) * 2 * RAD_WARPS;
    TileBounds *Rw = reinterpret_cast<TileBounds *>(tail) + warp;                                  tail += sizeof(TileBounds) * RAD_WARPS;
    uint64_t *bar = reinterpret_cast<uint64_t *>(tail) + warp * 2;                                 tail += 8 * 2 * RAD_WARPS;
    uint16_t *stage = reinterpret_cast<uint16_t *>(tail) + warp * RAD_STAGE;
    const uint32_t bar0 = smem_u32(bar);                   /* barrier of buffer b: bar0 + 8 b */
    if (lane == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned parity = 0;                                   /* bit b: phase parity buffer b's barrier completes next */
    unsigned nb = 0;                                       /* buffer the next copy goes into */
    RadOut o = { 0, 0, false };
    unsigned tested = 0, tile_loads = 0;
    for (uint32_t it = 0, it_end = 0;; ++it) {
        if (it >= it_end) {                                /* next run of up to RAD_ITEM_RUN consecutive items */
            uint32_t run = (uint32_t)RAD_ITEM_RUN;
            if (lane == 0) {
#if RAD_ITEM_GUIDED
                /* guided: towards the end of a launch the runs shrink (what is left / 4 runs per warp), so that the last warps finish together */
                const uint32_t seen = *(volatile uint32_t *)item_cursor;
                const uint32_t left = seen < n_items ? n_items - seen : 0u;
                const uint32_t fair = left / (gridDim.x * (uint32_t)RAD_WARPS * 4u);
                run = fair < run ? (fair ? fair : 1u) : run;
#endif
                it = atomicAdd(item_cursor, run);
            }
            it = __shfl_sync(0xffffffffu, it, 0);
            run = __shfl_sync(0xffffffffu, run, 0);
            if (it >= n_items) break;
            it_end = it + run < n_items ? it + run : n_items;
        }
        const uint2 item = items[it];
        const uint32_t rw = item.x;                        /* rows rw*32 .. rw*32+31 (storage order) */
        const uint32_t rt = rw / (RAD_TILE / 32);
        const V3 Pr = ld3(spos[(size_t)rw * 32 + lane]), Nr = ld3(snrm[(size_t)rw * 32 + lane]);
        __syncwarp();
        if (lane < 4) reinterpret_cast<float4 *>(Rw)[lane] = reinterpret_cast<const float4 *>(tb32 + rw)[lane];
        __syncwarp();
        const uint32_t mrt = (rt % tiles_per_rank) * world + rt / tiles_per_rank;       /* my tile's index on the Morton curve */
        const uint32_t mt = item.y * 32u + lane;           /* Morton tile index -> storage index (see rad_deal_kernel) */
        const uint32_t c = (mt % world) * tiles_per_rank + mt / world;
        unsigned tmask = __ballot_sync(0xffffffffu, mt < n_tiles && mt >= mrt && tile_pair_may_link(*Rw, tb[mt < n_tiles ? c : 0]));
        int pend = -1;
        while (tmask) {
            const unsigned tl = (unsigned)__ffs(tmask) - 1u;
            tmask &= tmask - 1u;
            const uint32_t ct = __shfl_sync(0xffffffffu, c, tl);
            if (lane == 0) {                               /* stage column tile ct into buf[nb] */
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const uint32_t ba = bar0 + 8u * nb;
                mbar_expect_tx(ba, 2u * TILE_BYTES + GB_BYTES);
                bulk_g2s(smem_u32(buf[nb].sp), spos + (size_t)ct * RAD_TILE, TILE_BYTES, ba);
                bulk_g2s(smem_u32(buf[nb].sn), snrm + (size_t)ct * RAD_TILE, TILE_BYTES, ba);
                bulk_g2s(smem_u32(buf[nb].gb), tbg + (size_t)ct * NG, GB_BYTES, ba);
                ++tile_loads;
            }
            if (pend >= 0) {
                const unsigned pb = nb ^ 1u;
                while (!mbar_try_wait(bar0 + 8u * pb, (parity >> pb) & 1u)) { }
                parity ^= 1u << pb;
                rad_sweep_tile<G>(buf[pb], *Rw, stage, Pr, Nr, rw * 32u, (uint32_t)pend, (uint32_t)pend == rt, lane, lt_mask, o, tested, cand, cand_cap, cand_count);
            }
            pend = (int)ct;
            nb ^= 1u;
        }
        if (pend >= 0) {
            const unsigned pb = nb ^ 1u;
            while (!mbar_try_wait(bar0 + 8u * pb, (parity >> pb) & 1u)) { }
            parity ^= 1u << pb;
            rad_sweep_tile<G>(buf[pb], *Rw, stage, Pr, Nr, rw * 32u, (uint32_t)pend, (uint32_t)pend == rt, lane, lt_mask, o, tested, cand, cand_cap, cand_count);
        }
    }
    if (!o.dead) rad_out_pad(o, cand, lane);
    count_add(counters, CNT_RAD_PAIRS, tested);
    count_add(counters, CNT_RAD_TILE_LOADS, tile_loads);
}

/* bounds of every super-tile = 32 consecutive tiles ON THE MORTON CURVE (storage order deals tiles round-robin to ranks) */
__global__ void rad_super_bounds_kernel(const TileBounds *__restrict__ tb, uint32_t n_tiles, uint32_t n_super, uint32_t world, uint32_t tiles_per_rank,
                                        TileBounds *__restrict__ tbs)
{
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (s >= n_super) return;
    const uint32_t mt = s * 32u + lane;
    float v[12];
    for (int a = 0; a < 6; ++a) { v[a] = INFINITY; v[6 + a] = -INFINITY; }
    if (mt < n_tiles) {
        const TileBounds t = tb[(mt % world) * tiles_per_rank + mt / world];
        v[0] = t.plo.x; v[1] = t.plo.y; v[2] = t.plo.z; v[3] = t.nlo.x; v[4] = t.nlo.y; v[5] = t.nlo.z;
        v[6] = t.phi.x; v[7] = t.phi.y; v[8] = t.phi.z; v[9] = t.nhi.x; v[10] = t.nhi.y; v[11] = t.nhi.z;
    }
#pragma unroll
    for (int a = 0; a < 12; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            const float t = __shfl_xor_sync(0xffffffffu, v[a], o);
            v[a] = a < 6 ? fminf(v[a], t) : fmaxf(v[a], t);
        }
    if (lane == 0) {
        TileBounds w;
        w.plo = make_float4(v[0], v[1], v[2], 0.f); w.nlo = make_float4(v[3], v[4], v[5], 0.f);
        w.phi = make_float4(v[6], v[7], v[8], 0.f); w.nhi = make_float4(v[9], v[10], v[11], 0.f);
        tbs[s] = w;
    }
}

/*
 * Work items of the pair sweep: (row warp, super-tile) pairs that survive the interval test.  One warp per
 * row warp, 32 super-tiles per step (one per lane).  FILL = false counts, FILL = true writes the items at
 * the scanned offsets -- row-warp-major, so warps running at the same time sweep neighbouring rows against
 * the same column tiles (L2 locality).  An item is at most 32 column tiles of work, which bounds the tail of
 * a launch; a whole row warp next to a wall is ~400 tiles = milliseconds.
 */
template <bool FILL>
__global__ void rad_items_kernel(const TileBounds *__restrict__ tb32, const TileBounds *__restrict__ tbs, uint32_t n_super, uint32_t world, uint32_t tiles_per_rank,
                                 uint32_t first_row_warp, uint32_t n_row_warps, uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets, uint2 *__restrict__ items)
{
    const uint32_t rwi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (rwi >= n_row_warps) return;
    const uint32_t rw = first_row_warp + rwi, rt = rw / (RAD_TILE / 32);
    const TileBounds R = tb32[rw];
    const uint32_t mrt = (rt % tiles_per_rank) * world + rt / tiles_per_rank;
    uint32_t n = 0;
    const uint32_t off = FILL ? offsets[rwi] : 0u;
    for (uint32_t sbase = (mrt / 32u) & ~31u; sbase < n_super; sbase += 32u) {
        const uint32_t s = sbase + lane;
        const bool ok = s < n_super && s * 32u + 31u >= mrt && tile_pair_may_link(R, tbs[s < n_super ? s : 0]);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (FILL && ok) items[off + n + __popc(m & ((1u << lane) - 1u))] = make_uint2(rw, s);
        n += (uint32_t)__popc(m);
    }
    if (!FILL && lane == 0) counts[rwi] = n;
}

template <int G> static size_t rad_sweep_smem() { return sizeof(RadColBuf<G>) * 2 * RAD_WARPS + sizeof(TileBounds) * RAD_WARPS + 16 * RAD_WARPS + 2 * RAD_STAGE * RAD_WARPS; }

/*
 * Visibility of the candidates: blocked segment -> directed link(s) keyed (row sorted position, partner original index).
 *
 * The candidate array is a sequence of RAD_CHUNK-slot chunks, each written by ONE sweep warp: the candidates of one row
 * warp (32 neighbouring lumels) against a few Morton-consecutive column tiles, i.e. a spatially compact bundle of
 * segments.  A warp takes whole chunks.  ENTRY = true: it first reduces the chunk's bounding box (one extra pass over the
 * 12-byte candidates; the positions come out of L1/L2 again in the second pass), lane 0 descends the scene tree once for
 * the whole chunk (bvh_entry.h) and leaves the entry set in shared memory; the 1024 segments then start at the entry set
 * instead of the root.  ENTRY = false is the plain walk from the root (kept for A/B measurement: LTR_RAD_ENTRY=0).
 */
#ifndef LB_VIS_FLUSH
#define LB_VIS_FLUSH 10
#endif
#ifndef LB_VIS_FMA
#define LB_VIS_FMA 0            /* 1 = box tests in FMA form with an exact re-test at the leaves (gpu_internal.cuh: bvh4_anyhit_entries_f): measured SLOWER on B200 (271 vs 235 ms, profiles/r02_ab_runs.md), kept for A/B */
#endif
#ifndef LB_VIS_LEAN
#define LB_VIS_LEAN 1           /* register-lean batch loop of the visibility kernel (see there); 0 = the round-1 software pipeline */
#endif
#ifndef LB_VIS_ENTRY2
#define LB_VIS_ENTRY2 1         /* 1 = version-2 entry sets (bvh_entry.h: shaft-culled search, leaf entries, <= BVH_ENTRY2_MAX entries): 224.5 -> 211.1 ms on config 4; 0 = version 1 (A/B) */
#endif
#ifndef LB_VIS_PACKET
#define LB_VIS_PACKET 0         /* 1 = per batch of 32 segments the WARP walks the tree once for the batch's shaft and lists the leaves in it; a ray only tests
                                 * the listed leaf boxes (needs LB_VIS_ENTRY2).  Bit-identical, and measured 3.8x SLOWER on config 4 (785.7 vs 208.6 ms,
                                 * profiles/r02_ab_runs.md): a real batch's shaft holds 35 leaves, not the 7 of the offline model, and the
                                 * warp's walk is one serial chain of dependent node loads.  Kept for A/B only */
#endif
#define VIS_LMAX 32             /* leaves listed per ray phase: one bit each in a ray's hit mask */
struct __align__(16) VisPacket {                  /* per warp, shared memory */
    float lo[VIS_LMAX][4];                        /* leaf box lo xyz + leaf code (raw bits) */
    float hi[VIS_LMAX][4];
    int stack[BVH_STACK];                         /* the warp's walk (uniform) */
    float rc[12];                                 /* the batch's boxes R and C */
};
#ifndef LB_VIS_SHAFT
#define LB_VIS_SHAFT 1          /* version-2 entry sets: 0 = bounding-box search only (A/B) */
#endif
#ifndef LB_VIS_MINBLOCKS
#define LB_VIS_MINBLOCKS 9      /* 56 registers: measured optimum on B200 (48 regs: +2 %, 40: +10 %, 72-80 uncapped: +12 %) */
#endif
template <bool ENTRY>
__global__ void __launch_bounds__(LB_BLOCK, LB_VIS_MINBLOCKS)
rad_visibility_kernel(const Bvh4Node *__restrict__ bvh, const Bvh4QNode *__restrict__ bvhq, const RayTri *__restrict__ raytris, const float4 *__restrict__ spos,
                      const uint32_t *__restrict__ sidx, const RadCand *__restrict__ cand, unsigned long long n_cand,
                      uint32_t my_k0, uint32_t my_k1, unsigned long long *__restrict__ keys, float *__restrict__ factors,
                      unsigned long long *link_count, uint4 *__restrict__ mirror, unsigned long long mirror_cap, unsigned long long *mirror_count,
                      uint32_t *chunk_cursor, unsigned long long *counters)
{
#if LB_VIS_ENTRY2
    __shared__ BvhEntrySet2 s_entry[LB_BLOCK / 32];
#if LB_VIS_PACKET
    __shared__ VisPacket s_packet[LB_BLOCK / 32];
    VisPacket &P = s_packet[threadIdx.x >> 5];
#endif
#else
    __shared__ BvhEntrySet s_entry[LB_BLOCK / 32];
#endif
#if LB_VIS_SMEMNODES
    __shared__ __align__(16) Bvh4Node s_enode[LB_BLOCK / 32][BVH_ENTRY_MAX];
    const Bvh4Node *sm_nodes = s_enode[threadIdx.x >> 5];
#else
    const Bvh4Node *sm_nodes = nullptr;
#endif
#if LB_VIS_FMA == 1
    __shared__ float s_ray[9][LB_BLOCK];
#endif
#if LB_VIS_LEAN
    __shared__ unsigned s_stat[4][LB_BLOCK];      /* per thread: segments, node visits, triangle tests, entry tests */
    s_stat[0][threadIdx.x] = s_stat[1][threadIdx.x] = s_stat[2][threadIdx.x] = s_stat[3][threadIdx.x] = 0u;
#endif
    unsigned segs = 0;
    TravStats ts = { 0, 0 };
    const unsigned lane = threadIdx.x & 31u;
#if LB_VIS_ENTRY2
    BvhEntrySet2 &E = s_entry[threadIdx.x >> 5];
#else
    BvhEntrySet &E = s_entry[threadIdx.x >> 5];
#endif
    const unsigned long long n_chunks = (n_cand + RAD_CHUNK - 1ull) / RAD_CHUNK;
    for (;;) {                                        /* persistent warps: chunks differ a lot in cost, so they are pulled from a cursor */
        uint32_t ch = 0;
        if (lane == 0) ch = atomicAdd(chunk_cursor, 1u);
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch >= n_chunks) break;
        const unsigned long long e0 = (unsigned long long)ch * RAD_CHUNK;
#if LB_VIS_ENTRY2
        if (ENTRY) {
            /* boxes of the chunk's row ends (R: the lumels of one row warp) and column ends (C) apart: the segments lie in their hull */
            float r[12];
#pragma unroll
            for (int a = 0; a < 3; ++a) { r[a] = r[6 + a] = INFINITY; r[3 + a] = r[9 + a] = -INFINITY; }
            for (unsigned i = 0; i < RAD_CHUNK; i += 128u) {          /* four independent candidate -> position load chains in flight */
                RadCand c4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long e = e0 + i + 32u * u + lane;
                    c4[u].a = RAD_PAD;
                    if (e < n_cand) c4[u] = cand[e];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool used = c4[u].a != RAD_PAD;
                    const float4 A = spos[used ? c4[u].a : 0u], B = spos[used ? c4[u].b : 0u];
                    if (used) {
                        r[0] = fminf(r[0], A.x); r[1] = fminf(r[1], A.y); r[2] = fminf(r[2], A.z); r[3] = fmaxf(r[3], A.x); r[4] = fmaxf(r[4], A.y); r[5] = fmaxf(r[5], A.z);
                        r[6] = fminf(r[6], B.x); r[7] = fminf(r[7], B.y); r[8] = fminf(r[8], B.z); r[9] = fmaxf(r[9], B.x); r[10] = fmaxf(r[10], B.y); r[11] = fmaxf(r[11], B.z);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int a = 0; a < 12; ++a) {
                    const float t = __shfl_xor_sync(0xffffffffu, r[a], o);
                    r[a] = (a % 6) < 3 ? fminf(r[a], t) : fmaxf(r[a], t);
                }
            }
            if (!(r[0] <= r[3])) continue;            /* warp-uniform: the chunk holds only unused slots */
            float lx = fminf(r[0], r[6]), ly = fminf(r[1], r[7]), lz = fminf(r[2], r[8]), hx = fmaxf(r[3], r[9]), hy = fmaxf(r[4], r[10]), hz = fmaxf(r[5], r[11]);
            const float maxabs = fmaxf(fmaxf(fmaxf(fabsf(lx), fabsf(hx)), fmaxf(fabsf(ly), fabsf(hy))), fmaxf(fabsf(lz), fabsf(hz)));
            bvh_entry_pad(lx, ly, lz, hx, hy, hz);
            __syncwarp();                             /* the previous chunk's set is no longer read */
            if (lane == 0) {
#pragma unroll
                for (int a = 0; a < 12; ++a) E.rc[a] = r[a];
            }
            __syncwarp();
            bvh4_entry_search2_warp<LB_VIS_SHAFT != 0>(bvh, lx, ly, lz, hx, hy, hz, maxabs, E, lane);
        }
#else
        if (ENTRY) {
            float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
            for (unsigned i = 0; i < RAD_CHUNK; i += 128u) {          /* four independent candidate -> position load chains in flight */
                RadCand c4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long e = e0 + i + 32u * u + lane;
                    c4[u].a = RAD_PAD;
                    if (e < n_cand) c4[u] = cand[e];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool used = c4[u].a != RAD_PAD;
                    const float4 A = spos[used ? c4[u].a : 0u], B = spos[used ? c4[u].b : 0u];
                    if (used) {
                        lx = fminf(lx, fminf(A.x, B.x)); ly = fminf(ly, fminf(A.y, B.y)); lz = fminf(lz, fminf(A.z, B.z));
                        hx = fmaxf(hx, fmaxf(A.x, B.x)); hy = fmaxf(hy, fmaxf(A.y, B.y)); hz = fmaxf(hz, fmaxf(A.z, B.z));
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
                hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
            }
            if (!(lx <= hx)) continue;                /* warp-uniform: the chunk holds only unused slots */
            bvh_entry_pad(lx, ly, lz, hx, hy, hz);
            bvh_entry_search_warp<Bvh4Access>(bvh, lx, ly, lz, hx, hy, hz, E, lane);   /* syncs the warp before it overwrites the previous chunk's set */
#if LB_VIS_SMEMNODES
            for (unsigned q = lane; q < (unsigned)E.n * 8u; q += 32u)                  /* 8 float4 per node record */
                reinterpret_cast<float4 *>(s_enode[threadIdx.x >> 5])[q] = __ldg(reinterpret_cast<const float4 *>(bvh + E.node[q >> 3]) + (q & 7u));
            __syncwarp();
#endif
        }
#endif
#if LB_VIS_LEAN
        /* Register-lean batch loop.  Across the walk only `blocked`, the batch position and the lane's statistics slots are
         * alive: the candidate record, the two lumel indices and the factor are READ AGAIN for the 4 % of segments that are
         * blocked (L1 hits), the records of later batches are brought close by address-only prefetches (two batches ahead)
         * and the next batch's record is loaded just to prefetch its lumels, then dropped.  The walk statistics live in
         * shared memory.  ptxas at 56 registers: 132 -> 0 bytes of spills in the node loop's kernel (the round-1 form kept
         * two prefetched records, the current one, both indices and four counters in registers across the walk). */
        for (unsigned i = 0; i < RAD_CHUNK; i += 32u) {
            if (e0 + i >= n_cand) break;              /* warp-uniform */
            const unsigned long long e = e0 + i + lane;
            const RadCand *cp = cand + e;
            bool blocked = false, valid = false;
            {
                if (i + 64u < RAD_CHUNK && e + 64u < n_cand) asm volatile("prefetch.global.L1 [%0];" :: "l"(cp + 64));
                if (i + 32u < RAD_CHUNK && e + 32u < n_cand) {
                    const RadCand cn = cp[32];
                    if (cn.a != RAD_PAD) {
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(spos + cn.a));
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(spos + cn.b));
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(sidx + cn.a));
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(sidx + cn.b));
                    }
                }
                RadCand c = { RAD_PAD, 0, 0.f };
                if (e < n_cand) c = *cp;
                valid = c.a != RAD_PAD;
                if (!__any_sync(0xffffffffu, valid)) continue;
#if LB_VIS_ENTRY2 && LB_VIS_PACKET
                if (ENTRY) {
                    /*
                     * Packet form.  The 32 segments of a batch come from one row warp and one or two 8-lumel column groups: a thin
                     * shaft.  The WARP walks the tree once for that shaft, from the chunk's entry set, with the bounding-box and
                     * shaft-plane tests of the entry search (bvh_entry.h), and lists the leaves in it (7.4 on average on config 4,
                     * after 5 node visits); every ray then tests the listed leaf boxes with the plain slab test -- all lanes in
                     * lock step, no stack, no node loads -- and the triangles of the boxes it enters.  A ray tests a triangle iff
                     * its slab test accepts the leaf box, exactly as on a walk from the root (accepting a leaf box implies
                     * accepting its ancestors' boxes, and a leaf dropped by the shaft lies farther from every segment of the
                     * batch than the slab test's slack: bvh_entry.h).  Lists longer than VIS_LMAX are handled in rounds.
                     */
                    V3 mA = mk3(0.f, 0.f, 0.f), dseg = mA, l1h = mA;
                    float ix = 0.f, iy = 0.f, iz = 0.f;
                    float r[12];
#pragma unroll
                    for (int a = 0; a < 3; ++a) { r[a] = r[6 + a] = INFINITY; r[3 + a] = r[9 + a] = -INFINITY; }
                    if (valid) {
                        const float4 Pa = spos[c.a], Pb = spos[c.b];          /* row end, column end */
                        const bool a_first = sidx[c.a] < sidx[c.b];            /* the reference traces from the lower lumel index */
                        const V3 A = ld3(a_first ? Pa : Pb), B = ld3(a_first ? Pb : Pa);
                        const V3 dn = norm3(B - A);
                        mA = A + dn * LB_SMALL;
                        const V3 mB = B - dn * LB_SMALL;
                        dseg = mB - mA;
                        ix = lb_slab_inv(dseg.x); iy = lb_slab_inv(dseg.y); iz = lb_slab_inv(dseg.z);
                        l1h = mk3(lb_slab_origin_hi(mA.x, dseg.x), lb_slab_origin_hi(mA.y, dseg.y), lb_slab_origin_hi(mA.z, dseg.z));
                        r[0] = r[3] = Pa.x; r[1] = r[4] = Pa.y; r[2] = r[5] = Pa.z; r[6] = r[9] = Pb.x; r[7] = r[10] = Pb.y; r[8] = r[11] = Pb.z;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int a = 0; a < 12; ++a) {
                            const float t = __shfl_xor_sync(0xffffffffu, r[a], o);
                            r[a] = (a % 6) < 3 ? fminf(r[a], t) : fmaxf(r[a], t);
                        }
                    }
                    float qlx = fminf(r[0], r[6]), qly = fminf(r[1], r[7]), qlz = fminf(r[2], r[8]), qhx = fmaxf(r[3], r[9]), qhy = fmaxf(r[4], r[10]), qhz = fmaxf(r[5], r[11]);
                    const float maxabs = fmaxf(fmaxf(fmaxf(fabsf(qlx), fabsf(qhx)), fmaxf(fabsf(qly), fabsf(qhy))), fmaxf(fabsf(qlz), fabsf(qhz)));
                    bvh_entry_pad(qlx, qly, qlz, qhx, qhy, qhz);
                    __syncwarp();
                    if (lane < 12u) P.rc[lane] = r[lane];
                    __syncwarp();
                    float p_nu0, p_nv0, p_ru0, p_rv0, p_lim0, p_nu1 = 0.f, p_nv1 = 0.f, p_ru1 = 0.f, p_rv1 = 0.f, p_lim1 = 0.f;
                    bvh_shaft_plane(P.rc, P.rc + 6, maxabs, (int)(lane >> 2), p_nu0, p_nv0, p_ru0, p_rv0, p_lim0);
                    if (lane < 16u) bvh_shaft_plane(P.rc, P.rc + 6, maxabs, 8 + (int)(lane >> 2), p_nu1, p_nv1, p_ru1, p_rv1, p_lim1);
                    const int pa0 = (int)(lane >> 4);
                    const unsigned cc = lane & 3u;
                    int sp = 0, nl = 0;
                    unsigned visits = 0, ntri = 0, nbox = 0;
                    /* one step of the warp's walk: box `cc` of a node (or of a group of four entries), seen by the eight lanes cc, cc + 4, .. */
#define VIS_CONSIDER(HAS, CODE, LX, LY, LZ, HX, HY, HZ)                                                                                  \
                    {                                                                                                                    \
                        bool out_;                                                                                                       \
                        {                                                                                                                \
                            const float lu = pa0 == 0 ? (LX) : (LY), hu = pa0 == 0 ? (HX) : (HY), lv = pa0 == 0 ? (LY) : (LZ), hv = pa0 == 0 ? (HY) : (HZ); \
                            const float bu = p_nu0 > 0.f ? lu : hu, bv = p_nv0 > 0.f ? lv : hv;                                           \
                            out_ = p_nu0 * (bu - p_ru0) + p_nv0 * (bv - p_rv0) > p_lim0;                                                 \
                        }                                                                                                                \
                        {                                                                                                                \
                            const float bu = p_nu1 > 0.f ? (LZ) : (HZ), bv = p_nv1 > 0.f ? (LX) : (HX);                                   \
                            out_ = out_ || (p_nu1 * (bu - p_ru1) + p_nv1 * (bv - p_rv1) > p_lim1);                                       \
                        }                                                                                                                \
                        const unsigned om_ = LB_VIS_SHAFT ? __ballot_sync(0xffffffffu, out_) : 0u;                                       \
                        const bool in_ = lane < 4u && (HAS) && (CODE) != BVH4_EMPTY && (LX) <= qhx && (HX) >= qlx && (LY) <= qhy && (HY) >= qly && \
                                         (LZ) <= qhz && (HZ) >= qlz && !(om_ & (0x11111111u << cc));                                     \
                        const unsigned hm_ = __ballot_sync(0xffffffffu, in_), lm_ = __ballot_sync(0xffffffffu, in_ && (CODE) < 0), im_ = hm_ & ~lm_; \
                        if (in_) {                                                                                                       \
                            if ((CODE) < 0) {                                                                                            \
                                const int at_ = nl + __popc(lm_ & ((1u << lane) - 1u));                                                  \
                                *reinterpret_cast<float4 *>(P.lo[at_]) = make_float4((LX), (LY), (LZ), __int_as_float(CODE));            \
                                *reinterpret_cast<float4 *>(P.hi[at_]) = make_float4((HX), (HY), (HZ), 0.f);                             \
                            } else P.stack[sp + __popc(im_ & ((1u << lane) - 1u))] = (CODE);                                             \
                        }                                                                                                                \
                        nl += __popc(lm_); sp += __popc(im_);                                                                            \
                        __syncwarp();                                                                                                    \
                    }
                    for (int e4 = 0; e4 < E.n; e4 += 4) {                  /* seed: the chunk's entries, four at a time */
                        const int ei = e4 + (int)cc;
                        const bool has = ei < E.n;
                        const float4 lo = *reinterpret_cast<const float4 *>(E.lo[has ? ei : 0]), hi = *reinterpret_cast<const float4 *>(E.hi[has ? ei : 0]);
                        const int code = __float_as_int(lo.w);
                        VIS_CONSIDER(has, code, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)
                    }
                    bool more;
                    do {
                        while (sp > 0 && nl <= VIS_LMAX - 4) {
                            const int node = P.stack[sp - 1];
                            __syncwarp();                             /* every lane has read the top before a push overwrites it */
                            --sp;
                            const Bvh4Node &N = bvh[node];
                            const int code = __ldg(&N.c[cc]);
                            const float lx = __ldg(&N.lox[cc]), ly = __ldg(&N.loy[cc]), lz = __ldg(&N.loz[cc]), hx = __ldg(&N.hix[cc]), hy = __ldg(&N.hiy[cc]), hz = __ldg(&N.hiz[cc]);
                            ++visits;
                            VIS_CONSIDER(true, code, lx, ly, lz, hx, hy, hz)
                        }
                        more = sp > 0;
                        if (valid && !blocked && nl) {                    /* the rays' turn: nl listed leaf boxes, the same words on every lane */
                            unsigned m = 0;
                            for (int j = 0; j < nl; ++j) {
                                const float4 lo = *reinterpret_cast<const float4 *>(P.lo[j]), hi = *reinterpret_cast<const float4 *>(P.hi[j]);
                                const float x0 = (lo.x - mA.x) * ix, x1 = (hi.x - l1h.x) * ix, y0 = (lo.y - mA.y) * iy, y1 = (hi.y - l1h.y) * iy;
                                const float z0 = (lo.z - mA.z) * iz, z1 = (hi.z - l1h.z) * iz;
                                const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                                const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), 1.f));
                                if (t0 <= t1 + 2e-6f) m |= 1u << j;
                            }
                            nbox += (unsigned)nl;
                            while (m && !blocked) {
                                const int j = __ffs(m) - 1;
                                m &= m - 1u;
                                const unsigned code = ~(unsigned)__float_as_int(P.lo[j][3]);
                                const RayTri *tp = raytris + (code >> 3);
                                for (unsigned t = code & 7u; t; --t, ++tp) {
                                    RayTri T;
                                    load_raytri(tp, T);
                                    ++ntri;
                                    if (seg_tri_prepared<true>(mA, dseg, T) < 1.0f) blocked = true;     /* scene-BVH triangles are "useful" by construction */
                                }
                            }
                        }
                        __syncwarp();                                 /* the list is rewritten in the next round */
                        nl = 0;
                    } while (more);
#undef VIS_CONSIDER
                    if (valid) { s_stat[0][threadIdx.x] += 1u; s_stat[2][threadIdx.x] += ntri; s_stat[3][threadIdx.x] += nbox; }
                    if (lane == 0) s_stat[1][threadIdx.x] += 2u * visits;
                } else
#endif
                if (valid) {                              /* RAD_PAD: unused slot of a warp's output chunk */
                    const bool a_first = sidx[c.a] < sidx[c.b];         /* the reference traces from the lower lumel index */
                    const V3 A = ld3(spos[a_first ? c.a : c.b]), B = ld3(spos[a_first ? c.b : c.a]);
                    const V3 dn = norm3(B - A);
                    const V3 mA = A + dn * LB_SMALL, mB = B - dn * LB_SMALL;
                    TravStats t1 = { 0, 0, 0 };
#if LB_VIS_ENTRY2
                    blocked = ENTRY ? bvh4_anyhit_entries2<LB_VIS_FLUSH>(bvh, raytris, E, mA, mB, t1) : bvh4_anyhit<LB_VIS_FLUSH>(bvh, raytris, mA, mB, t1);
#elif LB_VIS_Q8
                    blocked = ENTRY ? bvh4q_anyhit_entries<LB_VIS_FLUSH>(bvhq, raytris, E, mA, mB, t1) : bvh4_anyhit<LB_VIS_FLUSH>(bvh, raytris, mA, mB, t1);
#elif LB_VIS_FMA == 2
                    blocked = ENTRY ? bvh4_anyhit_entries_f2<LB_VIS_FLUSH>(bvh, raytris, E, mA, mB, t1) : bvh4_anyhit<LB_VIS_FLUSH>(bvh, raytris, mA, mB, t1);
#elif LB_VIS_FMA
                    blocked = ENTRY ? bvh4_anyhit_entries_f<LB_VIS_FLUSH>(bvh, raytris, E, mA, mB, &s_ray[0][threadIdx.x], t1) : bvh4_anyhit<LB_VIS_FLUSH>(bvh, raytris, mA, mB, t1);
#else
                    blocked = ENTRY ? bvh4_anyhit_entries<LB_VIS_FLUSH>(bvh, raytris, E, mA, mB, t1, sm_nodes) : bvh4_anyhit<LB_VIS_FLUSH>(bvh, raytris, mA, mB, t1);
#endif
                    s_stat[0][threadIdx.x] += 1u; s_stat[1][threadIdx.x] += t1.nodes; s_stat[2][threadIdx.x] += t1.tris; s_stat[3][threadIdx.x] += t1.entries;
                }
            }
            const unsigned bm = __ballot_sync(0xffffffffu, blocked);
            if (!bm) continue;                        /* 96 % of the segments are free: most batches end here */
            unsigned emit = 0;                        /* bit 0: link for row a (always mine), bit 1: row b is mine too, bit 2: row b lives on another rank */
            RadCand c = { 0, 0, 0.f };
            uint32_t oa = 0, ob = 0;
            if (blocked) {
                asm volatile("" : "+l"(cp));          /* the record is read again (it was not kept across the walk) */
                c = *cp;
                oa = sidx[c.a]; ob = sidx[c.b];
                emit = 1u | ((c.b >= my_k0 && c.b < my_k1) ? 2u : 4u);
            }
            const unsigned cnt = __popc(emit & 3u);
            unsigned pre = cnt;
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= (unsigned)o) pre += t; }
            const unsigned total = __shfl_sync(0xffffffffu, pre, 31);
            {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(link_count, (unsigned long long)total);
                base = __shfl_sync(0xffffffffu, base, 0);
                unsigned long long at = base + (pre - cnt);
                if (emit & 1u) { keys[at] = ((unsigned long long)c.a << 32) | ob; factors[at] = c.factor; ++at; }
                if (emit & 2u) { keys[at] = ((unsigned long long)c.b << 32) | oa; factors[at] = c.factor; }
            }
            const unsigned mm = __ballot_sync(0xffffffffu, (emit & 4u) != 0);
            if (mm) {                                 /* mirrored link for a row owned by another rank: queued for the exchange */
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(mirror_count, (unsigned long long)__popc(mm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (emit & 4u) {
                    const unsigned long long at = base + __popc(mm & ((1u << lane) - 1u));
                    if (at < mirror_cap) mirror[at] = make_uint4(c.b, oa, __float_as_uint(c.factor), 0u);
                }
            }
        }
    }
    segs = s_stat[0][threadIdx.x]; ts.nodes = s_stat[1][threadIdx.x]; ts.tris = s_stat[2][threadIdx.x]; ts.entries = s_stat[3][threadIdx.x];
#else
        /* software pipeline over the chunk's 32 batches: the candidate records are loaded two batches ahead, the lumels of
         * the next batch's candidates are prefetched into L1 while this batch is traced (ncu: a fifth of the kernel's stall
         * samples sat on the candidate -> position load chain) */
        RadCand cn1 = { RAD_PAD, 0, 0.f }, cn2 = { RAD_PAD, 0, 0.f };
        if (e0 + lane < n_cand) cn1 = cand[e0 + lane];
        if (e0 + 32u + lane < n_cand) cn2 = cand[e0 + 32u + lane];
        for (unsigned i = 0; i < RAD_CHUNK; i += 32u) {
            if (e0 + i >= n_cand) break;              /* warp-uniform */
            unsigned emit = 0;                        /* bit 0: link for row a (always mine), bit 1: row b is mine too, bit 2: row b lives on another rank */
            const RadCand c = cn1;
            uint32_t oa = 0, ob = 0;
            cn1 = cn2;
            {
                const unsigned long long e2 = e0 + i + 64u + lane;
                cn2.a = RAD_PAD;
                if (i + 64u < RAD_CHUNK && e2 < n_cand) cn2 = cand[e2];
            }
            if (cn1.a != RAD_PAD) {
                asm volatile("prefetch.global.L1 [%0];" :: "l"(spos + cn1.a));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(spos + cn1.b));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(sidx + cn1.a));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(sidx + cn1.b));
            }
            if (!__any_sync(0xffffffffu, c.a != RAD_PAD)) continue;
            if (c.a != RAD_PAD) {                     /* RAD_PAD: unused slot of a warp's output chunk */
                oa = sidx[c.a]; ob = sidx[c.b];
                const bool a_first = oa < ob;         /* the reference traces from the lower lumel index */
                const V3 A = ld3(spos[a_first ? c.a : c.b]), B = ld3(spos[a_first ? c.b : c.a]);
                const V3 dn = norm3(B - A);
                const V3 mA = A + dn * LB_SMALL, mB = B - dn * LB_SMALL;
                ++segs;
#if LB_VIS_FMA
                const bool blocked = ENTRY ? bvh4_anyhit_entries_f<LB_VIS_FLUSH>(bvh, raytris, E, mA, mB, &s_ray[0][threadIdx.x], ts) : bvh4_anyhit<LB_VIS_FLUSH>(bvh, raytris, mA, mB, ts);
#else
                const bool blocked = ENTRY ? bvh4_anyhit_entries<LB_VIS_FLUSH>(bvh, raytris, E, mA, mB, ts) : bvh4_anyhit<LB_VIS_FLUSH>(bvh, raytris, mA, mB, ts);
#endif
                if (blocked) emit = 1u | ((c.b >= my_k0 && c.b < my_k1) ? 2u : 4u);
            }
            const unsigned cnt = __popc(emit & 3u);
            /* warp-aggregated append: exclusive prefix of cnt over the warp */
            unsigned pre = cnt;
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= (unsigned)o) pre += t; }
            const unsigned total = __shfl_sync(0xffffffffu, pre, 31);
            if (total) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(link_count, (unsigned long long)total);
                base = __shfl_sync(0xffffffffu, base, 0);
                unsigned long long at = base + (pre - cnt);
                if (emit & 1u) { keys[at] = ((unsigned long long)c.a << 32) | ob; factors[at] = c.factor; ++at; }
                if (emit & 2u) { keys[at] = ((unsigned long long)c.b << 32) | oa; factors[at] = c.factor; }
            }
            /* mirrored link for a row owned by another rank: queued for the exchange */
            const unsigned mm = __ballot_sync(0xffffffffu, (emit & 4u) != 0);
            if (mm) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(mirror_count, (unsigned long long)__popc(mm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (emit & 4u) {
                    const unsigned long long at = base + __popc(mm & ((1u << lane) - 1u));
                    if (at < mirror_cap) mirror[at] = make_uint4(c.b, oa, __float_as_uint(c.factor), 0u);
                }
            }
        }
    }
#endif
    count_add(counters, CNT_RAD_SEGMENTS, segs);
    count_add(counters, CNT_RAY_NODE_VISITS, ts.nodes);
    count_add(counters, CNT_RAY_TRI_TESTS, ts.tris);
    count_add(counters, CNT_RAY_ENTRY_TESTS, ts.entries);
}

/* after the exchange: keep the mirrored links whose row is mine */
__global__ void rad_mirror_filter_kernel(const uint4 *__restrict__ all, unsigned long long stride, const unsigned long long *__restrict__ counts,
                                         uint32_t world, uint32_t me, uint32_t my_k0, uint32_t my_k1,
                                         unsigned long long *__restrict__ keys, float *__restrict__ factors, unsigned long long *link_count)
{
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t q = 0; q < world; ++q) {
        if (q == me) continue;
        const unsigned long long nq = counts[q], n_pad = (nq + 31ull) & ~31ull;
        for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_pad; e += (unsigned long long)gridDim.x * blockDim.x) {
            uint4 r = make_uint4(0, 0, 0, 0);
            bool keep = false;
            if (e < nq) { r = all[(unsigned long long)q * stride + e]; keep = r.x >= my_k0 && r.x < my_k1; }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (m) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(link_count, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) {
                    const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
                    keys[at] = ((unsigned long long)r.x << 32) | r.y;
                    factors[at] = __uint_as_float(r.z);
                }
            }
        }
    }
}

/* personalised exchange of the mirrored links (ctx->alltoallv): records are bucketed by the rank that owns their row
 * (rows are dealt in contiguous runs of rows_per_rank sorted positions), so that a record crosses NVLink once, to one
 * rank, instead of reaching all of them in an all-gather (8 ranks on config 4: 2.1 GB into every rank -> 0.26 GB). */
__global__ void rad_mirror_count_kernel(const uint4 *__restrict__ rec, unsigned long long n, uint32_t rows_per_rank, uint32_t world, unsigned long long *__restrict__ counts)
{
    __shared__ unsigned s_cnt[64];
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x)
        atomicAdd(&s_cnt[rec[e].x / rows_per_rank], 1u);
    __syncthreads();
    if (threadIdx.x < world && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void rad_mirror_bucket_kernel(const uint4 *__restrict__ rec, unsigned long long n, uint32_t rows_per_rank, const unsigned long long *__restrict__ bucket_off,
                                         unsigned long long *__restrict__ cursor, uint4 *__restrict__ out)
{
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < ((n + 31ull) & ~31ull); e += (unsigned long long)gridDim.x * blockDim.x) {
        const bool valid = e < n;
        const uint4 r = valid ? rec[e] : make_uint4(0, 0, 0, 0);
        const unsigned owner = valid ? r.x / rows_per_rank : 0xffffffffu;
        /* lanes of the same owner reserve their slots with one atomic */
        const unsigned peers = __match_any_sync(0xffffffffu, owner);
        const unsigned lane = threadIdx.x & 31u, leader = (unsigned)__ffs(peers) - 1u;
        unsigned long long base = 0;
        if (valid && lane == leader) base = atomicAdd(cursor + owner, (unsigned long long)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) out[bucket_off[owner] + base + __popc(peers & ((1u << lane) - 1u))] = r;
    }
}

__global__ void rad_mirror_append_kernel(const uint4 *__restrict__ rec, unsigned long long n, unsigned long long *__restrict__ keys, float *__restrict__ factors)
{
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 r = rec[e];
        keys[e] = ((unsigned long long)r.x << 32) | r.y;
        factors[e] = __uint_as_float(r.z);
    }
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

/* material / energy state per SORTED row: diffuse(+area), total, out */
__global__ void rad_init_kernel(const float4 *__restrict__ lrgb, const float4 *__restrict__ lrad, const float *__restrict__ diffuse3,
                                const float *__restrict__ emissive3, const uint32_t *__restrict__ sidx, uint32_t n_probes, uint64_t n,
                                uint64_t k0, uint64_t k1, float4 *__restrict__ diff, float4 *__restrict__ total, float4 *__restrict__ out)
{
    uint64_t k = k0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const uint32_t i = sidx[k];
    if (i == RAD_PAD) { diff[k] = total[k] = out[k] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const bool probe = i < n_probes;
    V3 d = probe ? mk3(0.f) : mk3(1.f);
    V3 e = ld3(lrgb[i]);
    if (!probe && diffuse3) {
        d = mk3(diffuse3[(size_t)i * 3], diffuse3[(size_t)i * 3 + 1], diffuse3[(size_t)i * 3 + 2]);
        e = e + mk3(emissive3[(size_t)i * 3], emissive3[(size_t)i * 3 + 1], emissive3[(size_t)i * 3 + 2]);
    }
    const float area = probe ? 0.f : lrad[i].w;
    diff[k] = make_float4(d.x, d.y, d.z, area);
    total[k] = make_float4(e.x, e.y, e.z, 0.f);
    out[k] = make_float4(e.x, e.y, e.z, 0.f);
}

/* E = (out * diffuse) * area for my sorted rows, the quantity that crosses a link (and NVLink) */
__global__ void rad_energy_kernel(const float4 *__restrict__ diff, const float4 *__restrict__ out, uint64_t k0, uint64_t k1, float4 *__restrict__ Es)
{
    uint64_t k = k0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const float4 d = diff[k], o = out[k];
    V3 e = (ld3(o) * ld3(d)) * d.w;
    Es[k] = make_float4(e.x, e.y, e.z, 0.f);
}

/* sorted layout -> original lumel order (links name partners by original index) */
__global__ void rad_unpermute_kernel(const float4 *__restrict__ src_sorted, const uint32_t *__restrict__ sidx, uint64_t n, float4 *__restrict__ dst_orig)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && sidx[k] != RAD_PAD) dst_orig[sidx[k]] = src_sorted[k];
}

__global__ void rad_bounce_kernel(const uint64_t *__restrict__ rowoff, const uint32_t *__restrict__ other, const float *__restrict__ factor,
                                  const float4 *__restrict__ Eo, uint64_t k0, uint64_t k1, float4 *__restrict__ total, float4 *__restrict__ out)
{
    uint64_t k = k0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const uint64_t a = rowoff[k - k0], b = rowoff[k - k0 + 1];
    V3 t = ld3(total[k]);
    V3 in = mk3(0.f);
    for (uint64_t e = a; e < b; ++e) {
        const V3 c = ld3(__ldg(Eo + other[e])) * factor[e];
        t = t + c;
        in = in + c;
    }
    total[k] = make_float4(t.x, t.y, t.z, 0.f);
    out[k] = make_float4(in.x, in.y, in.z, 0.f);
}

template <class T> static int grow_buf(ltrgpu_Ctx *ctx, T **p, size_t *cap, size_t used, size_t need)
{
    if (need <= *cap) return 0;
    size_t ncap = *cap ? *cap : (size_t)1 << 20;
    while (ncap < need) ncap *= 2;
    T *q = nullptr;
    CU_TRY(ctx, lb_malloc((void **)&q, ncap * sizeof(T)));
    if (*p && used) CU_TRY(ctx, cudaMemcpyAsync(q, *p, used * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (*p) lb_free(*p);
    *p = q; *cap = ncap;
    return 0;
}

static int rad_allgather(ltrgpu_Ctx *ctx, float4 *buf, uint64_t chunk_elems, const char *what)
{
    if (ctx->world <= 1) return 0;
    if (!ctx->allgather || ctx->allgather(ctx->allgather_user, buf + chunk_elems * ctx->rank, buf, chunk_elems * sizeof(float4), ctx->stream)) {
        snprintf(ctx->err, sizeof(ctx->err), "radiosity: all-gather of %s failed", what);
        return 1;
    }
    return 0;
}

struct RadHostMaterials { const float *diffuse3, *emissive3; };
static int rad_host_materials(void *user, const float **d, const float **e)
{
    const RadHostMaterials *m = (const RadHostMaterials *)user;
    *d = m->diffuse3; *e = m->emissive3;
    return 0;
}

extern "C" int ltrgpu_radiosity(ltrgpu_Ctx *ctx, const float *diffuse3, const float *emissive3, int bounces)
{
    RadHostMaterials m = { diffuse3, emissive3 };
    return ltrgpu_radiosity_ex(ctx, rad_host_materials, &m, bounces);
}

extern "C" int ltrgpu_radiosity_ex(ltrgpu_Ctx *ctx, ltrgpu_materials_fn materials, void *materials_user, int bounces)
{
    const float *diffuse3 = nullptr, *emissive3 = nullptr;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n = ctx->n_lumels;
    if (n == 0 || bounces <= 0) return 0;
    if (n > 0x7ffffff0ull) { snprintf(ctx->err, sizeof(ctx->err), "too many lumels for radiosity"); return 1; }
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, st));

    /* tiles, padded so every rank owns the same number of whole tiles */
    const uint32_t world = (uint32_t)(ctx->world > 0 ? ctx->world : 1);
    const uint32_t tiles_raw = (uint32_t)((n + RAD_TILE - 1) / RAD_TILE);
    const uint32_t tiles_per_rank = (tiles_raw + world - 1) / world;
    const uint32_t n_tiles = tiles_per_rank * world;
    const uint64_t n_pad = (uint64_t)n_tiles * RAD_TILE;
    const uint32_t my_t0 = tiles_per_rank * (uint32_t)ctx->rank, my_t1 = my_t0 + tiles_per_rank;
    const uint64_t k0 = (uint64_t)my_t0 * RAD_TILE, k1 = (uint64_t)my_t1 * RAD_TILE, n_rows = k1 - k0;

    float4 *spos = nullptr, *snrm = nullptr, *diff = nullptr, *total = nullptr, *out = nullptr, *Es = nullptr, *Eo = nullptr, *lrgb_full = nullptr;
    TileBounds *tb = nullptr, *tb32 = nullptr, *tbg = nullptr, *tbs = nullptr;
    uint32_t *mkeys = nullptr, *mkeys_alt = nullptr, *sidx = nullptr, *sidx_alt = nullptr;
    float *d_bounds = nullptr, *d_diffuse = nullptr, *d_emissive = nullptr;
    RadCand *cand = nullptr;
    unsigned long long *d_cnt = nullptr;           /* [0] candidates, [1] links, [2] mirrored links */
    uint4 *mirror = nullptr, *mirror_all = nullptr;
    uint32_t *item_cnt = nullptr, *item_off = nullptr;
    uint2 *items = nullptr;
    unsigned long long *d_mcounts = nullptr;
    size_t mirror_cap = 0, mirror_used = 0;
    unsigned long long *keys = nullptr, *keys_alt = nullptr;
    float *fac = nullptr, *fac_alt = nullptr;
    size_t link_cap = 0, link_cap_f = 0, link_used = 0;
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    float ms_pairs = 0, ms_vis = 0;
    int rc = 1;

#define RAD_TRY(x) do { if (x) goto done; } while (0)
#define RAD_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto done; } } while (0)
#define RAD_LAUNCHED() do { ctx->host_counters.kernel_launches++; RAD_CU(cudaGetLastError()); } while (0)
    /* LTR_TRACE=1: host-side phase timings on stderr (adds a stream sync per phase) */
    const bool trace = getenv("LTR_TRACE") != nullptr;
    /* column group size of the warp-level culling (4, 8, 16 or 32 lumels); LTR_RAD_GROUP overrides for experiments */
    int group = RAD_GROUP_DEFAULT;
    bool rad_entry = true;                         /* visibility rays start at their chunk's entry set; LTR_RAD_ENTRY=0: at the root (A/B) */
    if (const char *e = getenv("LTR_RAD_ENTRY")) rad_entry = atoi(e) != 0;
    if (const char *e = getenv("LTR_RAD_GROUP")) { const int g = atoi(e); if (g == 4 || g == 8 || g == 16 || g == 32) group = g; }
    struct timespec tr0; clock_gettime(CLOCK_MONOTONIC, &tr0);
#define RAD_TRACE(label) do { if (trace) { cudaStreamSynchronize(st); struct timespec t_; clock_gettime(CLOCK_MONOTONIC, &t_); \
        fprintf(stderr, "[ltr rank %d] radiosity %-22s %8.2f ms\n", ctx->rank, label, (t_.tv_sec - tr0.tv_sec) * 1e3 + (t_.tv_nsec - tr0.tv_nsec) * 1e-6); tr0 = t_; } } while (0)

    {
        /* the emitted light of EVERY lumel is needed: the direct stage has left the colours of all lumels on every rank
         * (its own exchange, gpu_direct.cu) */
        lrgb_full = ctx->d_lrgb;

        /* ---- 1. Morton sort ---- */
        RAD_TRY(dev_alloc(ctx, &mkeys, n)); RAD_TRY(dev_alloc(ctx, &mkeys_alt, n));
        RAD_TRY(dev_alloc(ctx, &sidx, n_pad)); RAD_TRY(dev_alloc(ctx, &sidx_alt, n));
        RAD_TRY(dev_alloc(ctx, &d_bounds, 6));
        {
            const float init[6] = { INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY };
            RAD_CU(cudaMemcpyAsync(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
            rad_bounds_reduce_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(ctx->d_lpos, n, d_bounds);
            RAD_LAUNCHED();
            float hb[6];
            RAD_CU(cudaMemcpyAsync(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
            RAD_CU(cudaStreamSynchronize(st));
            float3 lo = make_float3(hb[0], hb[1], hb[2]);
            float3 inv = make_float3(hb[3] > hb[0] ? 1.0f / (hb[3] - hb[0]) : 0.f, hb[4] > hb[1] ? 1.0f / (hb[4] - hb[1]) : 0.f, hb[5] > hb[2] ? 1.0f / (hb[5] - hb[2]) : 0.f);
            rad_morton_kernel<<<grid_for(n, 256), 256, 0, st>>>(ctx->d_lpos, n, lo, inv, mkeys, sidx);
            RAD_LAUNCHED();
            cub::DoubleBuffer<uint32_t> kb(mkeys, mkeys_alt), vb(sidx, sidx_alt);
            RAD_CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp_bytes, kb, vb, (int)n, 0, 30, st));
            RAD_CU(lb_malloc(&sort_tmp, sort_tmp_bytes ? sort_tmp_bytes : 16));
            RAD_CU(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_tmp_bytes, kb, vb, (int)n, 0, 30, st));
            ctx->host_counters.kernel_launches += 4;
            if (vb.Current() != sidx) RAD_CU(cudaMemcpyAsync(sidx, vb.Current(), n * 4, cudaMemcpyDeviceToDevice, st));
            RAD_CU(cudaStreamSynchronize(st));
            lb_free(sort_tmp); sort_tmp = nullptr; sort_tmp_bytes = 0;
        }
        {
            uint32_t *sdeal = nullptr;
            RAD_TRY(dev_alloc(ctx, &sdeal, n_pad));
            rad_deal_kernel<<<grid_for(n_pad, 256), 256, 0, st>>>(sidx, n, n_pad, world, tiles_per_rank, sdeal);
            RAD_LAUNCHED();
            lb_free(sidx);
            sidx = sdeal;
        }
        RAD_TRACE("morton sort");
        RAD_TRY(dev_alloc(ctx, &spos, n_pad)); RAD_TRY(dev_alloc(ctx, &snrm, n_pad));
        RAD_TRY(dev_alloc(ctx, &tb, n_tiles)); RAD_TRY(dev_alloc(ctx, &tb32, (size_t)n_tiles * (RAD_TILE / 32)));
        if (group != 32) RAD_TRY(dev_alloc(ctx, &tbg, (size_t)n_tiles * (RAD_TILE / group)));
        rad_gather_kernel<<<grid_for(n_pad, 256), 256, 0, st>>>(ctx->d_lpos, ctx->d_lnrm, sidx, n, n_pad, spos, snrm);
        RAD_LAUNCHED();
        switch (group) {
        case 4:  rad_tile_bounds_kernel<4><<<n_tiles, RAD_TILE, 0, st>>>(spos, snrm, sidx, tb, tb32, tbg); break;
        case 8:  rad_tile_bounds_kernel<8><<<n_tiles, RAD_TILE, 0, st>>>(spos, snrm, sidx, tb, tb32, tbg); break;
        case 16: rad_tile_bounds_kernel<16><<<n_tiles, RAD_TILE, 0, st>>>(spos, snrm, sidx, tb, tb32, tbg); break;
        default: rad_tile_bounds_kernel<32><<<n_tiles, RAD_TILE, 0, st>>>(spos, snrm, sidx, tb, tb32, tb32); break;
        }
        RAD_LAUNCHED();
        const uint32_t n_super = (n_tiles + 31u) / 32u;
        RAD_TRY(dev_alloc(ctx, &tbs, n_super));
        rad_super_bounds_kernel<<<grid_for((uint64_t)n_super * 32, 128), 128, 0, st>>>(tb, n_tiles, n_super, world, tiles_per_rank, tbs);
        RAD_LAUNCHED();
        /* pair sweep launch shape: persistent warps, as many CTAs as fit (shared memory bound: two staged column tiles per warp) */
        size_t sweep_smem = 0;
        int sweep_ctas = 1;
        {
            const void *fn = nullptr;
            switch (group) {
            case 4:  fn = (const void *)rad_candidates_kernel<4>;  sweep_smem = rad_sweep_smem<4>();  break;
            case 8:  fn = (const void *)rad_candidates_kernel<8>;  sweep_smem = rad_sweep_smem<8>();  break;
            case 16: fn = (const void *)rad_candidates_kernel<16>; sweep_smem = rad_sweep_smem<16>(); break;
            default: fn = (const void *)rad_candidates_kernel<32>; sweep_smem = rad_sweep_smem<32>(); break;
            }
            RAD_CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem));
            RAD_CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            RAD_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sweep_ctas, fn, RAD_WARPS * 32, sweep_smem));
            if (sweep_ctas < 1) sweep_ctas = 1;
        }

        /* ---- 2a. work items: (row warp, super-tile) pairs surviving the interval test, row-warp-major ---- */
        const uint32_t my_rw0 = my_t0 * (RAD_TILE / 32), my_row_warps = tiles_per_rank * (RAD_TILE / 32);
        uint32_t n_items = 0;
        {
            RAD_TRY(dev_alloc(ctx, &item_cnt, (size_t)my_row_warps + 1)); RAD_TRY(dev_alloc(ctx, &item_off, (size_t)my_row_warps + 1));
            RAD_CU(cudaMemsetAsync(item_cnt + my_row_warps, 0, 4, st));
            rad_items_kernel<false><<<grid_for((uint64_t)my_row_warps * 32, 128), 128, 0, st>>>(tb32, tbs, n_super, world, tiles_per_rank, my_rw0, my_row_warps, item_cnt, nullptr, nullptr);
            RAD_LAUNCHED();
            RAD_CU(cub::DeviceScan::ExclusiveSum(nullptr, sort_tmp_bytes, item_cnt, item_off, (int)my_row_warps + 1, st));
            RAD_CU(lb_malloc(&sort_tmp, sort_tmp_bytes ? sort_tmp_bytes : 16));
            RAD_CU(cub::DeviceScan::ExclusiveSum(sort_tmp, sort_tmp_bytes, item_cnt, item_off, (int)my_row_warps + 1, st));
            ctx->host_counters.kernel_launches += 2;
            RAD_CU(cudaMemcpyAsync(&n_items, item_off + my_row_warps, 4, cudaMemcpyDeviceToHost, st));
            RAD_CU(cudaStreamSynchronize(st));
            lb_free(sort_tmp); sort_tmp = nullptr; sort_tmp_bytes = 0;
            RAD_TRY(dev_alloc(ctx, &items, (size_t)n_items));
            rad_items_kernel<true><<<grid_for((uint64_t)my_row_warps * 32, 128), 128, 0, st>>>(tb32, tbs, n_super, world, tiles_per_rank, my_rw0, my_row_warps, nullptr, item_off, items);
            RAD_LAUNCHED();
            dev_free(&item_cnt); dev_free(&item_off);
        }
        RAD_TRACE("tile bounds + work items");

        /* ---- 2b-4. candidates and visibility, in batches of work items bounded by the candidate buffer ---- */
        RAD_TRY(dev_alloc(ctx, &d_cnt, 4));
        /* candidate buffer: a sixteenth of the device's TOTAL memory, between 16 Mi and 1 Gi records (12 B each).  Not of the
         * FREE memory: blocks cached by lb_malloc count as used, so a size derived from it drifted from bake to bake, missed
         * the cache every time it crossed a size class and paid a 12 GB cudaMalloc (~42 ms, seen as bake-time outliers). */
        unsigned long long cand_cap = 16ull << 20;
        {
            /* asked ONCE per device and process: cudaMemGetInfo takes the driver's memory lock -- 0.5 ms usually, but 13 / 49 / 79 ms
             * in 3 of 28 resident bakes of one measurement (profiles/r02_ab_runs.md, "step jitter"), all of it on the bake's span */
            static size_t total_of_device[64];
            size_t free_b = 0, total_b = (ctx->device >= 0 && ctx->device < 64) ? total_of_device[ctx->device] : 0;
            if (total_b || cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                if (ctx->device >= 0 && ctx->device < 64) total_of_device[ctx->device] = total_b;
                unsigned long long want = (unsigned long long)(total_b / 16 / sizeof(RadCand));
                if (want > cand_cap) cand_cap = want;
                if (cand_cap > (1024ull << 20)) cand_cap = 1024ull << 20;
            }
        }
        if (const char *e = getenv("LTR_RAD_CAND_CAP")) { const unsigned long long v = strtoull(e, nullptr, 10); if (v >= 4096) cand_cap = v; }   /* tests: force many batches */
        RAD_TRY(dev_alloc(ctx, &cand, cand_cap));
        /* batch of items: probe with a small one, then size each batch from the measured yield */
        uint32_t batch = (uint32_t)ctx->num_sms * (uint32_t)sweep_ctas * RAD_WARPS * 8u;
        RAD_TRACE("candidate buffer");
        for (uint32_t i0 = 0; i0 < n_items;) {
            const uint32_t i1 = batch < n_items - i0 ? i0 + batch : n_items;
            RAD_CU(cudaMemsetAsync(d_cnt, 0, 32, st));
            RAD_CU(cudaEventRecord(ctx->ev_k0, st));
            {
                const uint32_t resident = (uint32_t)ctx->num_sms * (uint32_t)sweep_ctas;
                const uint32_t want = (i1 - i0 + RAD_WARPS - 1) / RAD_WARPS;
                const uint32_t grid = want < resident ? want : resident;
#define RAD_SWEEP(GS, TBG) rad_candidates_kernel<GS><<<grid, RAD_WARPS * 32, sweep_smem, st>>>(spos, snrm, tb, tb32, TBG, tbs, n_tiles, world, tiles_per_rank, \
                                                                 items + i0, i1 - i0, (uint32_t *)(d_cnt + 3), cand, cand_cap, d_cnt, ctx->d_counters)
                switch (group) {
                case 4:  RAD_SWEEP(4, tbg); break;
                case 8:  RAD_SWEEP(8, tbg); break;
                case 16: RAD_SWEEP(16, tbg); break;
                default: RAD_SWEEP(32, tb32); break;
                }
#undef RAD_SWEEP
            }
            RAD_LAUNCHED();
            RAD_CU(cudaEventRecord(ctx->ev_k1, st));
            unsigned long long h_cnt[4];
            RAD_CU(cudaMemcpyAsync(h_cnt, d_cnt, 32, cudaMemcpyDeviceToHost, st));
            RAD_CU(cudaStreamSynchronize(st));
            { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1); ms_pairs += ms; }
            if (h_cnt[0] > cand_cap) {
                if (i1 - i0 == 1) { snprintf(ctx->err, sizeof(ctx->err), "radiosity: one work item produced %llu candidates (> %llu)", h_cnt[0], cand_cap); goto done; }
                /* redo this batch smaller: scale by the overshoot with 30 % headroom */
                uint32_t nb = (uint32_t)((double)(i1 - i0) * (double)cand_cap / (double)h_cnt[0] * 0.7);
                batch = nb ? nb : 1;
                continue;
            }
            const unsigned long long nc = h_cnt[0];
            if (nc) {
                struct timespec tg0, tg1;
                if (trace) clock_gettime(CLOCK_MONOTONIC, &tg0);
                RAD_TRY(grow_buf(ctx, &keys, &link_cap, link_used, link_used + 2 * nc));
                RAD_TRY(grow_buf(ctx, &fac, &link_cap_f, link_used, link_used + 2 * nc));
                if (world > 1) RAD_TRY(grow_buf(ctx, &mirror, &mirror_cap, mirror_used, mirror_used + nc));
                if (trace) {
                    clock_gettime(CLOCK_MONOTONIC, &tg1);
                    fprintf(stderr, "[ltr rank %d] radiosity batch %u items, %llu candidates, link buffers grown in %.2f ms (capacity %zu)\n", ctx->rank, i1 - i0, nc,
                            (tg1.tv_sec - tg0.tv_sec) * 1e3 + (tg1.tv_nsec - tg0.tv_nsec) * 1e-6, link_cap);
                }
                /* one warp per RAD_CHUNK candidates at a time, pulled from a cursor (the upper half of d_cnt[3], zero since the memset above) */
                unsigned long long want = ((nc + RAD_CHUNK - 1) / RAD_CHUNK + LB_BLOCK / 32 - 1) / (LB_BLOCK / 32);
                unsigned cap = (unsigned)ctx->num_sms * 16;
                unsigned blocks = want > cap ? cap : (unsigned)want;
                RAD_CU(cudaEventRecord(ctx->ev_k0, st));
#define RAD_VIS(ENTRY) rad_visibility_kernel<ENTRY><<<blocks, LB_BLOCK, 0, st>>>(ctx->d_bvh4, ctx->d_bvh4q, ctx->d_raytris, spos, sidx, cand, nc, (uint32_t)k0, (uint32_t)k1, \
                                                                  keys + link_used, fac + link_used, d_cnt + 1,                                                \
                                                                  mirror ? mirror + mirror_used : nullptr, mirror ? mirror_cap - mirror_used : 0, d_cnt + 2,  \
                                                                  (uint32_t *)(d_cnt + 3) + 1, ctx->d_counters)
                if (rad_entry) RAD_VIS(true); else RAD_VIS(false);
#undef RAD_VIS
                RAD_LAUNCHED();
                RAD_CU(cudaEventRecord(ctx->ev_k1, st));
                RAD_CU(cudaMemcpyAsync(h_cnt, d_cnt, 32, cudaMemcpyDeviceToHost, st));
                RAD_CU(cudaStreamSynchronize(st));
                { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1); ms_vis += ms; }
                link_used += h_cnt[1];
                mirror_used += h_cnt[2];
                ctx->host_counters.rad_batches++;
            }
            {
                const double per_item = (double)(nc ? nc : 1) / (double)(i1 - i0);
                double nb = 0.7 * (double)cand_cap / per_item;
                if (nb < 1.0) nb = 1.0;
                if (nb > 4.0e9) nb = 4.0e9;
                batch = (uint32_t)nb;
            }
            i0 = i1;
        }
        dev_free(&items);
        dev_free(&cand);
        RAD_TRACE("candidates+visibility");
        if (trace) {
            unsigned long long c[CNT_COUNT];
            cudaMemcpy(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[ltr rank %d] radiosity pairs %llu segments %llu links(local) %zu  pairs-kernel %.1f ms  visibility-kernel %.1f ms\n", ctx->rank,
                    c[CNT_RAD_PAIRS], c[CNT_RAD_SEGMENTS], link_used, ms_pairs, ms_vis);
        }

        /* ---- 4b. exchange the mirrored links: each rank's records are all-gathered (padded to the largest
         *          count), every rank keeps the ones whose row it owns.  ~16 B per cross-rank link. ---- */
        /* From 4 ranks up (with 2 an all-gather moves the same bytes).  NCCL opens its point-to-point connections at the first
         * send/recv of a communicator: seconds, once per process (8 ranks: 6.6 s measured) -- a caller that bakes once on many
         * GPUs can keep the all-gather with LTR_RAD_MIRROR_ALLGATHER=1. */
        const char *env_a2a = getenv("LTR_RAD_MIRROR_ALLGATHER");
        if (world >= 4 && world <= 64 && ctx->alltoallv && !(env_a2a && env_a2a[0] == '1')) {
            /* counts[r] = my records whose row rank r owns; the world x world matrix of them is all-gathered (tiny) */
            unsigned long long *d_cur = nullptr, *d_boff = nullptr, *d_matrix = nullptr;
            uint4 *bucketed = nullptr, *incoming = nullptr;
            RAD_TRY(dev_alloc(ctx, &d_matrix, (size_t)world * world)); RAD_TRY(dev_alloc(ctx, &d_cur, world)); RAD_TRY(dev_alloc(ctx, &d_boff, world + 1));
            unsigned long long *my_row = d_matrix + (size_t)ctx->rank * world;
            RAD_CU(cudaMemsetAsync(my_row, 0, 8 * world, st));
            RAD_CU(cudaMemsetAsync(d_cur, 0, 8 * world, st));
            if (mirror_used) {
                rad_mirror_count_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(mirror, mirror_used, (uint32_t)n_rows, world, my_row);
                RAD_LAUNCHED();
            }
            if (!ctx->allgather || ctx->allgather(ctx->allgather_user, my_row, d_matrix, 8 * (size_t)world, st)) {
                snprintf(ctx->err, sizeof(ctx->err), "radiosity: all-gather of the mirrored-link counts failed"); goto done;
            }
            std::vector<unsigned long long> M((size_t)world * world);
            RAD_CU(cudaMemcpyAsync(M.data(), d_matrix, 8 * M.size(), cudaMemcpyDeviceToHost, st));
            RAD_CU(cudaStreamSynchronize(st));
            std::vector<uint64_t> soff(world + 1, 0), roff(world + 1, 0);
            std::vector<unsigned long long> boff(world + 1, 0);
            for (uint32_t r = 0; r < world; ++r) {
                const unsigned long long to_r = r == (uint32_t)ctx->rank ? 0 : M[(size_t)ctx->rank * world + r], from_r = r == (uint32_t)ctx->rank ? 0 : M[(size_t)r * world + ctx->rank];
                boff[r + 1] = boff[r] + M[(size_t)ctx->rank * world + r];
                soff[r] = boff[r] * sizeof(uint4);
                roff[r + 1] = roff[r] + from_r * sizeof(uint4);
                (void)to_r;
            }
            soff[world] = boff[world] * sizeof(uint4);
            /* send ranges are the buckets themselves (the bucket of my own rank is empty: my rows never go to the mirror list) */
            const unsigned long long n_in = roff[world] / sizeof(uint4);
            RAD_TRY(dev_alloc(ctx, &bucketed, mirror_used ? mirror_used : 1)); RAD_TRY(dev_alloc(ctx, &incoming, n_in ? n_in : 1));
            RAD_CU(cudaMemcpyAsync(d_boff, boff.data(), 8 * (world + 1), cudaMemcpyHostToDevice, st));
            if (mirror_used) {
                rad_mirror_bucket_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(mirror, mirror_used, (uint32_t)n_rows, d_boff, d_cur, bucketed);
                RAD_LAUNCHED();
            }
            if (ctx->alltoallv(ctx->allgather_user, bucketed, soff.data(), incoming, roff.data(), st)) {
                snprintf(ctx->err, sizeof(ctx->err), "radiosity: exchange of the mirrored links failed"); goto done;
            }
            RAD_TRY(grow_buf(ctx, &keys, &link_cap, link_used, link_used + n_in));
            RAD_TRY(grow_buf(ctx, &fac, &link_cap_f, link_used, link_used + n_in));
            if (n_in) {
                rad_mirror_append_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(incoming, n_in, keys + link_used, fac + link_used);
                RAD_LAUNCHED();
            }
            RAD_CU(cudaStreamSynchronize(st));
            link_used += n_in;
            dev_free(&d_matrix); dev_free(&d_cur); dev_free(&d_boff); dev_free(&bucketed); dev_free(&incoming); dev_free(&mirror);
            RAD_TRACE("mirror link exchange (personalised)");
        } else if (world > 1) {
            RAD_TRY(dev_alloc(ctx, &d_mcounts, world));
            RAD_CU(cudaMemsetAsync(d_mcounts, 0, 8 * world, st));
            unsigned long long mine_cnt = mirror_used;
            RAD_CU(cudaMemcpyAsync(d_mcounts + ctx->rank, &mine_cnt, 8, cudaMemcpyHostToDevice, st));
            if (!ctx->allgather || ctx->allgather(ctx->allgather_user, d_mcounts + ctx->rank, d_mcounts, 8, st)) {
                snprintf(ctx->err, sizeof(ctx->err), "radiosity: all-gather of link counts failed"); goto done;
            }
            std::vector<unsigned long long> hc(world);
            RAD_CU(cudaMemcpyAsync(hc.data(), d_mcounts, 8 * world, cudaMemcpyDeviceToHost, st));
            RAD_CU(cudaStreamSynchronize(st));
            unsigned long long stride = 1, incoming = 0;
            for (uint32_t q = 0; q < world; ++q) { if (hc[q] > stride) stride = hc[q]; if (q != (uint32_t)ctx->rank) incoming += hc[q]; }
            RAD_TRY(dev_alloc(ctx, &mirror_all, (size_t)stride * world));
            if (mirror_used) RAD_CU(cudaMemcpyAsync(mirror_all + (size_t)stride * ctx->rank, mirror, mirror_used * sizeof(uint4), cudaMemcpyDeviceToDevice, st));
            if (ctx->allgather(ctx->allgather_user, mirror_all + (size_t)stride * ctx->rank, mirror_all, (size_t)stride * sizeof(uint4), st)) {
                snprintf(ctx->err, sizeof(ctx->err), "radiosity: all-gather of mirrored links failed"); goto done;
            }
            RAD_TRY(grow_buf(ctx, &keys, &link_cap, link_used, link_used + incoming));
            RAD_TRY(grow_buf(ctx, &fac, &link_cap_f, link_used, link_used + incoming));
            RAD_CU(cudaMemsetAsync(d_cnt, 0, 32, st));
            rad_mirror_filter_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(mirror_all, stride, d_mcounts, world, (uint32_t)ctx->rank, (uint32_t)k0, (uint32_t)k1,
                                                                     keys + link_used, fac + link_used, d_cnt + 1);
            RAD_LAUNCHED();
            unsigned long long got[4];
            RAD_CU(cudaMemcpyAsync(got, d_cnt, 32, cudaMemcpyDeviceToHost, st));
            RAD_CU(cudaStreamSynchronize(st));
            link_used += got[1];
            dev_free(&mirror_all); dev_free(&mirror);
            RAD_TRACE("mirror link exchange");
        }

        /* ---- 5. sort links by (row, partner) -> CSR in reference accumulation order ---- */
        if (link_used) {
            RAD_TRY(dev_alloc(ctx, &keys_alt, link_used)); RAD_TRY(dev_alloc(ctx, &fac_alt, link_used));
            cub::DoubleBuffer<unsigned long long> kb(keys, keys_alt);
            cub::DoubleBuffer<float> vb(fac, fac_alt);
            int end_bit = 33;                             /* key = row position << 32 | partner index: the bits above the row field are zero */
            while (end_bit < 64 && (n_pad >> (end_bit - 32)) != 0) ++end_bit;
            RAD_CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp_bytes, kb, vb, (int64_t)link_used, 0, end_bit, st));
            RAD_CU(lb_malloc(&sort_tmp, sort_tmp_bytes ? sort_tmp_bytes : 16));
            RAD_CU(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_tmp_bytes, kb, vb, (int64_t)link_used, 0, end_bit, st));
            ctx->host_counters.kernel_launches += 8;
            RAD_CU(cudaStreamSynchronize(st));
            if (kb.Current() != keys) { unsigned long long *t = keys; keys = keys_alt; keys_alt = t; }
            if (vb.Current() != fac) { float *t = fac; fac = fac_alt; fac_alt = t; }
            dev_free(&keys_alt); dev_free(&fac_alt);
        }
        dev_free(&ctx->d_rad_rowoff); dev_free(&ctx->d_rad_other); dev_free(&ctx->d_rad_factor); dev_free(&ctx->d_rad_sidx);
        RAD_TRY(dev_alloc(ctx, &ctx->d_rad_rowoff, n_rows + 1));
        RAD_TRY(dev_alloc(ctx, &ctx->d_rad_other, link_used));
        rad_row_offsets_kernel<<<grid_for(n_rows + 1, 256), 256, 0, st>>>(keys, link_used, k0, n_rows, ctx->d_rad_rowoff);
        RAD_LAUNCHED();
        if (link_used) {
            rad_split_keys_kernel<<<grid_for(link_used, 256), 256, 0, st>>>(keys, link_used, ctx->d_rad_other);
            RAD_LAUNCHED();
        }
        dev_free(&keys);
        ctx->d_rad_factor = fac; fac = nullptr;
        ctx->d_rad_sidx = sidx; sidx = nullptr;
        ctx->rad_rows = n_rows; ctx->rad_links = link_used; ctx->rad_k0 = k0; ctx->rad_n = n;
        {
            unsigned long long lc = link_used;
            RAD_CU(cudaMemcpyAsync(ctx->d_counters + CNT_RAD_LINKS, &lc, 8, cudaMemcpyHostToDevice, st));
            RAD_CU(cudaStreamSynchronize(st));
        }

        RAD_TRACE("link sort + CSR");
        /* ---- bounces ---- */
        RAD_TRY(dev_alloc(ctx, &diff, n_pad)); RAD_TRY(dev_alloc(ctx, &total, n_pad)); RAD_TRY(dev_alloc(ctx, &out, n_pad));
        RAD_TRY(dev_alloc(ctx, &Es, n_pad)); RAD_TRY(dev_alloc(ctx, &Eo, n + LB_PAD));
        if (materials && materials(materials_user, &diffuse3, &emissive3)) { snprintf(ctx->err, sizeof(ctx->err), "material callbacks failed"); rc = 1; goto done; }
        if (diffuse3) { RAD_TRY(dev_upload(ctx, &d_diffuse, diffuse3, n * 3)); RAD_TRY(dev_upload(ctx, &d_emissive, emissive3, n * 3)); }
        rad_init_kernel<<<grid_for(n_rows, 256), 256, 0, st>>>(lrgb_full, ctx->d_lrad, d_diffuse, d_emissive, ctx->d_rad_sidx, ctx->n_probes, n, k0, k1, diff, total, out);
        RAD_LAUNCHED();
        for (int b = 0; b < bounces; ++b) {
            rad_energy_kernel<<<grid_for(n_rows, 256), 256, 0, st>>>(diff, out, k0, k1, Es);
            RAD_LAUNCHED();
            RAD_TRY(rad_allgather(ctx, Es, (uint64_t)tiles_per_rank * RAD_TILE, "bounce energy"));
            rad_unpermute_kernel<<<grid_for(n_pad, 256), 256, 0, st>>>(Es, ctx->d_rad_sidx, n_pad, Eo);
            RAD_LAUNCHED();
            rad_bounce_kernel<<<grid_for(n_rows, 128), 128, 0, st>>>(ctx->d_rad_rowoff, ctx->d_rad_other, ctx->d_rad_factor, Eo, k0, k1, total, out);
            RAD_LAUNCHED();
        }
        RAD_TRACE("bounces");
        /* commit: total light back to original lumel order, for every lumel on every rank */
        RAD_TRY(rad_allgather(ctx, total, (uint64_t)tiles_per_rank * RAD_TILE, "total light"));
        rad_unpermute_kernel<<<grid_for(n_pad, 256), 256, 0, st>>>(total, ctx->d_rad_sidx, n_pad, ctx->d_lrgb);
        RAD_LAUNCHED();
        RAD_CU(cudaEventRecord(ctx->ev1, st));
        RAD_CU(cudaStreamSynchronize(st));
        RAD_TRACE("commit");
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->host_counters.ms_radiosity += ms;
        ctx->host_counters.ms_rad_pairs += ms_pairs;
        ctx->host_counters.ms_rad_vis += ms_vis;
        rc = 0;
    }
done:
    lb_free(spos); lb_free(snrm); lb_free(diff); lb_free(total); lb_free(out); lb_free(Es); lb_free(Eo); lb_free(tb); lb_free(tb32); lb_free(tbg); lb_free(tbs); lb_free(item_cnt); lb_free(item_off); lb_free(items);
    lb_free(mkeys); lb_free(mkeys_alt); lb_free(sidx); lb_free(sidx_alt); lb_free(d_bounds);
    lb_free(d_diffuse); lb_free(d_emissive); lb_free(cand); lb_free(d_cnt); lb_free(keys); lb_free(keys_alt);
    lb_free(fac); lb_free(fac_alt); lb_free(sort_tmp); lb_free(mirror); lb_free(mirror_all); lb_free(d_mcounts);
    return rc;
#undef RAD_TRY
#undef RAD_CU
#undef RAD_LAUNCHED
#undef RAD_TRACE
}

/* debug dump: rows are in SORTED (Morton) order on the device; hand them back in original lumel order */
extern "C" int ltrgpu_download_links(ltrgpu_Ctx *ctx, uint64_t *row_offset, uint32_t *other, float *factor, uint64_t *rows, uint64_t *count)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    *rows = ctx->world > 1 ? ctx->rad_rows : ctx->rad_n;
    *count = ctx->rad_links;
    if (!ctx->d_rad_rowoff || !row_offset) return 0;
    const uint64_t nr = ctx->rad_rows, nl = ctx->rad_links, n = ctx->rad_n;
    uint64_t *ro = (uint64_t *)malloc((nr + 1) * 8);
    uint32_t *oth = (uint32_t *)malloc((nl ? nl : 1) * 4), *sidx = (uint32_t *)malloc((n ? n : 1) * 4);
    float *fa = (float *)malloc((nl ? nl : 1) * 4);
    cudaError_t e = cudaMemcpyAsync(ro, ctx->d_rad_rowoff, (nr + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && nl) e = cudaMemcpyAsync(oth, ctx->d_rad_other, nl * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && nl) e = cudaMemcpyAsync(fa, ctx->d_rad_factor, nl * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sidx, ctx->d_rad_sidx, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    int rc = 0;
    if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "download_links: %s", cudaGetErrorString(e)); rc = 1; }
    else if (ctx->world > 1) {                      /* sharded: rows stay in this rank's sorted order */
        memcpy(row_offset, ro, (nr + 1) * 8);
        if (nl) { memcpy(other, oth, nl * 4); memcpy(factor, fa, nl * 4); }
    } else {
        /* single GPU: permute rows back to original lumel order (row k of the device CSR is lumel sidx[k]) */
        uint64_t *len = (uint64_t *)calloc(n + 1, 8);
        for (uint64_t k = 0; k < n && k < nr; ++k) len[sidx[k]] = ro[k + 1] - ro[k];
        uint64_t acc = 0;
        for (uint64_t i = 0; i < n; ++i) { row_offset[i] = acc; acc += len[i]; }
        row_offset[n] = acc;
        for (uint64_t k = 0; k < n && k < nr; ++k) {
            const uint64_t dst = row_offset[sidx[k]];
            memcpy(other + dst, oth + ro[k], (ro[k + 1] - ro[k]) * 4);
            memcpy(factor + dst, fa + ro[k], (ro[k + 1] - ro[k]) * 4);
        }
        free(len);
    }
    free(ro); free(oth); free(sidx); free(fa);
    return rc;
}
