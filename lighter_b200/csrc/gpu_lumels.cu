/*
 * gpu_lumels.cu -- lumel generation on the GPU.
 *
 * Reference behaviour restated (file:line into /root/reference):
 *   raster    lighter.cpp:32-66, lighter_math.cpp:245-302  two passes (margin 0.501 then 0) over the
 *             lightmap-UV triangles in part/index order, last writer wins, barycentrics clamped
 *             individually to [0,1]
 *   compact   lighter.cpp:424-443,467-475  lumel <=> texel whose interpolated normal != 0, row-major
 *   offset    lighter_math.cpp:991-1044    concave-edge push-out against every instance, in order
 *   overlap   lighter.cpp:449-465          <=100 closest-hit probes along the normal
 *
 * GPU formulation: "last writer wins" becomes an atomicMax over a (pass, triangle-order) key per
 * texel, one warp per (pass, triangle) sweeping the triangle's texel box; the winner's attributes
 * are then recomputed per texel, flags are scanned to give the reference's row-major lumel order,
 * and one thread per lumel runs the two sequential corrections.
 */
#include "gpu_internal.cuh"

#include <stdlib.h>

struct F2 { float x, y; };
__device__ __forceinline__ F2 f2(float x, float y) { F2 v; v.x = x; v.y = y; return v; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return f2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float dot2(F2 a, F2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float cross2(F2 a, F2 b) { return (a.x * b.y) - (a.y * b.x); }
__device__ __forceinline__ F2 perp2(F2 v) { return f2(v.y, -v.x); }
__device__ __forceinline__ F2 norm2(F2 v)
{
    float l2 = v.x * v.x + v.y * v.y;
    if (l2 == 0) return f2(0, 0);
    float inv = 1.0f / sqrtf(l2);
    return f2(v.x * inv, v.y * inv);
}
__device__ __forceinline__ F2 ldf2(const float2 *p) { float2 v = __ldg(p); return f2(v.x, v.y); }

__device__ __forceinline__ float heron(float a, float b, float c)
{
    float p = (a + b + c) * 0.5f;
    float q = p * (p - a) * (p - b) * (p - c);
    return q < 0 ? 0.f : sqrtf(q);
}

struct TexelSpace {
    const ltrgpu_Inst *inst;
    const ltrgpu_RasterTri *rtris;
    const V3 *wpos, *wnrm;
    const float2 *vtex, *ltex;
};

__device__ __forceinline__ uint64_t texel_off(const ltrgpu_Inst &I) { return ((uint64_t)I.texel_off_hi << 32) | I.texel_off_lo; }

/* ---------------------------------------------------------------------------------------------
 * pass kernel: one warp per triangle, atomicMax of the ordered key over covered texels
 * ------------------------------------------------------------------------------------------- */
__global__ void raster_kernel(TexelSpace ts, uint32_t n_rtris, float margin, uint32_t passbit, uint32_t *__restrict__ texkey)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rtris; r += warps_per_grid) {
        const ltrgpu_RasterTri T = ts.rtris[r];
        const ltrgpu_Inst I = ts.inst[T.inst];
        const int W = (int)I.lm_w, H = (int)I.lm_h;
        F2 p1 = ldf2(ts.ltex + T.i0), p2 = ldf2(ts.ltex + T.i1), p3 = ldf2(ts.ltex + T.i2);
        int maxX = (int)(fmaxr(p1.x, fmaxr(p2.x, p3.x)) + margin);
        int minX = (int)(fminr(p1.x, fminr(p2.x, p3.x)) - margin);
        int maxY = (int)(fmaxr(p1.y, fmaxr(p2.y, p3.y)) + margin);
        int minY = (int)(fminr(p1.y, fminr(p2.y, p3.y)) - margin);
        if (maxX < 0 || minX >= W || maxY < 0 || minY >= H) continue;
        if (minX < 0) minX = 0;
        if (maxX >= W) maxX = W - 1;
        if (minY < 0) minY = 0;
        if (maxY >= H) maxY = H - 1;
        F2 p1p2 = p2 - p1, p1p3 = p3 - p1;
        if (cross2(p1p2, p1p3) == 0) continue;
        F2 n1 = norm2(perp2(p1p2));
        F2 n2 = norm2(perp2(p3 - p2));
        F2 n3 = norm2(perp2(p1p3));
        n3 = f2(-n3.x, -n3.y);
        float d1 = dot2(n1, p1), d2 = dot2(n2, p2), d3 = dot2(n3, p3);
        const float MG = margin;
        const uint32_t key = passbit | (r + 1u);
        const uint64_t base = texel_off(I);
        const int bw = maxX - minX + 1, bh = maxY - minY + 1;
        const int total = bw * bh;
        for (int k = (int)lane; k < total; k += 32) {
            int y = minY + k / bw, x = minX + k % bw;
            F2 p = f2((float)x, (float)y);
            float pd1 = dot2(p, n1), pd2 = dot2(p, n2), pd3 = dot2(p, n3);
            bool in = (pd1 <= d1 + MG && pd2 <= d2 + MG && pd3 <= d3 + MG) || (pd1 + MG >= d1 && pd2 + MG >= d2 && pd3 + MG >= d3);
            if (in) atomicMax(texkey + base + (uint64_t)x + (uint64_t)W * y, key);
        }
    }
}

/* attributes of texel g as written by its winning triangle */
__device__ __forceinline__ bool resolve_texel(const TexelSpace &ts, uint64_t g, uint32_t key, bool full,
                                              V3 &P, V3 &N, float4 &XD, uint32_t &inst_id, uint32_t &loc)
{
    if (key == 0) return false;
    const ltrgpu_RasterTri T = ts.rtris[(key & 0x7fffffffu) - 1u];
    const ltrgpu_Inst I = ts.inst[T.inst];
    uint64_t li = g - texel_off(I);
    int x = (int)(li % I.lm_w), y = (int)(li / I.lm_w);
    F2 p1 = ldf2(ts.ltex + T.i0), p2 = ldf2(ts.ltex + T.i1), p3 = ldf2(ts.ltex + T.i2);
    F2 p1p2 = p2 - p1, p1p3 = p3 - p1;
    float vcross = cross2(p1p2, p1p3);
    F2 q = f2((float)x - p1.x, (float)y - p1.y);
    float s = cross2(q, p1p3) / vcross;
    float t = cross2(p1p2, q) / vcross;
    s = fmaxr(fminr(s, 1.0f), 0.0f);
    t = fmaxr(fminr(t, 1.0f), 0.0f);
    V3 b1 = ts.wnrm[T.i0], b2 = ts.wnrm[T.i1], b3 = ts.wnrm[T.i2];
    N = b1 + (b2 - b1) * s + (b3 - b1) * t;
    inst_id = T.inst;
    loc = (uint32_t)li;
    if (!full) return !is_zero3(N);
    V3 a1 = ts.wpos[T.i0], a2 = ts.wpos[T.i1], a3 = ts.wpos[T.i2];
    P = a1 + (a2 - a1) * s + (a3 - a1) * t;
    /* world area per texel area of this triangle (ref: lighter_math.cpp:150-181) */
    float lmarea, coarea;
    {
        F2 e1 = p2 - p1, e2 = p3 - p2, e3 = p1 - p3;
        lmarea = heron(sqrtf(e1.x * e1.x + e1.y * e1.y), sqrtf(e2.x * e2.x + e2.y * e2.y), sqrtf(e3.x * e3.x + e3.y * e3.y));
        coarea = heron(len3(a2 - a1), len3(a3 - a2), len3(a1 - a3));
    }
    float area = lmarea > 0 ? coarea / lmarea : 0.f;
    F2 c1 = ldf2(ts.vtex + T.i0), c2 = ldf2(ts.vtex + T.i1), c3 = ldf2(ts.vtex + T.i2);
    float part = (float)T.part;
    XD.x = c1.x + (c2.x - c1.x) * s + (c3.x - c1.x) * t;
    XD.y = c1.y + (c2.y - c1.y) * s + (c3.y - c1.y) * t;
    XD.z = part + (part - part) * s + (part - part) * t;
    XD.w = area + (area - area) * s + (area - area) * t;
    return !is_zero3(N);
}

__global__ void lumel_flags_kernel(TexelSpace ts, uint64_t n_texels, const uint32_t *__restrict__ texkey, uint32_t *__restrict__ flags)
{
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_texels) return;
    V3 P, N; float4 XD; uint32_t inst, loc;
    flags[g] = resolve_texel(ts, g, texkey[g], false, P, N, XD, inst, loc) ? 1u : 0u;
}

/* ---------------------------------------------------------------------------------------------
 * exclusive scan of u32 (three small kernels; chunk = 2048 values per CTA)
 * ------------------------------------------------------------------------------------------- */
#define SCAN_T 256
#define SCAN_V 8
#define SCAN_CHUNK (SCAN_T * SCAN_V)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[SCAN_T / 32];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += n; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < SCAN_T / 32 ? warp_sums[lane] : 0;
        uint32_t winc = w;
        for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= (unsigned)o) winc += n; }
        if (lane < SCAN_T / 32) warp_sums[lane] = winc - w;
        if (lane == SCAN_T / 32 - 1) *total = winc;
    }
    __syncthreads();
    uint32_t r = inc - v + warp_sums[wid];
    __syncthreads();
    return r;
}

__global__ void scan_reduce_kernel(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ block_sums)
{
    __shared__ uint32_t total;
    uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * SCAN_V;
    uint32_t s = 0;
    for (int k = 0; k < SCAN_V; ++k) if (base + k < n) s += in[base + k];
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void scan_sums_kernel(uint32_t *block_sums, uint32_t nblocks, uint32_t *grand_total)
{
    __shared__ uint32_t total;
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += SCAN_T) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? block_sums[i] : 0;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (i < nblocks) block_sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void scan_apply_kernel(uint32_t *__restrict__ data, uint64_t n, const uint32_t *__restrict__ block_sums)
{
    __shared__ uint32_t total;
    uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * SCAN_V;
    uint32_t v[SCAN_V], s = 0;
    for (int k = 0; k < SCAN_V; ++k) { v[k] = base + k < n ? data[base + k] : 0; s += v[k]; }
    uint32_t ex = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
    for (int k = 0; k < SCAN_V; ++k) { if (base + k < n) data[base + k] = ex; ex += v[k]; }
}

__global__ void gather_inst_offsets_kernel(const ltrgpu_Inst *inst, uint32_t n_inst, uint64_t n_texels, const uint32_t *scan,
                                           const uint32_t *total, uint32_t n_probes, uint64_t *out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_inst) return;
    if (i == 0) { out[0] = 0; return; }
    uint64_t off = i < n_inst ? texel_off(inst[i]) : n_texels;
    out[i] = (uint64_t)n_probes + (off < n_texels ? scan[off] : *total);
}

/* ---------------------------------------------------------------------------------------------
 * emit: write the compacted lumel records
 * ------------------------------------------------------------------------------------------- */
__global__ void lumel_emit_kernel(TexelSpace ts, uint64_t n_texels, const uint32_t *__restrict__ texkey, const uint32_t *__restrict__ scan,
                                  uint32_t n_probes, float4 *lpos, float4 *lnrm, float4 *lrad, uint32_t *lloc, uint32_t *linst)
{
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_texels) return;
    V3 P, N; float4 XD; uint32_t inst, loc;
    if (!resolve_texel(ts, g, texkey[g], true, P, N, XD, inst, loc)) return;
    uint64_t i = (uint64_t)n_probes + scan[g];
    N = norm3(N);
    lpos[i] = make_float4(P.x, P.y, P.z, 0.f);
    lnrm[i] = make_float4(N.x, N.y, N.z, 0.f);
    lrad[i] = XD;
    lloc[i] = loc;
    linst[i] = inst;
}

__global__ void probe_emit_kernel(const V3 *pos, const V3 *nrm, uint32_t n, float4 *lpos, float4 *lnrm, float4 *lrad, uint32_t *lloc, uint32_t *linst)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lpos[i] = make_float4(pos[i].x, pos[i].y, pos[i].z, 0.f);
    lnrm[i] = make_float4(nrm[i].x, nrm[i].y, nrm[i].z, 0.f);
    lrad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    lloc[i] = 0;
    linst[i] = 0;
}

/* ---------------------------------------------------------------------------------------------
 * reference-order tree walks (used only here; see reftree.h for why order matters)
 * ------------------------------------------------------------------------------------------- */
struct RefView { const RefNode *nodes; const int32_t *items; const float *tris9; const PreparedTri *ptris; const float4 *boxes; };

__device__ __forceinline__ void ld_tri9(const float *t, V3 &a, V3 &b, V3 &c)
{
    a = mk3(t[0], t[1], t[2]); b = mk3(t[3], t[4], t[5]); c = mk3(t[6], t[7], t[8]);
}

/* does the segment's own box overlap the triangle box widened by 1e-4 relative + absolute?  (necessary for a hit) */
__device__ __forceinline__ bool seg_box_overlap(V3 a, V3 b, float4 lo, float4 hi)
{
    const float e = 1e-4f;
    const float ex = e * (1.0f + fmaxf(fabsf(lo.x), fabsf(hi.x))), ey = e * (1.0f + fmaxf(fabsf(lo.y), fabsf(hi.y))), ez = e * (1.0f + fmaxf(fabsf(lo.z), fabsf(hi.z)));
    return !(fminf(a.x, b.x) > hi.x + ex || fmaxf(a.x, b.x) < lo.x - ex || fminf(a.y, b.y) > hi.y + ey || fmaxf(a.y, b.y) < lo.y - ey ||
             fminf(a.z, b.z) > hi.z + ez || fmaxf(a.z, b.z) < lo.z - ez);
}

/* closest hit inside one instance tree, first-met wins ties (ref: lighter_math.cpp:835-871) */
__device__ float reftree_closest(const RefView &v, V3 from, V3 to, int *tid)
{
    RefRay r = make_ref_ray(from, to);
    float closest = 2.0f;
    int hitid = -1;
    int stack[24];
    int sp = 0;
    stack[sp++] = 0;
    /* The reference's ray/box test accepts boxes BEHIND the origin (lighter_math.cpp:618-650 has no tmax >= 0 check), so its
     * walk follows the whole line through the instance; a triangle can only be hit inside the segment's own box, and a node
     * box contains every triangle below it, so nodes that miss that box (widened like the per-triangle pre-test) are skipped:
     * same hits, same first-met order among them.  On a 500 k-triangle instance this is most of the walk. */
    const float sbe = 1e-4f;
    const V3 slo = min3(from, to), shi = max3(from, to);
    while (sp) {
        int node = stack[--sp];
        RefNode N = v.nodes[node];
        if (!ref_ray_box(r, N.lo, N.hi)) continue;
        if (slo.x > N.hi.x + sbe * (1.0f + fabsf(N.hi.x)) || shi.x < N.lo.x - sbe * (1.0f + fabsf(N.lo.x)) ||
            slo.y > N.hi.y + sbe * (1.0f + fabsf(N.hi.y)) || shi.y < N.lo.y - sbe * (1.0f + fabsf(N.lo.y)) ||
            slo.z > N.hi.z + sbe * (1.0f + fabsf(N.hi.z)) || shi.z < N.lo.z - sbe * (1.0f + fabsf(N.lo.z))) continue;
        if (N.ido != -1) {
            int cnt = v.items[N.ido];
            for (int k = 0; k < cnt; ++k) {
                int id = v.items[N.ido + 1 + k];
                {   /* conservative pre-test: a segment that misses the triangle's box (widened) cannot hit the triangle.
                     * Degenerate reference trees put thousands of triangles in one leaf (a flat 40x40 ceiling ends up
                     * as ONE leaf of 7848), and the reference tests them all. */
                    const float4 bl = __ldg(v.boxes + 2ull * id), bh = __ldg(v.boxes + 2ull * id + 1);
                    if (!seg_box_overlap(from, to, bl, bh)) continue;
                }
                V3 a, b, c;
                ld_tri9(v.tris9 + 9ull * id, a, b, c);
                float d = seg_tri(from, to, a, b, c);
                if (d < closest) { closest = d; hitid = id; }
            }
        }
        if (N.ch != -1 && sp + 2 <= 24) { stack[sp++] = N.ch; stack[sp++] = node + 1; }
    }
    if (hitid != -1) *tid = hitid;
    return closest;
}

/* The tests a triangle must pass before it may move the sample (ref: SampleOffsetQuery, lighter_math.cpp:991-1038), up to but
 * not including the closest-hit probe; on success the target position and the slide direction. */
__device__ __forceinline__ bool offset_tri_tests(const PreparedTri &PT, V3 P, V3 N, float dist, V3 &Pnew, V3 &projTPN)
{
    float ndst = point_tri_distance_prepared(P, PT);
    if (!(ndst < dist)) return false;
    V3 TPN = -tri_back_normal(PT.t0, PT.t1, PT.t2);
    float TPD = dot3(TPN, PT.t0);
    float sigdst = dot3(TPN, P) - TPD;
    if (!(sigdst <= LB_SMALL)) return false;
    V3 Pextr = P + norm3(TPN + N) * -sigdst * 1.41f;
    if (!point_proj_on_tri(Pextr, PT)) return false;
    float projDot = dot3(TPN, N);
    if (!(fabsf(projDot) < 0.95f)) return false;
    projTPN = norm3(TPN - N * projDot);
    float dotFactor = 1.0f - fabsf(projDot);
    Pnew = P + projTPN * (-sigdst / dotFactor + LB_SMALL);
    return true;
}

/* One candidate triangle of the concave-edge offset, evaluated with the CURRENT (possibly already moved)
 * sample position -- the reference mutates P while it walks (lighter_math.cpp:991-1038). */
__device__ __forceinline__ void offset_one_tri(const RefView &v, int id, V3 &P, V3 N, float dist)
{
    {   /* conservative pre-test with the CURRENT position: the distance to a triangle is at least the distance to
         * its box, so a box farther than dist (with slack for rounding) cannot pass `ndst < dist` */
        const float4 bl = __ldg(v.boxes + 2ull * id), bh = __ldg(v.boxes + 2ull * id + 1);
        const float dx = fmaxf(fmaxf(bl.x - P.x, P.x - bh.x), 0.f), dy = fmaxf(fmaxf(bl.y - P.y, P.y - bh.y), 0.f), dz = fmaxf(fmaxf(bl.z - P.z, P.z - bh.z), 0.f);
        /* slack: 1 % of dist plus the worst rounding error of the reference's float evaluation at these coordinate
         * magnitudes (dot products of ~|P|-sized terms: a few ulps of |P|), so the pre-test can never reject a triangle
         * the exact test would accept */
        const float lim = dist * 1.01f + 2e-6f * (1.0f + fabsf(P.x) + fabsf(P.y) + fabsf(P.z));
        if (dx * dx + dy * dy + dz * dz > lim * lim) return;
    }
    PreparedTri PT;
    load_prepared(v.ptris + id, PT);
    V3 Pnew, projTPN;
    if (!offset_tri_tests(PT, P, N, dist, Pnew, projTPN)) return;
    int tid = -1;
    float d = reftree_closest(v, P + (N + projTPN) * LB_SMALL, Pnew, &tid);
    if (d >= 0.9f || tid == id) P = Pnew;
}

/* Concave-edge offset against one instance (ref: lighter_math.cpp:991-1044).  The node culling uses the box
 * around the INITIAL position (the reference builds its query box once), so the walk only has to list the
 * triangles of the overlapped nodes in reference order; they are then evaluated in that order, in batches,
 * with every lane of the warp inside the same loop -- the evaluation is where the time goes, and a lane-
 * private walk-and-evaluate loop ran with 6 of 32 lanes active (ncu). */
#define OFFSET_BATCH 24
__device__ void reftree_offset_sample(const RefView &v, V3 &P, V3 N, float dist, bool active)
{
    const V3 qlo = P - mk3(dist), qhi = P + mk3(dist);
    int stack[32];                                        /* reference trees are at most 16 deep: never fills */
    int list[OFFSET_BATCH];
    int sp = 0, ln = 0;
    if (active) stack[sp++] = 0;
    for (;;) {
        /* walk until the batch is full or the tree is exhausted */
        int pend_ido = -1, pend_k = 0, pend_cnt = 0;
        while (sp && ln < OFFSET_BATCH) {
            int node = stack[--sp];
            if (node < 0) {                               /* resume marker: remaining items of a node that did not fit the last batch */
                pend_ido = stack[--sp]; pend_k = -node - 1; pend_cnt = v.items[pend_ido];
            } else {
                RefNode Nd = v.nodes[node];
                if (qlo.x > Nd.hi.x || qhi.x < Nd.lo.x || qlo.y > Nd.hi.y || qhi.y < Nd.lo.y || qlo.z > Nd.hi.z || qhi.z < Nd.lo.z) continue;
                if (Nd.ch != -1 && sp + 2 <= 28) { stack[sp++] = Nd.ch; stack[sp++] = node + 1; }
                if (Nd.ido == -1) continue;
                pend_ido = Nd.ido; pend_k = 0; pend_cnt = v.items[Nd.ido];
            }
            while (pend_k < pend_cnt && ln < OFFSET_BATCH) list[ln++] = v.items[pend_ido + 1 + pend_k++];
            if (pend_k < pend_cnt) { stack[sp++] = pend_ido; stack[sp++] = -pend_k - 1; }     /* continue this node first next time */
        }
        const unsigned any = __ballot_sync(0xffffffffu, ln > 0);
        if (!any) break;
        int maxn = ln;
        for (int o = 16; o > 0; o >>= 1) maxn = max(maxn, __shfl_xor_sync(0xffffffffu, maxn, o));
        for (int k = 0; k < maxn; ++k)
            if (k < ln) offset_one_tri(v, list[k], P, N, dist);
        ln = 0;
    }
}

/*
 * Which lumels can the concave-edge offset move at all?  The reference walks every instance's tree for every lumel
 * (lighter.cpp:445-446) and mutates P as it goes, so the outcome depends on the order of the triangles -- but the FIRST
 * triangle to move a sample is evaluated at the sample's original position, and must pass every test of SampleOffsetQuery
 * there.  A lumel for which no triangle of any instance passes those tests at P0 is therefore left where it is, whatever
 * the order.  This kernel answers that question with a distance-bounded walk of a flat BVH over the triangles of ALL
 * instance trees (the scene BVH when every instance casts shadows), evaluating the reference's own test chain
 * (offset_tri_tests) at the leaves; lumels with a potential mover are listed for the exact reference-order pass.
 * Typically a few per cent of the lumels (the texels along concave edges); one coherent query per lumel instead of a
 * divergent walk of reference-order trees whose leaves can hold thousands of triangles (flat ceilings: volume-0 nodes
 * are never split, lighter_math.cpp:712-736).
 */
__device__ __forceinline__ bool lumel_has_mover(const BvhNode *__restrict__ nodes, const PreparedTri *__restrict__ tris, V3 P, V3 N, float dist)
{
    /* same slack as the per-triangle pre-test of offset_one_tri: never rejects a triangle the exact test would accept */
    const float lim = dist * 1.01f + 2e-6f * (1.0f + fabsf(P.x) + fabsf(P.y) + fabsf(P.z));
    const float lim2 = lim * lim;
    int stack[BVH_STACK];
    int sp = 0, node = 0;
    for (;;) {
        const float4 *n4 = reinterpret_cast<const float4 *>(nodes + node);
        const float4 a = __ldg(n4), b = __ldg(n4 + 1), c = __ldg(n4 + 2);
        const int4 k = __ldg(reinterpret_cast<const int4 *>(n4 + 3));
        const float d0 = box_dist2(P, a.x, a.y, a.z, a.w, b.x, b.y), d1 = box_dist2(P, b.z, b.w, c.x, c.y, c.z, c.w);
        int next = -1;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int cc = side ? k.y : k.x;
            if (!((side ? d1 : d0) <= lim2)) continue;
            if (cc < 0) {
                const unsigned code = ~cc, first = code >> 3, cnt = code & 7u;
                for (unsigned t = 0; t < cnt; ++t) {
                    PreparedTri PT;
                    load_prepared(tris + first + t, PT);
                    V3 Pnew, dir;
                    if (offset_tri_tests(PT, P, N, dist, Pnew, dir)) return true;
                }
            } else if (next < 0) next = cc;
            else stack[sp++] = cc;
        }
        if (next >= 0) node = next;
        else if (sp) node = stack[--sp];
        else return false;
    }
}

__global__ void __launch_bounds__(LB_BLOCK)
lumel_classify_kernel(const BvhNode *__restrict__ bvh, const PreparedTri *__restrict__ ptris, uint64_t first, uint64_t n_lumels,
                      const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, const float4 *__restrict__ lrad,
                      uint32_t *__restrict__ list, uint32_t *list_count)
{
    const uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool flag = false;
    if (i < n_lumels) flag = bvh == nullptr || lumel_has_mover(bvh, ptris, ld3(lpos[i]), ld3(lnrm[i]), sqrtf(lrad[i].w));
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (!m) return;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(list_count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (flag) list[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(i - first);
}

/* The exact pass for the listed lumels: every instance in order, its tree walked in the reference's order with the sample
 * position as it stands (ref: lighter.cpp:445-446, lighter_math.cpp:991-1044). */
#ifndef LB_FIX_MINBLOCKS
#define LB_FIX_MINBLOCKS 6        /* 80 registers: the exact pass only sees the few per cent of lumels near concave edges; no spills */
#endif
__global__ void __launch_bounds__(LB_BLOCK, LB_FIX_MINBLOCKS)
lumel_offset_kernel(const ltrgpu_Inst *__restrict__ inst, uint32_t n_inst, RefView all, uint64_t first, const uint32_t *__restrict__ list,
                    const uint32_t *__restrict__ list_count, float4 *lpos, const float4 *__restrict__ lnrm, const float4 *__restrict__ lrad)
{
    const uint32_t n = *list_count;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; w0 < n; w0 += warps * 32u) {
        const bool active = w0 + lane < n;
        uint64_t i = 0;
        V3 P = mk3(0.f), N = mk3(0.f);
        float dist = 0.f;
        if (active) { i = first + list[w0 + lane]; P = ld3(lpos[i]); N = ld3(lnrm[i]); dist = sqrtf(lrad[i].w); }
        for (uint32_t base = 0; base < n_inst; base += 32) {
            /* instances of this group whose root box meets the union of the lanes' query boxes AS THE SAMPLES STAND NOW;
             * the union is taken again whenever an instance has moved a sample (the reference builds the query box of
             * every instance around the current position) */
            int done_to = (int)base - 1;
            bool fresh = true;
            unsigned mask = 0;
            for (;;) {
                if (fresh) {
                    float ul[3] = { active ? P.x - dist : INFINITY, active ? P.y - dist : INFINITY, active ? P.z - dist : INFINITY };
                    float uh[3] = { active ? P.x + dist : -INFINITY, active ? P.y + dist : -INFINITY, active ? P.z + dist : -INFINITY };
                    for (int o = 16; o > 0; o >>= 1)
                        for (int a = 0; a < 3; ++a) { ul[a] = fminf(ul[a], __shfl_xor_sync(0xffffffffu, ul[a], o)); uh[a] = fmaxf(uh[a], __shfl_xor_sync(0xffffffffu, uh[a], o)); }
                    const uint32_t m = base + lane;
                    bool ok = false;
                    if (m < n_inst && (int)m > done_to) {
                        const RefNode R = all.nodes[inst[m].node_off];
                        ok = !(ul[0] > R.hi.x || uh[0] < R.lo.x || ul[1] > R.hi.y || uh[1] < R.lo.y || ul[2] > R.hi.z || uh[2] < R.lo.z);
                    }
                    mask = __ballot_sync(0xffffffffu, ok);
                    fresh = false;
                }
                if (!mask) break;
                const uint32_t mm = base + (uint32_t)__ffs(mask) - 1u;
                mask &= mask - 1u;
                done_to = (int)mm;
                const ltrgpu_Inst I = inst[mm];
                RefView v = { all.nodes + I.node_off, all.items + I.item_off, all.tris9 + 9ull * I.tri_off, all.ptris + I.tri_off, all.boxes + 2ull * I.tri_off };
                const V3 before = P;
                reftree_offset_sample(v, P, N, dist, active);     /* lanes whose own box misses the tree fall out at its root */
                if (__any_sync(0xffffffffu, P.x != before.x || P.y != before.y || P.z != before.z)) fresh = true;
            }
        }
        if (active) lpos[i] = make_float4(P.x, P.y, P.z, 0.f);
    }
}

/* Overlap correction (ref: lighter.cpp:449-465): up to 100 closest-hit probes along the normal; one lumel per thread,
 * neighbouring lumels in a warp. */
#ifndef LB_CORR_MINBLOCKS
#define LB_CORR_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(LB_BLOCK, LB_CORR_MINBLOCKS)
lumel_correct_kernel(const BvhNode *__restrict__ bvh, const RayTri *__restrict__ raytris, const PreparedTri *__restrict__ ptris, const uint32_t *__restrict__ tri_orig,
                     uint64_t first, uint64_t n_lumels, float max_correct_dist, float corr_min_dot,
                     float4 *lpos, const float4 *__restrict__ lnrm, unsigned long long *counters)
{
    const uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned rays = 0;
    if (i < n_lumels) {
        V3 P = ld3(lpos[i]);
        const V3 N = ld3(lnrm[i]);
        int itsleft = 100;
        float md = max_correct_dist;
        const V3 PEnd = P + N * md;
        while (md > LB_SMALL && itsleft-- > 0) {
            V3 dn = norm3(PEnd - P);
            V3 mA = P + dn * LB_SMALL, mB = PEnd - dn * LB_SMALL;
            int slot = -1;
            TravStats ts = { 0, 0 };
            float q = bvh_segment<false>(bvh, raytris, tri_orig, mA, mB, &slot, ts);
            ++rays;
            V3 hitnrm = mk3(0.f);
            if (slot >= 0) {
                const PreparedTri *T = ptris + slot;
                hitnrm = tri_back_normal(T->t0, T->t1, T->t2);
            }
            if (-dot3(hitnrm, N) < corr_min_dot) break;
            if (q < LB_SMALL) q = LB_SMALL;
            P = P * (1.0f - q) + PEnd * q;
            md *= (1.0f - q);
        }
        lpos[i] = make_float4(P.x, P.y, P.z, 0.f);
    }
    count_add(counters, CNT_CORR_RAYS, rays);
}

__global__ void fill_rgb_kernel(float4 *rgb, uint64_t n, float r, float g, float b)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rgb[i] = make_float4(r, g, b, 0.f);
}

extern "C" int ltrgpu_generate_lumels(ltrgpu_Ctx *ctx, uint64_t *inst_lumel_off)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, st));
    TexelSpace ts = { ctx->d_inst, ctx->d_rtris, ctx->d_wpos, ctx->d_wnrm, ctx->d_vtex, ctx->d_ltex };
    const uint64_t nt = ctx->n_texels;
    if (nt >= 0xffffffffull) { snprintf(ctx->err, sizeof(ctx->err), "texel space too large"); return 1; }
    if (dev_alloc(ctx, &ctx->d_texkey, nt)) return 1;
    if (dev_alloc(ctx, &ctx->d_texidx, nt + 1)) return 1;
    CU_TRY(ctx, cudaMemsetAsync(ctx->d_texkey, 0, (nt ? nt : 1) * 4, st));

    uint32_t *d_block_sums = nullptr, *d_total = nullptr;
    uint64_t *d_off = nullptr;
    uint32_t nblocks = (uint32_t)((nt + SCAN_CHUNK - 1) / SCAN_CHUNK);
    if (dev_alloc(ctx, &d_block_sums, nblocks)) return 1;
    if (dev_alloc(ctx, &d_total, 1)) return 1;
    if (dev_alloc(ctx, &d_off, ctx->n_inst + 1)) return 1;
    CU_TRY(ctx, cudaMemsetAsync(d_total, 0, 4, st));

    if (ctx->n_rtris && nt) {
        const float margin1 = 0.5f + LB_SMALL;
        unsigned warps = ctx->n_rtris;
        unsigned blocks = grid_for((uint64_t)warps * 32, 256);
        unsigned cap = (unsigned)ctx->num_sms * 64;
        if (blocks > cap) blocks = cap;
        raster_kernel<<<blocks, 256, 0, st>>>(ts, ctx->n_rtris, margin1, 0u, ctx->d_texkey);
        CU_LAUNCH_CHECK(ctx);
        raster_kernel<<<blocks, 256, 0, st>>>(ts, ctx->n_rtris, 0.0f, 0x80000000u, ctx->d_texkey);
        CU_LAUNCH_CHECK(ctx);
    }
    if (nt) {
        lumel_flags_kernel<<<grid_for(nt, 256), 256, 0, st>>>(ts, nt, ctx->d_texkey, ctx->d_texidx);
        CU_LAUNCH_CHECK(ctx);
        scan_reduce_kernel<<<nblocks, SCAN_T, 0, st>>>(ctx->d_texidx, nt, d_block_sums);
        CU_LAUNCH_CHECK(ctx);
        scan_sums_kernel<<<1, SCAN_T, 0, st>>>(d_block_sums, nblocks, d_total);
        CU_LAUNCH_CHECK(ctx);
        scan_apply_kernel<<<nblocks, SCAN_T, 0, st>>>(ctx->d_texidx, nt, d_block_sums);
        CU_LAUNCH_CHECK(ctx);
    }
    gather_inst_offsets_kernel<<<grid_for(ctx->n_inst + 1, 128), 128, 0, st>>>(ctx->d_inst, ctx->n_inst, nt, ctx->d_texidx, d_total,
                                                                              ctx->n_probes, d_off);
    CU_LAUNCH_CHECK(ctx);
    free(ctx->h_inst_lumel_off);
    ctx->h_inst_lumel_off = (uint64_t *)malloc(sizeof(uint64_t) * (ctx->n_inst + 1));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->h_inst_lumel_off, d_off, sizeof(uint64_t) * (ctx->n_inst + 1), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    ctx->host_counters.d2h_bytes += sizeof(uint64_t) * (ctx->n_inst + 1);
    /* instance 0 holds the probes */
    ctx->h_inst_lumel_off[0] = 0;
    if (ctx->n_inst > 1) ctx->h_inst_lumel_off[1] = ctx->n_probes;
    const uint64_t n = ctx->h_inst_lumel_off[ctx->n_inst];
    ctx->n_lumels = n;
    memcpy(inst_lumel_off, ctx->h_inst_lumel_off, sizeof(uint64_t) * (ctx->n_inst + 1));

    if (dev_alloc(ctx, &ctx->d_lpos, n + LB_PAD)) return 1;
    if (dev_alloc(ctx, &ctx->d_lnrm, n + LB_PAD)) return 1;
    if (dev_alloc(ctx, &ctx->d_lrad, n + LB_PAD)) return 1;
    if (dev_alloc(ctx, &ctx->d_lrgb, n + LB_PAD)) return 1;
    if (dev_alloc(ctx, &ctx->d_lloc, n + LB_PAD)) return 1;
    if (dev_alloc(ctx, &ctx->d_linst, n + LB_PAD)) return 1;
    if (ctx->n_probes) {
        probe_emit_kernel<<<grid_for(ctx->n_probes, 128), 128, 0, st>>>(ctx->d_probe_pos, ctx->d_probe_nrm, ctx->n_probes, ctx->d_lpos,
                                                                       ctx->d_lnrm, ctx->d_lrad, ctx->d_lloc, ctx->d_linst);
        CU_LAUNCH_CHECK(ctx);
    }
    if (nt) {
        lumel_emit_kernel<<<grid_for(nt, 256), 256, 0, st>>>(ts, nt, ctx->d_texkey, ctx->d_texidx, ctx->n_probes, ctx->d_lpos, ctx->d_lnrm,
                                                            ctx->d_lrad, ctx->d_lloc, ctx->d_linst);
        CU_LAUNCH_CHECK(ctx);
    }
    if (n > ctx->n_probes) {
        /* The two sequential corrections are the expensive part of lumel generation; with several GPUs
         * each rank corrects its own contiguous lumel range and the positions are all-gathered
         * (16 B/lumel).  Raster, scan and emit above are replicated: they are cheap and deterministic. */
        const uint64_t world = ctx->world > 0 ? (uint64_t)ctx->world : 1;
        const uint64_t chunk = (n + world - 1) / world;
        uint64_t b = chunk * (uint64_t)ctx->rank, e = b + chunk;
        if (b > n) b = n;
        if (e > n) e = n;
        if (b < ctx->n_probes) b = ctx->n_probes;
        RefView all = { ctx->d_rnodes, ctx->d_ritems, ctx->d_rtree_tris, ctx->d_rtree_ptris, ctx->d_rtree_boxes };
        if (e > b) {
            /* 1. which lumels can the concave-edge offset move (flat BVH over the triangles of every instance tree; without
             *    one -- LTR_LUMEL_CLASSIFY=0, or a scene too small to have it -- every lumel takes the exact pass) */
            uint32_t *d_list = nullptr, *d_list_count = nullptr;
            if (dev_alloc(ctx, &d_list, e - b)) return 1;
            if (dev_alloc(ctx, &d_list_count, 1)) return 1;
            CU_TRY(ctx, cudaMemsetAsync(d_list_count, 0, 4, st));
            const bool classify = ctx->d_lbvh != nullptr && !getenv("LTR_LUMEL_CLASSIFY_OFF");
            lumel_classify_kernel<<<grid_for(e - b, LB_BLOCK), LB_BLOCK, 0, st>>>(classify ? ctx->d_lbvh : nullptr, ctx->d_lbvh_ptris, b, e, ctx->d_lpos, ctx->d_lnrm,
                                                                                 ctx->d_lrad, d_list, d_list_count);
            CU_LAUNCH_CHECK(ctx);
            /* 2. the exact reference-order pass over the listed lumels (persistent warps over a device-side count) */
            lumel_offset_kernel<<<(unsigned)ctx->num_sms * 8u, LB_BLOCK, 0, st>>>(ctx->d_inst, ctx->n_inst, all, b, d_list, d_list_count, ctx->d_lpos, ctx->d_lnrm, ctx->d_lrad);
            CU_LAUNCH_CHECK(ctx);
            /* 3. overlap correction of every lumel */
            if (ctx->params.max_correct_dist) {
                lumel_correct_kernel<<<grid_for(e - b, LB_BLOCK), LB_BLOCK, 0, st>>>(ctx->d_bvh, ctx->d_raytris, ctx->d_ptris, ctx->d_tri_orig, b, e,
                                                                                    ctx->params.max_correct_dist, ctx->params.corr_min_dot, ctx->d_lpos, ctx->d_lnrm, ctx->d_counters);
                CU_LAUNCH_CHECK(ctx);
            }
            if (getenv("LTR_TRACE")) {
                uint32_t hc = 0;
                CU_TRY(ctx, cudaMemcpyAsync(&hc, d_list_count, 4, cudaMemcpyDeviceToHost, st));
                CU_TRY(ctx, cudaStreamSynchronize(st));
                fprintf(stderr, "[ltr rank %d] lumel offset: %u of %llu lumels take the reference-order pass\n", ctx->rank, hc, (unsigned long long)(e - b));
            }
            lb_free(d_list); lb_free(d_list_count);
        }
        if (world > 1) {
            if (!ctx->allgather || ctx->allgather(ctx->allgather_user, ctx->d_lpos + chunk * ctx->rank, ctx->d_lpos, chunk * sizeof(float4), st)) {
                snprintf(ctx->err, sizeof(ctx->err), "lumel generation: all-gather of positions failed");
                return 1;
            }
        }
    }
    if (n) {
        fill_rgb_kernel<<<grid_for(n, 256), 256, 0, st>>>(ctx->d_lrgb, n, ctx->params.ambient[0], ctx->params.ambient[1], ctx->params.ambient[2]);
        CU_LAUNCH_CHECK(ctx);
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->host_counters.ms_samples += ms;
    lb_free(d_block_sums); lb_free(d_total); lb_free(d_off);
    ctx->sh_begin = 0; ctx->sh_end = n;          /* refined by ltrgpu_set_shard */
    return 0;
}

/* one ltr_SampleRequest's worth of fields per mesh lumel (ref: lighter.cpp:690-706) */
__global__ void sample_request_kernel(const float4 *__restrict__ lpos, const float4 *__restrict__ lnrm, const float4 *__restrict__ lrad,
                                      const uint32_t *__restrict__ lloc, const uint32_t *__restrict__ linst, const ltrgpu_Inst *__restrict__ inst,
                                      uint64_t first, uint32_t count, ltrgpu_SampleReq *__restrict__ out)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint64_t i = first + k;
    const float4 P = lpos[i], R = lrad[i];
    const V3 N = norm3(ld3(lnrm[i]));
    const uint32_t in = linst[i], L = lloc[i];
    const uint32_t w = inst[in].lm_w, h = inst[in].lm_h;
    const int lx = (int)(L % w), ly = (int)(L / w);
    ltrgpu_SampleReq q;
    q.pos[0] = P.x; q.pos[1] = P.y; q.pos[2] = P.z;
    q.nrm[0] = N.x; q.nrm[1] = N.y; q.nrm[2] = N.z;
    q.tex0[0] = R.x; q.tex0[1] = R.y;
    q.tex1[0] = ((float)lx + 0.5f) / (float)w; q.tex1[1] = ((float)ly + 0.5f) / (float)h;
    q.part_id = (uint32_t)R.z; q.inst = in;
    out[k] = q;
}

#define AUX_TRY(ctx, call)                                                                         \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf((ctx)->aux_err, sizeof((ctx)->aux_err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

extern "C" int ltrgpu_sample_requests_begin(ltrgpu_Ctx *ctx)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->aux_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    if (!ctx->ev_lumels) CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_lumels, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) if (!ctx->ev_req[b]) CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_req[b], cudaEventDisableTiming));
    /* the second stream may read the lumel arrays once everything queued on the bake stream so far has run */
    CU_TRY(ctx, cudaEventRecord(ctx->ev_lumels, ctx->stream));
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_lumels, 0));
    ctx->aux_err[0] = 0;
    return 0;
}

extern "C" int ltrgpu_sample_requests_issue(ltrgpu_Ctx *ctx, uint64_t first, uint32_t count, ltrgpu_SampleReq *host_pinned, int slot)
{
    AUX_TRY(ctx, cudaSetDevice(ctx->device));             /* the calling thread is not the bake thread */
    if (!count) return 0;
    if (first + count > ctx->n_lumels || (slot != 0 && slot != 1)) { snprintf(ctx->aux_err, sizeof(ctx->aux_err), "sample request range out of bounds"); return 1; }
    if (ctx->req_cap[slot] < count) {
        if (ctx->d_req[slot]) lb_free(ctx->d_req[slot]);
        ctx->d_req[slot] = nullptr; ctx->req_cap[slot] = 0;
        AUX_TRY(ctx, lb_malloc(&ctx->d_req[slot], (size_t)count * sizeof(ltrgpu_SampleReq)));
        ctx->req_cap[slot] = count;
    }
    sample_request_kernel<<<grid_for(count, 256), 256, 0, ctx->aux_stream>>>(ctx->d_lpos, ctx->d_lnrm, ctx->d_lrad, ctx->d_lloc, ctx->d_linst, ctx->d_inst,
                                                                               first, count, ctx->d_req[slot]);
    AUX_TRY(ctx, cudaGetLastError());
    AUX_TRY(ctx, cudaMemcpyAsync(host_pinned, ctx->d_req[slot], (size_t)count * sizeof(ltrgpu_SampleReq), cudaMemcpyDeviceToHost, ctx->aux_stream));
    AUX_TRY(ctx, cudaEventRecord(ctx->ev_req[slot], ctx->aux_stream));
    ctx->aux_launches += 1; ctx->aux_d2h_bytes += (unsigned long long)count * sizeof(ltrgpu_SampleReq);
    return 0;
}

extern "C" int ltrgpu_sample_requests_wait(ltrgpu_Ctx *ctx, int slot)
{
    AUX_TRY(ctx, cudaEventSynchronize(ctx->ev_req[slot]));
    return 0;
}

extern "C" const char *ltrgpu_aux_error(ltrgpu_Ctx *ctx) { return ctx->aux_err; }

extern "C" int ltrgpu_download_lumels(ltrgpu_Ctx *ctx, float *pos3, float *nrm3, uint32_t *loc, float *radinfo4, float *rgb3)
{
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = ctx->n_lumels;
    if (!n) return 0;
    float4 *tmp = (float4 *)malloc(sizeof(float4) * n);
    struct { float4 *src; float *dst; int k; } jobs[4] = { { ctx->d_lpos, pos3, 3 }, { ctx->d_lnrm, nrm3, 3 }, { ctx->d_lrad, radinfo4, 4 }, { ctx->d_lrgb, rgb3, 3 } };
    for (int j = 0; j < 4; ++j) {
        if (!jobs[j].dst) continue;
        cudaError_t e = cudaMemcpyAsync(tmp, jobs[j].src, sizeof(float4) * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { free(tmp); snprintf(ctx->err, sizeof(ctx->err), "download_lumels: %s", cudaGetErrorString(e)); return 1; }
        ctx->host_counters.d2h_bytes += sizeof(float4) * n;
        for (uint64_t i = 0; i < n; ++i) {
            jobs[j].dst[i * jobs[j].k + 0] = tmp[i].x; jobs[j].dst[i * jobs[j].k + 1] = tmp[i].y; jobs[j].dst[i * jobs[j].k + 2] = tmp[i].z;
            if (jobs[j].k == 4) jobs[j].dst[i * 4 + 3] = tmp[i].w;
        }
    }
    free(tmp);
    if (loc) {
        CU_TRY(ctx, cudaMemcpyAsync(loc, ctx->d_lloc, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->host_counters.d2h_bytes += 4 * n;
    }
    return 0;
}
