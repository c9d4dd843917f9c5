/* bvh.cpp -- host builder for the flat scene BVH (see bvh.h): binned SAH over all three axes,
 * parallel over sub-trees, flattened depth-first so that a parent and its first child are adjacent. */
#include "bvh.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <memory>
#include <utility>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <emmintrin.h>
#include <sys/mman.h>

namespace {

const float FMAXV = 3.402823466e+38f;

struct Tmp {
    Box3 box;
    int32_t left, right;       /* -1 for a leaf */
    uint32_t first, count;
    int32_t inner;             /* inner nodes in this sub-tree (0 for a leaf); -1 = not known yet */
    int32_t inner_even;        /* ... of which at even depth below this node (itself included): the 4-wide nodes of the sub-tree */
    int32_t height;            /* levels of inner nodes below and including this one */
};

inline void grow(Box3 &b, const Box3 &o) { b.lo = min3(b.lo, o.lo); b.hi = max3(b.hi, o.hi); }
inline void grow(Box3 &b, V3 p) { b.lo = min3(b.lo, p); b.hi = max3(b.hi, p); }
inline float half_area(const Box3 &b)
{
    V3 e = b.hi - b.lo;
    return e.x * e.y + e.y * e.z + e.z * e.x;
}
inline float axis_of(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

/* One primitive of the build: its box and original index, kept in a PERMUTED array so that every pass
 * over a node's range is sequential in memory (an index array into fixed boxes costs a cache miss per
 * triangle at the top levels, which is where the time goes at 1M triangles). */
struct Prim { Box3 b; uint32_t id; };
inline V3 centroid(const Prim &p) { return (p.b.lo + p.b.hi) * 0.5f; }

const int NB = 16;                                /* SAH bins per axis */
const uint32_t SPARSE_MAX = 96;                   /* nodes up to this size bin without clearing (bin_range_sparse) */
const uint32_t FINE_MIN = 4096;                   /* unit of the dynamically scheduled part of the build */
const uint32_t COOP_MIN = 65536;                  /* nodes above this size are built by all threads together */

struct Bounds { Box3 bb, cb; };

/* SSE2 (baseline x86-64) forms of the two per-triangle passes.  A Prim is 7 floats: lo.xyz, hi.xyz, id -- an unaligned
 * 4-float load at float 0 gives lo in lanes 0-2, one at float 3 gives hi in lanes 0-2 (lane 3 holds a neighbouring field and
 * is never read back: min/max/add/mul are lane-wise).  min/max of floats is exact, and the centroid / bin arithmetic is the
 * same IEEE single-precision sequence as the scalar code, so the tree is the one the scalar builder produces. */
inline __m128 ld_lo(const Prim &p) { return _mm_loadu_ps(&p.b.lo.x); }
inline __m128 ld_hi(const Prim &p) { return _mm_loadu_ps(&p.b.lo.x + 3); }
inline V3 v3_of(__m128 v) { alignas(16) float f[4]; _mm_store_ps(f, v); return mk3(f[0], f[1], f[2]); }

struct alignas(16) Bins {
    __m128 lo[3][NB], hi[3][NB];
    uint32_t cnt[3][NB];
    uint32_t mask[3];                             /* bit b: bin b of that axis holds something (the only valid bins) */
    void clear()                                  /* dense use: every bin initialised, masks derived afterwards (set_masks) */
    {
        const __m128 big = _mm_set1_ps(FMAXV), small = _mm_set1_ps(-FMAXV);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < NB; ++b) { lo[a][b] = big; hi[a][b] = small; cnt[a][b] = 0; }
    }
    void merge(const Bins &o)
    {
        for (int a = 0; a < 3; ++a) for (int b = 0; b < NB; ++b) { lo[a][b] = _mm_min_ps(lo[a][b], o.lo[a][b]); hi[a][b] = _mm_max_ps(hi[a][b], o.hi[a][b]); cnt[a][b] += o.cnt[a][b]; }
    }
    void set_masks()
    {
        for (int a = 0; a < 3; ++a) { uint32_t m = 0; for (int b = 0; b < NB; ++b) m |= (cnt[a][b] ? 1u : 0u) << b; mask[a] = m; }
    }
};

inline void bounds_of(const Prim *p, size_t n, Bounds &o)
{
    const __m128 half = _mm_set1_ps(0.5f);
    __m128 blo = _mm_set1_ps(FMAXV), bhi = _mm_set1_ps(-FMAXV), clo = blo, chi = bhi;
    for (size_t i = 0; i < n; ++i) {
        const __m128 l = ld_lo(p[i]), h = ld_hi(p[i]);
        const __m128 c = _mm_mul_ps(_mm_add_ps(l, h), half);
        blo = _mm_min_ps(blo, l); bhi = _mm_max_ps(bhi, h);
        clo = _mm_min_ps(clo, c); chi = _mm_max_ps(chi, c);
    }
    o.bb.lo = v3_of(blo); o.bb.hi = v3_of(bhi); o.cb.lo = v3_of(clo); o.cb.hi = v3_of(chi);
}
inline int bin_of(float c, float lo, float scale)
{
    int b = (int)((c - lo) * scale);
    return b < 0 ? 0 : (b >= NB ? NB - 1 : b);
}
inline int clamp_bin(int b) { return b < 0 ? 0 : (b >= NB ? NB - 1 : b); }
/* all three axes in one pass over the primitives (an unused axis has scale 0: everything lands in its bin 0, which nobody reads) */
inline void bin_range(const Prim *p, size_t n, const Box3 &cb, const bool use[3], const float scale[3], Bins &o)
{
    (void)use;
    const __m128 half = _mm_set1_ps(0.5f), cblo = _mm_set_ps(0.f, cb.lo.z, cb.lo.y, cb.lo.x), sc = _mm_set_ps(0.f, scale[2], scale[1], scale[0]);
    for (size_t i = 0; i < n; ++i) {
        const __m128 l = ld_lo(p[i]), h = ld_hi(p[i]);
        const __m128 c = _mm_mul_ps(_mm_add_ps(l, h), half);
        const __m128i bi = _mm_cvttps_epi32(_mm_mul_ps(_mm_sub_ps(c, cblo), sc));
        const int bx = clamp_bin(_mm_cvtsi128_si32(bi)), by = clamp_bin(_mm_cvtsi128_si32(_mm_shuffle_epi32(bi, 1))), bz = clamp_bin(_mm_cvtsi128_si32(_mm_shuffle_epi32(bi, 2)));
        o.cnt[0][bx]++; o.lo[0][bx] = _mm_min_ps(o.lo[0][bx], l); o.hi[0][bx] = _mm_max_ps(o.hi[0][bx], h);
        o.cnt[1][by]++; o.lo[1][by] = _mm_min_ps(o.lo[1][by], l); o.hi[1][by] = _mm_max_ps(o.hi[1][by], h);
        o.cnt[2][bz]++; o.lo[2][bz] = _mm_min_ps(o.lo[2][bz], l); o.hi[2][bz] = _mm_max_ps(o.hi[2][bz], h);
    }
}
/* The same for a SMALL node (most nodes are: half a million of them hold < 16 triangles): nothing is cleared, a bin is
 * initialised by the first triangle that lands in it (mask bit), so the cost is per triangle, not per bin. */
inline void bin_range_sparse(const Prim *p, size_t n, const Box3 &cb, const float scale[3], Bins &o)
{
    o.mask[0] = o.mask[1] = o.mask[2] = 0;
    const __m128 half = _mm_set1_ps(0.5f), cblo = _mm_set_ps(0.f, cb.lo.z, cb.lo.y, cb.lo.x), sc = _mm_set_ps(0.f, scale[2], scale[1], scale[0]);
    for (size_t i = 0; i < n; ++i) {
        const __m128 l = ld_lo(p[i]), h = ld_hi(p[i]);
        const __m128 c = _mm_mul_ps(_mm_add_ps(l, h), half);
        const __m128i bi = _mm_cvttps_epi32(_mm_mul_ps(_mm_sub_ps(c, cblo), sc));
        const int bin[3] = { clamp_bin(_mm_cvtsi128_si32(bi)), clamp_bin(_mm_cvtsi128_si32(_mm_shuffle_epi32(bi, 1))), clamp_bin(_mm_cvtsi128_si32(_mm_shuffle_epi32(bi, 2))) };
        for (int a = 0; a < 3; ++a) {
            const int b = bin[a];
            if (o.mask[a] & (1u << b)) { o.cnt[a][b]++; o.lo[a][b] = _mm_min_ps(o.lo[a][b], l); o.hi[a][b] = _mm_max_ps(o.hi[a][b], h); }
            else { o.mask[a] |= 1u << b; o.cnt[a][b] = 1; o.lo[a][b] = l; o.hi[a][b] = h; }
        }
    }
}
inline float half_area_v(__m128 lo, __m128 hi)
{
    alignas(16) float e[4];
    _mm_store_ps(e, _mm_sub_ps(hi, lo));
    return e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
}

/* A team of threads that lives for one build: the cooperative top of the tree runs ~55 short parallel passes, and spawning
 * 15 threads for each of them cost more than the passes themselves (0.3-1 ms per spawn round on the bench host).  Plain
 * mutex + condition variables, one job at a time, the caller is member 0. */
class ThreadTeam {
public:
    explicit ThreadTeam(int members) : T(members)
    {
        for (int t = 1; t < T; ++t) th.emplace_back([this, t]() { member(t); });
    }
    ~ThreadTeam()
    {
        { std::lock_guard<std::mutex> g(m); quit = true; }
        cv_work.notify_all();
        for (auto &x : th) x.join();
    }
    int size() const { return T; }
    /* runs f(0) .. f(T-1), one call per member, and returns when all are done */
    void run(const std::function<void(int)> &f)
    {
        if (T <= 1) { f(0); return; }
        { std::lock_guard<std::mutex> g(m); job = &f; remaining = T - 1; ++epoch; }
        cv_work.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [this]() { return remaining == 0; });
        job = nullptr;
    }
private:
    void member(int t)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)> *j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_work.wait(lk, [&]() { return quit || epoch != seen; });
                if (quit) return;
                seen = epoch;
                j = job;
            }
            (*j)(t);
            {
                std::lock_guard<std::mutex> g(m);
                if (--remaining == 0) cv_done.notify_one();
            }
        }
    }
    const int T;
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    const std::function<void(int)> *job = nullptr;
    uint64_t epoch = 0;
    int remaining = 0;
    bool quit = false;
};
thread_local ThreadTeam *tl_team = nullptr;       /* the team of the build running on this thread, if any */

template <class F> void parallel_chunks(int threads, size_t n, F fn)      /* fn(chunk index, begin, end), chunk count == threads */
{
    const size_t per = (n + threads - 1) / threads;
    if (threads <= 1) { fn(0, 0, std::min(n, per)); return; }
    if (tl_team && tl_team->size() == threads) {
        tl_team->run([&](int t) { const size_t b = std::min(n, per * t), e = std::min(n, b + per); fn(t, b, e); });
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
        const size_t b = std::min(n, per * t), e = std::min(n, b + per);
        pool.emplace_back([=]() { fn(t, b, e); });
    }
    fn(0, 0, std::min(n, per));
    for (auto &th : pool) th.join();
}

/* `worker` on up to `members` threads (the team's, when this build has one of at least that size) */
template <class F> void parallel_workers(int members, F worker)
{
    if (members <= 1) { worker(); return; }
    if (tl_team && tl_team->size() >= members) { tl_team->run([&](int t) { if (t < members) worker(); }); return; }
    std::vector<std::thread> pool;
    for (int t = 1; t < members; ++t) pool.emplace_back(worker);
    worker();
    for (auto &th : pool) th.join();
}

struct Work {
    int32_t node; uint32_t first, count; int depth;
    bool has_nb = false;          /* nb = bounds of the range, accumulated by the parent's partition pass (saves a pass over the range) */
    Bounds nb;
};

/* bounds of one side of a partition, accumulated while the triangles are being moved */
struct SideAcc {
    __m128 blo, bhi, clo, chi;
    void clear() { blo = clo = _mm_set1_ps(FMAXV); bhi = chi = _mm_set1_ps(-FMAXV); }
    void add(__m128 l, __m128 h, __m128 c) { blo = _mm_min_ps(blo, l); bhi = _mm_max_ps(bhi, h); clo = _mm_min_ps(clo, c); chi = _mm_max_ps(chi, c); }
    void merge(const SideAcc &o) { blo = _mm_min_ps(blo, o.blo); bhi = _mm_max_ps(bhi, o.bhi); clo = _mm_min_ps(clo, o.clo); chi = _mm_max_ps(chi, o.chi); }
    Bounds get() const { Bounds b; b.bb.lo = v3_of(blo); b.bb.hi = v3_of(bhi); b.cb.lo = v3_of(clo); b.cb.hi = v3_of(chi); return b; }
};

/* bin (unclamped) of a centroid on axis A: the arithmetic of bin_range */
template <int A> inline int axis_bin(__m128 c, __m128 cblo, __m128 sc4)
{
    const __m128i bi = _mm_cvttps_epi32(_mm_mul_ps(_mm_sub_ps(c, cblo), sc4));
    return A == 0 ? _mm_cvtsi128_si32(bi) : _mm_cvtsi128_si32(_mm_shuffle_epi32(bi, A));
}

/* std::partition's two-pointer scheme written out (same swaps, same resulting order) with "left = centroid bin on axis A
 * below sp" as the predicate; every triangle is looked at exactly once, and while it is in registers its box and centroid
 * are added to the bounds of the side it lands on (the accumulators stay in registers: no call, no captures). */
template <int A> uint32_t partition_acc(Prim *P, uint32_t count, int sp, __m128 cblo, __m128 sc4, SideAcc &AL, SideAcc &AR)
{
    const __m128 half = _mm_set1_ps(0.5f);
    __m128 lbl = AL.blo, lbh = AL.bhi, lcl = AL.clo, lch = AL.chi, rbl = AR.blo, rbh = AR.bhi, rcl = AR.clo, rch = AR.chi;
#define LB_SIDE(PRIM, LEFT)                                                                                   \
    {                                                                                                         \
        const __m128 l_ = ld_lo(PRIM), h_ = ld_hi(PRIM);                                                      \
        const __m128 c_ = _mm_mul_ps(_mm_add_ps(l_, h_), half);                                               \
        LEFT = clamp_bin(axis_bin<A>(c_, cblo, sc4)) < sp;                                                    \
        if (LEFT) { lbl = _mm_min_ps(lbl, l_); lbh = _mm_max_ps(lbh, h_); lcl = _mm_min_ps(lcl, c_); lch = _mm_max_ps(lch, c_); } \
        else      { rbl = _mm_min_ps(rbl, l_); rbh = _mm_max_ps(rbh, h_); rcl = _mm_min_ps(rcl, c_); rch = _mm_max_ps(rch, c_); } \
    }
    Prim *first = P, *last = P + count;
    for (;;) {
        bool left;
        for (;;) { if (first == last) goto done; LB_SIDE(*first, left) if (left) ++first; else break; }
        --last;
        for (;;) { if (first == last) goto done; LB_SIDE(*last, left) if (!left) --last; else break; }
        const Prim t = *first; *first = *last; *last = t;
        ++first;
    }
done:
#undef LB_SIDE
    AL.blo = lbl; AL.bhi = lbh; AL.clo = lcl; AL.chi = lch; AR.blo = rbl; AR.bhi = rbh; AR.clo = rcl; AR.chi = rch;
    return (uint32_t)(first - P);
}

struct Builder {
    Prim *prim;
    Prim *scratch;                                /* same size as prim: target of the parallel stable partition */
    Tmp *tmp;                                     /* 2*count+1 entries, uninitialised (every node is written before it is read) */
    std::atomic<int32_t> next{0};
    int leaf_max;
    int threads;

    int32_t alloc() { return next.fetch_add(1); }

    /* Builds ONE node: bounds, SAH split, partition.  Returns false for a leaf; otherwise the children
     * (allocated, ranges set) are left in l / r.  coop = use every thread on this node's range.  The tree
     * does not depend on the thread count: bins merge associatively and the cooperative partition is a
     * stable one (deterministic), chosen by node size only. */
    bool split(const Work &w, bool coop, Work &l, Work &r)
    {
        Prim *P = prim + w.first;
        const uint32_t count = w.count;
        const int T = coop ? threads : 1;
        Bounds nb;
        if (w.has_nb) nb = w.nb;
        else if (T > 1) {
            std::vector<Bounds> part(T);
            parallel_chunks(T, count, [&](int t, size_t b, size_t e) { bounds_of(P + b, e - b, part[t]); });
            nb = part[0];
            for (int t = 1; t < T; ++t) { grow(nb.bb, part[t].bb); grow(nb.cb, part[t].cb); }
        } else bounds_of(P, count, nb);
        Tmp &N = tmp[w.node];
        N.box = nb.bb; N.first = w.first; N.count = count; N.left = N.right = -1; N.inner = 0; N.inner_even = 0; N.height = 0;
        if ((int)count <= leaf_max) return false;
        N.inner = -1;

        const Box3 cb = nb.cb;
        const float ext3[3] = { cb.hi.x - cb.lo.x, cb.hi.y - cb.lo.y, cb.hi.z - cb.lo.z };
        bool use[3]; float scale[3];
        for (int a = 0; a < 3; ++a) { use[a] = ext3[a] > 0; scale[a] = use[a] ? NB / ext3[a] : 0.f; }
        Bins bins;
        std::vector<Bins> part;                           /* cooperative path: per-chunk bins, kept for the partition's left counts */
        if (T > 1) {
            bins.clear();
            part.resize(T);
            parallel_chunks(T, count, [&](int t, size_t b, size_t e) { part[t].clear(); bin_range(P + b, e - b, cb, use, scale, part[t]); });
            for (int t = 0; t < T; ++t) bins.merge(part[t]);
            bins.set_masks();
        } else if (count > SPARSE_MAX) {
            bins.clear();
            bin_range(P, count, cb, use, scale, bins);
            bins.set_masks();
        } else bin_range_sparse(P, count, cb, scale, bins);

        int best_axis = -1, best_split = 0;
        float best_cost = FMAXV;
        for (int a = 0; a < 3; ++a) {
            if (!use[a]) continue;
            /* Sweep over the NON-EMPTY bins only.  Between two neighbouring non-empty bins every split position gives the
             * same two sides, hence the same cost; the dense sweep (positions NB-1 down to 1, strict <) keeps the highest of
             * them, which is the non-empty bin that opens the right side -- so the choice below is the dense sweep's. */
            float la[NB]; uint32_t lc[NB]; int bs[NB];
            int K = 0;
            __m128 alo = _mm_set1_ps(FMAXV), ahi = _mm_set1_ps(-FMAXV);
            uint32_t c = 0;
            for (uint32_t m = bins.mask[a]; m; m &= m - 1) {
                const int b = __builtin_ctz(m);
                alo = _mm_min_ps(alo, bins.lo[a][b]); ahi = _mm_max_ps(ahi, bins.hi[a][b]); c += bins.cnt[a][b];
                la[K] = half_area_v(alo, ahi); lc[K] = c; bs[K] = b; ++K;
            }
            alo = _mm_set1_ps(FMAXV); ahi = _mm_set1_ps(-FMAXV); c = 0;
            for (int k = K - 1; k > 0; --k) {
                const int b = bs[k];
                alo = _mm_min_ps(alo, bins.lo[a][b]); ahi = _mm_max_ps(ahi, bins.hi[a][b]); c += bins.cnt[a][b];
                float cost = la[k - 1] * lc[k - 1] + half_area_v(alo, ahi) * c;
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
            }
        }

        uint32_t mid;
        l.has_nb = r.has_nb = false;
        if (best_axis < 0) {
            mid = count / 2;                               /* coincident centroids: split by index */
        } else if (w.depth > 40) {                         /* failsafe against degenerate SAH chains */
            V3 e = cb.hi - cb.lo;
            int a = (e.x >= e.y && e.x >= e.z) ? 0 : (e.y >= e.z ? 1 : 2);
            mid = count / 2;
            std::nth_element(P, P + mid, P + count, [=](const Prim &p, const Prim &q) { return axis_of(centroid(p), a) < axis_of(centroid(q), a); });
        } else {
            const int a = best_axis, sp = best_split;
            /* side of a triangle = bin of its centroid on the chosen axis (the arithmetic of bin_range), and while the
             * triangle is in registers its box and centroid go into the bounds of the side it lands on */
            const __m128 half = _mm_set1_ps(0.5f), cblo = _mm_set_ps(0.f, cb.lo.z, cb.lo.y, cb.lo.x), sc4 = _mm_set_ps(0.f, scale[2], scale[1], scale[0]);
            auto side = [=](const Prim &t, SideAcc &L, SideAcc &R) {
                const __m128 l = ld_lo(t), h = ld_hi(t);
                const __m128 c = _mm_mul_ps(_mm_add_ps(l, h), half);
                alignas(16) int bi[4];
                _mm_store_si128((__m128i *)bi, _mm_cvttps_epi32(_mm_mul_ps(_mm_sub_ps(c, cblo), sc4)));
                const bool left = clamp_bin(bi[a]) < sp;
                (left ? L : R).add(l, h, c);
                return left;
            };
            SideAcc AL, AR;
            AL.clear(); AR.clear();
            if (T > 1) {
                /* stable partition through the scratch array: left counts per chunk from the chunk's bins, scan, scatter, copy back */
                std::vector<uint32_t> nl(T + 1, 0);
                Prim *S = scratch + w.first;
                for (int t = 0; t < T; ++t) { uint32_t c = 0; for (int b = 0; b < sp; ++b) c += part[t].cnt[a][b]; nl[t + 1] = nl[t] + c; }
                const uint32_t total_left = nl[T];
                const size_t per = ((size_t)count + T - 1) / T;
                std::vector<SideAcc> pl(T), pr(T);
                parallel_chunks(T, count, [&](int t, size_t b, size_t e) {
                    size_t lpos = nl[t], rpos = total_left + (std::min((size_t)count, per * t) - nl[t]);
                    SideAcc L, R;
                    L.clear(); R.clear();
                    for (size_t i = b; i < e; ++i) { if (side(P[i], L, R)) S[lpos++] = P[i]; else S[rpos++] = P[i]; }
                    pl[t] = L; pr[t] = R;
                });
                for (int t = 0; t < T; ++t) { AL.merge(pl[t]); AR.merge(pr[t]); }
                parallel_chunks(T, count, [&](int, size_t b, size_t e) { if (e > b) memcpy(P + b, S + b, (e - b) * sizeof(Prim)); });
                mid = total_left;
            } else if (count > COOP_MIN) {                 /* same (stable) order as the cooperative path: the tree does not depend on the thread count */
                mid = (uint32_t)(std::stable_partition(P, P + count, [&](const Prim &t) { return side(t, AL, AR); }) - P);
            } else {
                mid = a == 0 ? partition_acc<0>(P, count, sp, cblo, sc4, AL, AR) : a == 1 ? partition_acc<1>(P, count, sp, cblo, sc4, AL, AR)
                                                                                           : partition_acc<2>(P, count, sp, cblo, sc4, AL, AR);
            }
            if (mid == 0 || mid == count) mid = count / 2;
            else { l.has_nb = r.has_nb = true; l.nb = AL.get(); r.nb = AR.get(); }
        }
        l.node = alloc(); r.node = alloc();
        tmp[w.node].left = l.node; tmp[w.node].right = r.node;
        l.first = w.first; l.count = mid; r.first = w.first + mid; r.count = count - mid;
        l.depth = r.depth = w.depth + 1;
        return true;
    }

    void build_serial(const Work &w)
    {
        Work l, r;
        if (!split(w, false, l, r)) return;
        build_serial(l);
        build_serial(r);
        tmp[w.node].inner = 1 + tmp[l.node].inner + tmp[r.node].inner;
        tmp[w.node].inner_even = 1 + (tmp[l.node].inner - tmp[l.node].inner_even) + (tmp[r.node].inner - tmp[r.node].inner_even);
        tmp[w.node].height = 1 + std::max(tmp[l.node].height, tmp[r.node].height);
    }

    /* serial splits down to FINE_MIN triangles; the nodes at that size are left for phase B2 */
    void expand(const Work &w, std::vector<Work> &out)
    {
        if (w.count <= FINE_MIN) { out.push_back(w); return; }
        Work l, r;
        if (!split(w, false, l, r)) return;
        expand(l, out);
        expand(r, out);
    }

    /* sub-tree sizes of the nodes built cooperatively (their descendants built in phase B are known) */
    void finish_sizes(int32_t t)
    {
        Tmp &N = tmp[t];
        if (N.inner >= 0) return;
        finish_sizes(N.left); finish_sizes(N.right);
        N.inner = 1 + tmp[N.left].inner + tmp[N.right].inner;
        N.inner_even = 1 + (tmp[N.left].inner - tmp[N.left].inner_even) + (tmp[N.right].inner - tmp[N.right].inner_even);
        N.height = 1 + std::max(tmp[N.left].height, tmp[N.right].height);
    }

    void run(uint32_t count)
    {
        Work root = { alloc(), 0, count, 0 };
        const double t_0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
        /* phase A: the few nodes big enough to be worth every thread, one after the other */
        std::vector<Work> big{ root }, small;
        while (!big.empty()) {
            Work w = big.back(); big.pop_back();
            if (w.count <= COOP_MIN || threads <= 1) { small.push_back(w); continue; }
            Work l, r;
            if (split(w, true, l, r)) { big.push_back(l); big.push_back(r); }
        }
        /* phase B1: every medium sub-tree (<= COOP_MIN triangles; only ~16 of them at 1 M triangles -- one per thread of a
         * 16-core host, i.e. no slack for their uneven sizes) is split further by ONE thread down to FINE_MIN triangles;
         * phase B2: the resulting few hundred small sub-trees, largest first, are pulled from a shared cursor.  The tree is
         * the same as before: split() does not depend on who calls it. */
        const bool trace_b = getenv("LTR_TRACE_BVH") != nullptr;
        if (trace_b) fprintf(stderr, "[bvh] phase A %.1f ms: %zu sub-trees\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_0, small.size());
        auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t_a = tnow();
        std::vector<Work> fine;
        {
            std::vector<std::vector<Work>> part(small.size());
            std::atomic<size_t> cursor{0};
            auto worker = [&]() { for (size_t i; (i = cursor.fetch_add(1)) < small.size();) expand(small[i], part[i]); };
            parallel_workers((int)std::min<size_t>((size_t)std::max(1, threads), small.size()), worker);
            for (auto &v : part) fine.insert(fine.end(), v.begin(), v.end());
        }
        small.swap(fine);
        if (trace_b) fprintf(stderr, "[bvh] phase B1 %.1f ms: %zu sub-trees\n", tnow() - t_a, small.size());
        std::sort(small.begin(), small.end(), [](const Work &a, const Work &b) { return a.count != b.count ? a.count > b.count : a.first < b.first; });
        std::atomic<size_t> cursor{0};
        auto worker = [&]() { for (size_t i; (i = cursor.fetch_add(1)) < small.size();) build_serial(small[i]); };
        parallel_workers((int)std::min<size_t>((size_t)std::max(1, threads), small.size()), worker);
        finish_sizes(root.node);
        subtrees.swap(small);
    }
    std::vector<Work> subtrees;                   /* the phase-B roots: units of the parallel flatten as well */
};

inline int32_t leaf_code(uint32_t first, uint32_t count) { return ~(int32_t)((first << 3) | count); }

/* Pre-order layout: an inner node's first child (if inner) is the next slot, the second child follows the
 * first child's whole sub-tree.  With the sub-tree sizes known, every slot is a pure function of the path
 * from the root, so the phase-B sub-trees are emitted by the threads independently. */
struct Flattener {
    const Tmp *tmp;
    BvhNode *out;

    static void set_child(BvhNode &n, int which, const Box3 &b, int32_t code)
    {
        if (which == 0) { n.lo0x = b.lo.x; n.lo0y = b.lo.y; n.lo0z = b.lo.z; n.hi0x = b.hi.x; n.hi0y = b.hi.y; n.hi0z = b.hi.z; n.c0 = code; }
        else            { n.lo1x = b.lo.x; n.lo1y = b.lo.y; n.lo1z = b.lo.z; n.hi1x = b.hi.x; n.hi1y = b.hi.y; n.hi1z = b.hi.z; n.c1 = code; }
    }
    /* emits the sub-tree of INNER node t at `slot`; stops (and records) at the nodes listed in `cut` when given */
    void emit(int32_t t, int32_t slot, const std::vector<uint8_t> *is_cut, std::vector<std::pair<int32_t, int32_t>> *cut_out) const
    {
        struct Item { int32_t t, slot; };
        std::vector<Item> st;
        st.push_back({ t, slot });
        while (!st.empty()) {
            const Item it = st.back(); st.pop_back();
            if (is_cut && (*is_cut)[it.t]) { cut_out->push_back({ it.t, it.slot }); continue; }
            const Tmp &N = tmp[it.t];
            const Tmp &L = tmp[N.left], &R = tmp[N.right];
            BvhNode &o = out[it.slot];
            const int32_t ls = it.slot + 1, rs = it.slot + 1 + L.inner;
            set_child(o, 0, L.box, L.left < 0 ? leaf_code(L.first, L.count) : ls);
            set_child(o, 1, R.box, R.left < 0 ? leaf_code(R.first, R.count) : rs);
            o.pad0 = o.pad1 = 0;
            if (R.left >= 0) st.push_back({ N.right, rs });
            if (L.left >= 0) st.push_back({ N.left, ls });
        }
    }
};

/* The 4-wide tree straight from the build tree: a 4-wide node stands for an inner node at EVEN depth, its slots are
 * that node's grand-children (a leaf child stays a slot).  Pre-order layout again: with inner_even known per sub-tree
 * every index is a function of the path, so sub-trees are emitted independently by the threads. */
struct Flattener4 {
    const Tmp *tmp;
    Bvh4Node *out;
    struct Item { int32_t t, slot; };

    /* emits the 4-wide sub-tree of even-depth inner node t at `slot`; sub-trees of at most `defer_below` 4-wide nodes are
     * recorded in `deferred` instead of being emitted (0 = emit everything) */
    void emit(int32_t t, int32_t slot, int32_t defer_below, std::vector<Item> *deferred) const
    {
        std::vector<Item> st;
        st.push_back({ t, slot });
        while (!st.empty()) {
            const Item it = st.back(); st.pop_back();
            if (deferred && tmp[it.t].inner_even <= defer_below) { deferred->push_back(it); continue; }
            const Tmp &N = tmp[it.t];
            int32_t g[4]; int ng = 0;
            const int32_t kids[2] = { N.left, N.right };
            for (int k = 0; k < 2; ++k) {
                const Tmp &C = tmp[kids[k]];
                if (C.left >= 0) { g[ng++] = C.left; g[ng++] = C.right; } else g[ng++] = kids[k];
            }
            Bvh4Node node;
            int32_t next = it.slot + 1;
            Item push[4]; int np = 0;
            for (int j = 0; j < 4; ++j) {
                node.pad[j] = 0;
                if (j >= ng) {
                    node.lox[j] = node.loy[j] = node.loz[j] = FMAXV; node.hix[j] = node.hiy[j] = node.hiz[j] = -FMAXV;
                    node.c[j] = BVH4_EMPTY;
                    continue;
                }
                const Tmp &G = tmp[g[j]];
                node.lox[j] = G.box.lo.x; node.loy[j] = G.box.lo.y; node.loz[j] = G.box.lo.z;
                node.hix[j] = G.box.hi.x; node.hiy[j] = G.box.hi.y; node.hiz[j] = G.box.hi.z;
                if (G.left >= 0) { node.c[j] = next; push[np++] = { g[j], next }; next += G.inner_even; }
                else node.c[j] = leaf_code(G.first, G.count);
            }
            out[it.slot] = node;
            for (int j = np - 1; j >= 0; --j) st.push_back(push[j]);
        }
    }
};

} // namespace

/* Uninitialised work array backed by transparent huge pages where the kernel offers them on request (THP "madvise" mode):
 * the builder touches ~150 MB of fresh memory per 1 M triangles, and with 4 KiB pages the page faults are a visible part
 * of the build.  Falls back silently to ordinary pages. */
template <class T> struct HugeArray {
    T *p = nullptr;
    explicit HugeArray(size_t n) { if (n) p = static_cast<T *>(lb_big_alloc(n * sizeof(T))); }
    ~HugeArray() { free(p); }
    HugeArray(const HugeArray &) = delete;
    HugeArray &operator=(const HugeArray &) = delete;
    T *get() const { return p; }
    T &operator[](size_t i) const { return p[i]; }
};

void *lb_big_alloc(size_t n)
{
    const bool no_huge = getenv("LTR_NO_HUGEPAGES") != nullptr;      /* read per call: tests flip it */
    const bool poison = getenv("LTR_POISON_BIGALLOC") != nullptr;    /* tests: fresh pages are zero, which would hide an element nobody writes */
    void *q = nullptr;
    if (n >= (4u << 20) && !no_huge) {
        const size_t bytes = ((n + (2u << 20) - 1) >> 21) << 21;
        if (posix_memalign(&q, 2u << 20, bytes) == 0) {
#ifdef MADV_HUGEPAGE
            madvise(q, bytes, MADV_HUGEPAGE);
#endif
            if (poison) memset(q, 0xA5, bytes);
            return q;
        }
    }
    q = malloc(n ? n : 1);
    if (!q) { fprintf(stderr, "lighter_b200: out of host memory (%zu bytes)\n", n); abort(); }
    if (poison) memset(q, 0xA5, n ? n : 1);
    return q;
}

void build_scene_bvh(const float *tris9, size_t count, SceneBvh &out, int leaf_max, int threads)
{
    out.nodes.clear();
    out.order.resize(count);
    out.bounds.lo = mk3(FMAXV); out.bounds.hi = mk3(-FMAXV);
    if (leaf_max < 1) leaf_max = 1;
    if (leaf_max > 7) leaf_max = 7;
    const Box3 empty = { mk3(FMAXV), mk3(-FMAXV) };

    if (count == 0) {
        BvhNode n;
        Flattener::set_child(n, 0, empty, leaf_code(0, 0));
        Flattener::set_child(n, 1, empty, leaf_code(0, 0));
        n.pad0 = n.pad1 = 0;
        out.nodes.push_back(n);
        out.depth = 1;
        build_bvh4(out);
        return;
    }

    const bool trace = getenv("LTR_TRACE_BVH") != nullptr;
    auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tt = tnow();
    auto lap = [&](const char *w) { if (trace) { double t = tnow(); fprintf(stderr, "[bvh] %-12s %7.1f ms\n", w, t - tt); tt = t; } };
    if (threads <= 0) { const char *e = getenv("LTR_BVH_THREADS"); threads = e ? atoi(e) : 0; }      /* env: measurements */
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    /* uninitialised working arrays: zero-filling ~140 MB on one thread costs more than the top of the build */
    /* one team of threads for the whole build (only for scenes big enough to take the cooperative path) */
    std::unique_ptr<ThreadTeam> team(threads > 1 && count > COOP_MIN ? new ThreadTeam(threads) : nullptr);
    struct TeamScope { ThreadTeam *prev; explicit TeamScope(ThreadTeam *t) : prev(tl_team) { tl_team = t; } ~TeamScope() { tl_team = prev; } } team_scope(team.get());
    HugeArray<Prim> prim(count), scratch(threads > 1 && count > COOP_MIN ? count : 0);
    if (!prim.get() || (threads > 1 && count > COOP_MIN && !scratch.get())) { fprintf(stderr, "lighter_b200: out of host memory in the BVH build\n"); abort(); }
    {
        std::vector<Box3> part(threads, Box3{ mk3(FMAXV), mk3(-FMAXV) });
        Prim *pp = prim.get();
        parallel_chunks(count > COOP_MIN ? threads : 1, count, [&](int t, size_t b, size_t e) {
            Box3 acc = { mk3(FMAXV), mk3(-FMAXV) };
            for (size_t i = b; i < e; ++i) {
                const float *q = tris9 + 9 * i;
                V3 a = mk3(q[0], q[1], q[2]), bb = mk3(q[3], q[4], q[5]), c = mk3(q[6], q[7], q[8]);
                pp[i].b.lo = min3(a, min3(bb, c));
                pp[i].b.hi = max3(a, max3(bb, c));
                pp[i].id = (uint32_t)i;
                grow(acc, pp[i].b);
            }
            part[t] = acc;
        });
        for (int t = 0; t < threads; ++t) grow(out.bounds, part[t]);
    }

    lap("prims");
    Builder B;
    HugeArray<Tmp> tmp(2 * count + 1);
    if (!tmp.get()) { fprintf(stderr, "lighter_b200: out of host memory in the BVH build\n"); abort(); }
    B.prim = prim.get(); B.scratch = scratch.get(); B.tmp = tmp.get();
    B.leaf_max = leaf_max;
    B.threads = threads;
    B.run((uint32_t)count);
    lap("build");
    const int32_t root = 0;
    {
        const Prim *pp = prim.get();
        uint32_t *oo = out.order.data();
        parallel_chunks(count > COOP_MIN ? threads : 1, count, [=](int, size_t b, size_t e) { for (size_t i = b; i < e; ++i) oo[i] = pp[i].id; });
    }

    if (tmp[root].left < 0) {                   /* whole scene fits one leaf: wrap it */
        BvhNode n;
        Flattener::set_child(n, 0, tmp[root].box, leaf_code(0, (uint32_t)count));
        Flattener::set_child(n, 1, empty, leaf_code(0, 0));
        n.pad0 = n.pad1 = 0;
        out.nodes.push_back(n);
        out.depth = 1;
        build_bvh4(out);
        return;
    }
    out.nodes.resize((size_t)tmp[root].inner);
    out.depth = tmp[root].height;
    Flattener F{ tmp.get(), out.nodes.data() };
    /* the cooperative top of the tree serially (a few dozen nodes), cutting at the phase-B roots; those in parallel */
    std::vector<uint8_t> is_cut;
    std::vector<std::pair<int32_t, int32_t>> cuts;
    if (B.subtrees.size() > 1) {
        is_cut.assign((size_t)B.next.load(), 0);
        for (const Work &w : B.subtrees) if (tmp[w.node].left >= 0) is_cut[w.node] = 1;
        F.emit(root, 0, &is_cut, &cuts);
        std::atomic<size_t> cursor{0};
        auto worker = [&]() { for (size_t i; (i = cursor.fetch_add(1)) < cuts.size();) F.emit(cuts[i].first, cuts[i].second, nullptr, nullptr); };
        parallel_workers((int)std::min<size_t>((size_t)threads, cuts.size()), worker);
    } else {
        F.emit(root, 0, nullptr, nullptr);
    }
    lap("flatten");
    {
        out.nodes4.resize((size_t)tmp[root].inner_even);
        Flattener4 F4{ tmp.get(), out.nodes4.data() };
        std::vector<Flattener4::Item> tasks;
        if (threads > 1 && tmp[root].inner_even > 8192) {
            F4.emit(root, 0, tmp[root].inner_even / (threads * 8) + 64, &tasks);
            std::atomic<size_t> cursor{0};
            auto worker = [&]() { for (size_t i; (i = cursor.fetch_add(1)) < tasks.size();) F4.emit(tasks[i].t, tasks[i].slot, 0, nullptr); };
            parallel_workers((int)std::min<size_t>((size_t)threads, tasks.size()), worker);
        } else {
            F4.emit(root, 0, 0, nullptr);
        }
    }
    lap("flatten4");
}


void build_bvh4(SceneBvh &bvh)
{
    const BigVec<BvhNode> &n2 = bvh.nodes;
    BigVec<Bvh4Node> &n4 = bvh.nodes4;
    n4.clear();
    if (n2.empty()) return;
    n4.reserve(n2.size() / 2 + 1);
    struct Slot { Box3 b; int32_t code; };                 /* code: binary-tree child code (>=0 inner index in n2, <0 leaf) */
    auto child = [&](const BvhNode &n, int k) {
        Slot s;
        if (k == 0) { s.b.lo = mk3(n.lo0x, n.lo0y, n.lo0z); s.b.hi = mk3(n.hi0x, n.hi0y, n.hi0z); s.code = n.c0; }
        else        { s.b.lo = mk3(n.lo1x, n.lo1y, n.lo1z); s.b.hi = mk3(n.hi1x, n.hi1y, n.hi1z); s.code = n.c1; }
        return s;
    };
    /* iterative pre-order: (binary node to expand, slot in n4) */
    struct Item { int32_t n2i, n4i; };
    std::vector<Item> st;
    n4.push_back(Bvh4Node());
    st.push_back({ 0, 0 });
    while (!st.empty()) {
        const Item it = st.back(); st.pop_back();
        Slot slots[4];
        int ns = 0;
        for (int k = 0; k < 2; ++k) {
            const Slot c = child(n2[it.n2i], k);
            if (c.code >= 0) { slots[ns++] = child(n2[c.code], 0); slots[ns++] = child(n2[c.code], 1); }
            else slots[ns++] = c;
        }
        Bvh4Node node;
        int32_t kids[4] = { -1, -1, -1, -1 };
        for (int j = 0; j < 4; ++j) {
            if (j < ns && !(slots[j].code < 0 && ((~slots[j].code) & 7) == 0)) {
                node.lox[j] = slots[j].b.lo.x; node.loy[j] = slots[j].b.lo.y; node.loz[j] = slots[j].b.lo.z;
                node.hix[j] = slots[j].b.hi.x; node.hiy[j] = slots[j].b.hi.y; node.hiz[j] = slots[j].b.hi.z;
                if (slots[j].code >= 0) { kids[j] = slots[j].code; node.c[j] = (int32_t)n4.size(); n4.push_back(Bvh4Node()); }
                else node.c[j] = slots[j].code;
            } else {
                node.lox[j] = node.loy[j] = node.loz[j] = FMAXV; node.hix[j] = node.hiy[j] = node.hiz[j] = -FMAXV;
                node.c[j] = BVH4_EMPTY;
            }
            node.pad[j] = 0;
        }
        n4[it.n4i] = node;
        for (int j = 3; j >= 0; --j) if (kids[j] >= 0) st.push_back({ kids[j], node.c[j] });
    }
}
