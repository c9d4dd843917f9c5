/* bvh.cpp -- host builder for the flat scene BVH (see bvh.h): binned SAH over all three axes,
 * parallel over sub-trees, flattened depth-first so that a parent and its first child are adjacent. */
#include "bvh.h"

#include <algorithm>
#include <atomic>
#include <thread>

namespace {

const float FMAXV = 3.402823466e+38f;

struct Tmp {
    Box3 box;
    int32_t left, right;       /* -1 for a leaf */
    uint32_t first, count;
};

inline void grow(Box3 &b, const Box3 &o) { b.lo = min3(b.lo, o.lo); b.hi = max3(b.hi, o.hi); }
inline void grow(Box3 &b, V3 p) { b.lo = min3(b.lo, p); b.hi = max3(b.hi, p); }
inline float half_area(const Box3 &b)
{
    V3 e = b.hi - b.lo;
    return e.x * e.y + e.y * e.z + e.z * e.x;
}
inline float axis_of(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

struct Builder {
    const Box3 *pbox;
    const V3 *pcen;
    uint32_t *idx;
    std::vector<Tmp> tmp;
    std::atomic<int32_t> next{0};
    std::atomic<int> free_threads{0};
    int leaf_max;

    int32_t alloc() { return next.fetch_add(1); }

    void build(int32_t me, uint32_t first, uint32_t count, int depth)
    {
        Box3 bb = { mk3(FMAXV), mk3(-FMAXV) }, cb = bb;
        for (uint32_t i = first; i < first + count; ++i) { grow(bb, pbox[idx[i]]); grow(cb, pcen[idx[i]]); }
        Tmp &N = tmp[me];
        N.box = bb; N.first = first; N.count = count; N.left = N.right = -1;
        if ((int)count <= leaf_max) return;

        const int NB = 16;
        int best_axis = -1, best_split = 0;
        float best_cost = FMAXV;
        for (int a = 0; a < 3; ++a) {
            float lo = axis_of(cb.lo, a), ext = axis_of(cb.hi, a) - lo;
            if (!(ext > 0)) continue;
            Box3 bbx[NB];
            uint32_t cnt[NB];
            for (int b = 0; b < NB; ++b) { bbx[b].lo = mk3(FMAXV); bbx[b].hi = mk3(-FMAXV); cnt[b] = 0; }
            float scale = NB / ext;
            for (uint32_t i = first; i < first + count; ++i) {
                int b = (int)((axis_of(pcen[idx[i]], a) - lo) * scale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                cnt[b]++; grow(bbx[b], pbox[idx[i]]);
            }
            float la[NB]; uint32_t lc[NB];
            Box3 acc = { mk3(FMAXV), mk3(-FMAXV) };
            uint32_t c = 0;
            for (int b = 0; b < NB - 1; ++b) { grow(acc, bbx[b]); c += cnt[b]; la[b] = c ? half_area(acc) : 0; lc[b] = c; }
            acc.lo = mk3(FMAXV); acc.hi = mk3(-FMAXV); c = 0;
            for (int b = NB - 1; b > 0; --b) {
                grow(acc, bbx[b]); c += cnt[b];
                if (lc[b - 1] == 0 || c == 0) continue;
                float cost = la[b - 1] * lc[b - 1] + half_area(acc) * c;
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
            }
        }

        uint32_t mid;
        if (best_axis < 0) {
            mid = first + count / 2;                       /* coincident centroids: split by index */
        } else if (depth > 40) {                           /* failsafe against degenerate SAH chains */
            V3 e = cb.hi - cb.lo;
            int a = (e.x >= e.y && e.x >= e.z) ? 0 : (e.y >= e.z ? 1 : 2);
            const V3 *cen = pcen;
            mid = first + count / 2;
            std::nth_element(idx + first, idx + mid, idx + first + count,
                             [=](uint32_t p, uint32_t q) { return axis_of(cen[p], a) < axis_of(cen[q], a); });
        } else {
            float lo = axis_of(cb.lo, best_axis), scale = NB / (axis_of(cb.hi, best_axis) - lo);
            const V3 *cen = pcen; int a = best_axis, sp = best_split;
            uint32_t *m = std::partition(idx + first, idx + first + count, [=](uint32_t t) {
                int b = (int)((axis_of(cen[t], a) - lo) * scale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                return b < sp;
            });
            mid = (uint32_t)(m - idx);
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        int32_t l = alloc(), r = alloc();
        tmp[me].left = l; tmp[me].right = r;
        uint32_t lc = mid - first, rc = count - lc;
        if (lc > 16384 && rc > 16384 && free_threads.fetch_sub(1) > 0) {
            std::thread th([=]() { build(l, first, lc, depth + 1); });
            build(r, mid, rc, depth + 1);
            th.join();
            free_threads.fetch_add(1);
        } else {
            if (lc > 16384 && rc > 16384) free_threads.fetch_add(1);   /* undo the failed reservation */
            build(l, first, lc, depth + 1);
            build(r, mid, rc, depth + 1);
        }
    }
};

inline int32_t leaf_code(uint32_t first, uint32_t count) { return ~(int32_t)((first << 3) | count); }

struct Flattener {
    const std::vector<Tmp> &tmp;
    std::vector<BvhNode> &out;
    int maxdepth = 0;

    static void set_child(BvhNode &n, int which, const Box3 &b, int32_t code)
    {
        if (which == 0) { n.lo0x = b.lo.x; n.lo0y = b.lo.y; n.lo0z = b.lo.z; n.hi0x = b.hi.x; n.hi0y = b.hi.y; n.hi0z = b.hi.z; n.c0 = code; }
        else            { n.lo1x = b.lo.x; n.lo1y = b.lo.y; n.lo1z = b.lo.z; n.hi1x = b.hi.x; n.hi1y = b.hi.y; n.hi1z = b.hi.z; n.c1 = code; }
    }
    /* iterative DFS: (tmp index of an INNER node, slot in out) */
    void run(int32_t root)
    {
        struct Item { int32_t t, slot, depth; };
        std::vector<Item> st;
        out.push_back(BvhNode());
        st.push_back({ root, 0, 1 });
        while (!st.empty()) {
            Item it = st.back(); st.pop_back();
            if (it.depth > maxdepth) maxdepth = it.depth;
            const Tmp &N = tmp[it.t];
            int32_t kids[2] = { N.left, N.right };
            int32_t slots[2] = { -1, -1 };
            for (int k = 0; k < 2; ++k) {
                const Tmp &C = tmp[kids[k]];
                if (C.left < 0) set_child(out[it.slot], k, C.box, leaf_code(C.first, C.count));
                else { slots[k] = (int32_t)out.size(); out.push_back(BvhNode()); set_child(out[it.slot], k, C.box, slots[k]); }
            }
            out[it.slot].pad0 = out[it.slot].pad1 = 0;
            /* push second child first so the first child is processed (and laid out) next */
            if (slots[1] >= 0) st.push_back({ kids[1], slots[1], it.depth + 1 });
            if (slots[0] >= 0) st.push_back({ kids[0], slots[0], it.depth + 1 });
        }
    }
};

} // namespace

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
        return;
    }

    std::vector<Box3> pbox(count);
    std::vector<V3> pcen(count);
    for (size_t i = 0; i < count; ++i) {
        const float *t = tris9 + 9 * i;
        V3 a = mk3(t[0], t[1], t[2]), b = mk3(t[3], t[4], t[5]), c = mk3(t[6], t[7], t[8]);
        pbox[i].lo = min3(a, min3(b, c));
        pbox[i].hi = max3(a, max3(b, c));
        pcen[i] = (pbox[i].lo + pbox[i].hi) * 0.5f;
        out.order[i] = (uint32_t)i;
        grow(out.bounds, pbox[i]);
    }

    Builder B;
    B.pbox = pbox.data(); B.pcen = pcen.data(); B.idx = out.order.data();
    B.tmp.resize(2 * count + 1);
    B.leaf_max = leaf_max;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    B.free_threads.store(threads > 1 ? threads - 1 : 0);
    int32_t root = B.alloc();
    B.build(root, 0, (uint32_t)count, 0);

    if (B.tmp[root].left < 0) {                 /* whole scene fits one leaf: wrap it */
        BvhNode n;
        Flattener::set_child(n, 0, B.tmp[root].box, leaf_code(0, (uint32_t)count));
        Flattener::set_child(n, 1, empty, leaf_code(0, 0));
        n.pad0 = n.pad1 = 0;
        out.nodes.push_back(n);
        out.depth = 1;
        return;
    }
    Flattener F{ B.tmp, out.nodes };
    out.nodes.reserve(count);
    F.run(root);
    out.depth = F.maxdepth;
}
