/* reftree.cpp -- see reftree.h.  Host only. */
#include "reftree.h"

#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <future>
#include <thread>

namespace {

/* A sub-tree built on its own: node 0 is its root, child links and item offsets are relative to its own arrays.  The
 * reference appends nodes and items in depth-first order (node, its kept items, the whole first sub-tree, the second child,
 * its sub-tree: lighter_math.cpp:695-781), so a tree is the concatenation root + left + right with the offsets shifted --
 * which lets the two halves of a big node be built by different threads and still come out byte for byte as the serial
 * recursion would write them (pinned by tests/test_host.py against the reference's own dump and by the pre-pass fingerprints). */
struct Sub {
    std::vector<RefNode> nodes;
    std::vector<int32_t> items;
};

struct KeyId { float k; int32_t id; };

struct Builder {
    const Box3 *boxes;
    std::vector<float> vol;
    std::vector<float> key[3];           /* dot(axis, centre) per item, evaluated like the reference */
    std::atomic<int> spare_threads{0};   /* threads that may still be started for sub-trees */
    size_t fork_min = 16384;             /* a node with fewer items to push down is not worth a thread (LTR_REFTREE_FORK_MIN: tests) */

    static float volume(const Box3 &b) { return (b.hi.x - b.lo.x) * (b.hi.y - b.lo.y) * (b.hi.z - b.lo.z); }

    static void emit_items(Sub &S, int32_t node, const int32_t *ids, size_t n)
    {
        S.nodes[node].ido = (int32_t)S.items.size();
        S.items.push_back((int32_t)n);
        S.items.insert(S.items.end(), ids, ids + n);
    }

    /* appends the sub-tree T (built on its own) to S; returns the index its root got */
    static int32_t splice(Sub &S, const Sub &T)
    {
        const int32_t nb = (int32_t)S.nodes.size(), ib = (int32_t)S.items.size();
        S.nodes.insert(S.nodes.end(), T.nodes.begin(), T.nodes.end());
        for (size_t i = nb; i < S.nodes.size(); ++i) {
            if (S.nodes[i].ch != -1) S.nodes[i].ch += nb;
            if (S.nodes[i].ido != -1) S.nodes[i].ido += ib;
        }
        S.items.insert(S.items.end(), T.items.begin(), T.items.end());
        return nb;
    }

    /* builds the sub-tree of ids[0..n) into S, whose node `node` already exists (pushed by the caller) */
    void make(Sub &S, int32_t node, const int32_t *ids, size_t n, int depth)
    {
        Box3 bb = { mk3(3.402823466e+38f), mk3(-3.402823466e+38f) };
        for (size_t i = 0; i < n; ++i) { bb.lo = min3(bb.lo, boxes[ids[i]].lo); bb.hi = max3(bb.hi, boxes[ids[i]].hi); }
        S.nodes[node].lo = bb.lo;
        S.nodes[node].hi = bb.hi;
        S.nodes[node].ch = -1;
        S.nodes[node].ido = -1;
        const float nvol = volume(bb);

        if (n > 4 && depth < 16) {
            /* union of the boxes small enough to be pushed down */
            Box3 sb = { mk3(3.402823466e+38f), mk3(-3.402823466e+38f) };
            size_t nsplit = 0;
            for (size_t i = 0; i < n; ++i)
                if (vol[ids[i]] * 3 < nvol) { ++nsplit; sb.lo = min3(sb.lo, boxes[ids[i]].lo); sb.hi = max3(sb.hi, boxes[ids[i]].hi); }
            if (nsplit >= 4) {
                V3 ext = sb.hi - sb.lo;
                int axis = 2;
                if (ext.x > ext.y && ext.x > ext.z) axis = 0;
                else if (ext.y > ext.x && ext.y > ext.z) axis = 1;

                /* The reference sorts the ids with std::sort and a comparator that looks the keys up; sorting (key, id)
                 * records with the same comparison runs the same introsort over the same outcomes -- the same permutation,
                 * ties included -- without a cache miss per comparison. */
                const float *k = key[axis].data();
                std::vector<int32_t> keep;
                std::vector<KeyId> down;
                down.reserve(nsplit);
                for (size_t i = 0; i < n; ++i) {
                    if (vol[ids[i]] * 3 < nvol) down.push_back(KeyId{ k[ids[i]], ids[i] });
                    else keep.push_back(ids[i]);
                }
                if (!keep.empty()) emit_items(S, node, keep.data(), keep.size());
                std::sort(down.begin(), down.end(), [](const KeyId &a, const KeyId &b) { return a.k < b.k; });
                std::vector<int32_t> sorted(down.size());
                for (size_t i = 0; i < down.size(); ++i) sorted[i] = down[i].id;
                std::vector<KeyId>().swap(down);
                const size_t mid = sorted.size() / 2;

                /* big node and a thread to spare: the second half is built on its own while this thread does the first */
                bool forked = false;
                Sub right;
                std::future<void> fut;
                if (sorted.size() >= fork_min) {
                    int s = spare_threads.load();
                    while (s > 0 && !spare_threads.compare_exchange_weak(s, s - 1)) {}
                    if (s > 0) {
                        forked = true;
                        right.nodes.push_back(RefNode());
                        fut = std::async(std::launch::async, [this, &right, &sorted, mid, depth]() {
                            make(right, 0, sorted.data() + mid, sorted.size() - mid, depth + 1);
                            spare_threads.fetch_add(1);
                        });
                    }
                }
                S.nodes.push_back(RefNode());
                make(S, (int32_t)S.nodes.size() - 1, sorted.data(), mid, depth + 1);
                if (forked) {
                    fut.get();
                    S.nodes[node].ch = splice(S, right);
                } else {
                    S.nodes[node].ch = (int32_t)S.nodes.size();
                    S.nodes.push_back(RefNode());
                    make(S, (int32_t)S.nodes.size() - 1, sorted.data() + mid, sorted.size() - mid, depth + 1);
                }
                return;
            }
        }
        emit_items(S, node, ids, n);
    }
};

} // namespace

void RefTree::build(const Box3 *boxes, size_t count, int threads)
{
    nodes.clear();
    items.clear();
    RefNode root;
    root.lo = mk3(3.402823466e+38f);
    root.hi = mk3(-3.402823466e+38f);
    root.ch = -1;
    root.ido = -1;
    nodes.push_back(root);
    if (count == 0) return;

    Builder B;
    B.boxes = boxes;
    B.spare_threads = threads > 1 ? threads - 1 : 0;
    if (const char *e = getenv("LTR_REFTREE_FORK_MIN")) B.fork_min = (size_t)std::max(2L, atol(e));
    B.vol.resize(count);
    for (int a = 0; a < 3; ++a) B.key[a].resize(count);
    std::vector<int32_t> ids;
    ids.reserve(count);
    for (size_t i = 0; i < count; ++i) {
        const Box3 &b = boxes[i];
        B.vol[i] = Builder::volume(b);
        V3 c = (b.lo + b.hi) * 0.5f;
        B.key[0][i] = 1.f * c.x + 0.f * c.y + 0.f * c.z;
        B.key[1][i] = 0.f * c.x + 1.f * c.y + 0.f * c.z;
        B.key[2][i] = 0.f * c.x + 0.f * c.y + 1.f * c.z;
        if (box_valid(b)) ids.push_back((int32_t)i);
    }
    Sub S;
    S.nodes.swap(nodes);
    B.make(S, 0, ids.data(), ids.size(), 0);
    nodes.swap(S.nodes);
    items.swap(S.items);
}
