/* reftree.cpp -- see reftree.h.  Host only. */
#include "reftree.h"

#include <algorithm>

namespace {

struct Builder {
    const Box3 *boxes;
    std::vector<float> vol;
    std::vector<float> key[3];           /* dot(axis, centre) per item, evaluated like the reference */
    RefTree *T;

    static float volume(const Box3 &b) { return (b.hi.x - b.lo.x) * (b.hi.y - b.lo.y) * (b.hi.z - b.lo.z); }

    void emit_items(int32_t node, const int32_t *ids, size_t n)
    {
        T->nodes[node].ido = (int32_t)T->items.size();
        T->items.push_back((int32_t)n);
        T->items.insert(T->items.end(), ids, ids + n);
    }

    void make(int32_t node, int32_t *ids, size_t n, int depth)
    {
        Box3 bb = { mk3(3.402823466e+38f), mk3(-3.402823466e+38f) };
        for (size_t i = 0; i < n; ++i) { bb.lo = min3(bb.lo, boxes[ids[i]].lo); bb.hi = max3(bb.hi, boxes[ids[i]].hi); }
        T->nodes[node].lo = bb.lo;
        T->nodes[node].hi = bb.hi;
        T->nodes[node].ch = -1;
        T->nodes[node].ido = -1;
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

                std::vector<int32_t> keep, down;
                down.reserve(nsplit);
                for (size_t i = 0; i < n; ++i) (vol[ids[i]] * 3 < nvol ? down : keep).push_back(ids[i]);
                if (!keep.empty()) emit_items(node, keep.data(), keep.size());

                const float *k = key[axis].data();
                std::sort(down.begin(), down.end(), [k](int32_t a, int32_t b) { return k[a] < k[b]; });
                size_t mid = down.size() / 2;

                T->nodes.push_back(RefNode());
                make((int32_t)T->nodes.size() - 1, down.data(), mid, depth + 1);
                T->nodes[node].ch = (int32_t)T->nodes.size();
                T->nodes.push_back(RefNode());
                make((int32_t)T->nodes.size() - 1, down.data() + mid, down.size() - mid, depth + 1);
                return;
            }
        }
        emit_items(node, ids, n);
    }
};

} // namespace

void RefTree::build(const Box3 *boxes, size_t count)
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
    B.T = this;
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
    B.make(0, ids.data(), ids.size(), 0);
}
