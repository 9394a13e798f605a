// miro_bvh.cpp — binned-SAH binary build + collapse to the 4-wide GPU node layout.  See miro_bvh.h.
#include "miro_bvh.h"
#include <math.h>
#include <stdlib.h>
#include <float.h>
#include <algorithm>
#include <numeric>

namespace miro {
namespace {

constexpr int kBins = 16;
static uint32_t kMaxLeaf = MIRO_GPU_MAX_LEAF;   // tuning aid: MIRO_BVH_MAX_LEAF (the ABI's leaf reference holds up to 8)
static float kTraversalCost = 1.0f;      // one binary split level, in units of one triangle test (tunable: MIRO_BVH_TRAVERSAL_COST)
constexpr float kPrimCost = 1.0f;

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; } }
    void grow(const float* l, const float* h) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); } }
    void grow(const Box& b) { grow(b.lo, b.hi); }
    void growPoint(const float* p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0.f || dy < 0.f || dz < 0.f) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct BinNode {
    Box box;
    int32_t left = -1, right = -1;   // children (binary nodes), or
    uint32_t first = 0, count = 0;   // leaf range in the index array
    bool leaf() const { return left < 0; }
};

struct Builder {
    const std::vector<BuildPrim>& prims;
    std::vector<uint32_t> idx;
    std::vector<float> cx, cy, cz;   // centroids
    std::vector<BinNode> bn;
    uint32_t max_depth = 0;

    explicit Builder(const std::vector<BuildPrim>& p) : prims(p) {
        const size_t n = p.size();
        idx.resize(n); std::iota(idx.begin(), idx.end(), 0u);
        cx.resize(n); cy.resize(n); cz.resize(n);
        for (size_t i = 0; i < n; ++i) {
            cx[i] = 0.5f * (p[i].lo[0] + p[i].hi[0]); cy[i] = 0.5f * (p[i].lo[1] + p[i].hi[1]); cz[i] = 0.5f * (p[i].lo[2] + p[i].hi[2]);
        }
        bn.reserve(n ? 2 * n : 1);
    }
    float cen(uint32_t i, int axis) const { return axis == 0 ? cx[i] : (axis == 1 ? cy[i] : cz[i]); }

    bool homogeneous(uint32_t first, uint32_t count) const {
        for (uint32_t i = 1; i < count; ++i) if (prims[idx[first + i]].kind != prims[idx[first]].kind) return false;
        return true;
    }

    int32_t build(uint32_t first, uint32_t count, uint32_t depth) {
        max_depth = std::max(max_depth, depth);
        const int32_t me = (int32_t)bn.size();
        bn.emplace_back();
        Box box; box.reset();
        Box cbox; cbox.reset();
        for (uint32_t i = 0; i < count; ++i) {
            const uint32_t p = idx[first + i];
            box.grow(prims[p].lo, prims[p].hi);
            const float c[3] = {cx[p], cy[p], cz[p]};
            cbox.growPoint(c);
        }
        bn[me].box = box;
        const bool homog = homogeneous(first, count);
        if (count == 1 || (count <= 2 && homog)) { bn[me].first = first; bn[me].count = count; return me; }

        uint32_t mid = 0;
        bool have_split = false;
        if (!homog && count <= kMaxLeaf) {
            // a would-be leaf with mixed primitive kinds: separate the kinds (leaves are homogeneous)
            const uint32_t k0 = prims[idx[first]].kind;
            auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](uint32_t p) { return prims[p].kind == k0; });
            mid = (uint32_t)(it - idx.begin()) - first;
            have_split = mid > 0 && mid < count;
        }
        if (!have_split) {
            // binned SAH over the three axes
            float best_cost = FLT_MAX; int best_axis = -1, best_bin = -1;
            const float parent_area = std::max(box.area(), 1e-30f);
            for (int axis = 0; axis < 3; ++axis) {
                const float lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
                if (!(ext > 0.f)) continue;
                Box bb[kBins]; uint32_t bc[kBins];
                for (int b = 0; b < kBins; ++b) { bb[b].reset(); bc[b] = 0; }
                const float scale = kBins / ext;
                for (uint32_t i = 0; i < count; ++i) {
                    const uint32_t p = idx[first + i];
                    int b = (int)((cen(p, axis) - lo) * scale);
                    b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                    bb[b].grow(prims[p].lo, prims[p].hi); bc[b]++;
                }
                float right_area[kBins]; uint32_t right_cnt[kBins];
                Box acc; acc.reset(); uint32_t cnt = 0;
                for (int b = kBins - 1; b > 0; --b) { acc.grow(bb[b]); cnt += bc[b]; right_area[b] = acc.area(); right_cnt[b] = cnt; }
                acc.reset(); cnt = 0;
                for (int b = 0; b < kBins - 1; ++b) {
                    acc.grow(bb[b]); cnt += bc[b];
                    if (cnt == 0 || right_cnt[b + 1] == 0) continue;
                    const float cost = kTraversalCost + kPrimCost * (acc.area() * cnt + right_area[b + 1] * right_cnt[b + 1]) / parent_area;
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
                }
            }
            if (count <= kMaxLeaf && homog && !(best_cost < kPrimCost * count)) {
                bn[me].first = first; bn[me].count = count; return me;   // a leaf is cheaper
            }
            if (best_axis >= 0 && depth < 40) {   // beyond 40 levels fall through to balanced median splits
                const float lo = cbox.lo[best_axis], scale = kBins / (cbox.hi[best_axis] - cbox.lo[best_axis]);
                auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](uint32_t p) {
                    int b = (int)((cen(p, best_axis) - lo) * scale);
                    b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                    return b <= best_bin;
                });
                mid = (uint32_t)(it - idx.begin()) - first;
                have_split = mid > 0 && mid < count;
            }
            if (!have_split) {
                // all centroids coincide (or a degenerate partition): median split in index order
                if (count <= kMaxLeaf && homog) { bn[me].first = first; bn[me].count = count; return me; }
                int axis = 0; float e = -1.f;
                for (int k = 0; k < 3; ++k) if (box.hi[k] - box.lo[k] > e) { e = box.hi[k] - box.lo[k]; axis = k; }
                mid = count / 2;
                std::nth_element(idx.begin() + first, idx.begin() + first + mid, idx.begin() + first + count,
                                 [&](uint32_t a, uint32_t b) { return cen(a, axis) < cen(b, axis); });
            }
        }
        const int32_t l = build(first, mid, depth + 1);
        const int32_t r = build(first + mid, count - mid, depth + 1);
        bn[me].left = l; bn[me].right = r;
        return me;
    }
};

struct Collapser {
    Builder& b;
    std::vector<miro_gpu_node>& nodes;
    std::vector<uint32_t>* order;
    BvhStats st;

    int32_t leaf_ref(const BinNode& n) {
        const uint32_t kind = b.prims[b.idx[n.first]].kind;
        const uint32_t first = (uint32_t)order[kind].size();
        for (uint32_t i = 0; i < n.count; ++i) order[kind].push_back(b.prims[b.idx[n.first + i]].index);
        st.leaves++;
        return MIRO_GPU_LEAF(kind, first, n.count);
    }

    int32_t emit(int32_t bi, uint32_t depth) {
        const BinNode& n = b.bn[bi];
        if (n.leaf()) return leaf_ref(n);
        st.max_depth = std::max(st.max_depth, depth + 1);
        // gather up to four children: repeatedly open the inner child with the largest surface area
        int32_t kids[4]; int nk = 0;
        kids[nk++] = n.left; kids[nk++] = n.right;
        while (nk < 4) {
            int best = -1; float best_area = -1.f;
            for (int i = 0; i < nk; ++i) {
                const BinNode& c = b.bn[kids[i]];
                if (!c.leaf() && c.box.area() > best_area) { best_area = c.box.area(); best = i; }
            }
            if (best < 0) break;
            const BinNode& c = b.bn[kids[best]];
            kids[best] = c.left; kids[nk++] = c.right;
        }
        const int32_t me = (int32_t)nodes.size();
        nodes.emplace_back();
        st.nodes++;
        miro_gpu_node out;
        for (int i = 0; i < 4; ++i) {
            out.lo_x[i] = out.lo_y[i] = out.lo_z[i] = FLT_MAX;
            out.hi_x[i] = out.hi_y[i] = out.hi_z[i] = -FLT_MAX;
            out.child[i] = MIRO_GPU_CHILD_EMPTY; out.reserved[i] = 0;
        }
        const float parent_area = std::max(n.box.area(), 1e-30f);
        for (int i = 0; i < nk; ++i) {
            const BinNode& c = b.bn[kids[i]];
            out.lo_x[i] = c.box.lo[0]; out.lo_y[i] = c.box.lo[1]; out.lo_z[i] = c.box.lo[2];
            out.hi_x[i] = c.box.hi[0]; out.hi_y[i] = c.box.hi[1]; out.hi_z[i] = c.box.hi[2];
            st.sah_cost += (c.leaf() ? kPrimCost * c.count : kTraversalCost) * c.box.area() / parent_area;
        }
        for (int i = 0; i < nk; ++i) out.child[i] = emit(kids[i], depth + 1);
        nodes[me] = out;
        return me;
    }
};

}  // namespace

int32_t build_wide_bvh(const std::vector<BuildPrim>& prims, std::vector<miro_gpu_node>& nodes,
                       std::vector<uint32_t> order[3], BvhStats* stats) {
    if (prims.empty()) return MIRO_GPU_CHILD_EMPTY;
    if (const char* e = getenv("MIRO_BVH_MAX_LEAF")) { const int v = atoi(e); if (v >= 1 && v <= 8) kMaxLeaf = (uint32_t)v; }
    if (const char* e = getenv("MIRO_BVH_TRAVERSAL_COST")) { const float v = (float)atof(e); if (v > 0.f) kTraversalCost = v; }
    Builder b(prims);
    const int32_t root = b.build(0, (uint32_t)prims.size(), 0);
    Collapser c{b, nodes, order, BvhStats()};
    const int32_t ref = c.emit(root, 0);
    if (stats) *stats = c.st;
    return ref;
}

}  // namespace miro
