// miro_bvh.cpp — binned-SAH binary build + collapse to the 4-wide GPU node layout.  See miro_bvh.h.
#include "miro_bvh.h"
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <float.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <future>
#include <numeric>
#include <thread>

namespace miro {
namespace {

constexpr int kMaxBins = 64;
static int kBins = 16;                   // SAH bins per axis, object and spatial (MIRO_BVH_BINS, <= kMaxBins)
static size_t kSpatialMinPrims = 512;     // no spatial splits below this many primitives (MIRO_BVH_SPATIAL_MIN)
static float kSpatialAlpha = 1e-5f;      // spatial splits are tried where the object split's children overlap by more than this fraction of the root's area (MIRO_BVH_ALPHA)
static uint32_t kMaxLeaf = MIRO_GPU_MAX_LEAF;   // tuning aid: MIRO_BVH_MAX_LEAF (the ABI's leaf reference holds up to 8)
static float kTraversalCost = 1.0f;      // one binary split level, in units of one triangle test (tunable: MIRO_BVH_TRAVERSAL_COST)
constexpr float kPrimCost = 1.0f;
static double kSpatialBudget = 1.0;      // extra references spatial splits may create, as a fraction of the primitive count (MIRO_BVH_SPATIAL; 0: object splits only)
// Sub-trees of at least this many references are built as parallel tasks (MIRO_BVH_PARALLEL_MIN; 0: sequential build).  The tree does
// not depend on the number of threads: WHETHER a node's children run concurrently is decided by its reference count alone.
static size_t kParallelMin = 8192;

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
    uint32_t first = 0, count = 0;   // leaf range in Builder::leaf_prims
    bool leaf() const { return left < 0; }
};

// One reference to a primitive: the part of it (its box, possibly clipped by earlier spatial splits) a sub-tree is responsible
// for.  Without spatial splits every primitive has exactly one reference.
struct Ref {
    Box box;
    uint32_t prim;                   // index into Builder::prims
};

// Top-down build with object splits (binned SAH over reference centroids) and SPATIAL splits (Stich, Friedrichs, Dietrich 2009):
// where the two children of the best object split overlap, the node's box is also cut by axis-aligned planes, references that
// straddle the chosen plane are split in two (a triangle is clipped against the plane, so both halves get tight boxes), and the
// cheaper of the two kinds of split wins.  Long, thin or large triangles — which make every box they fall into overlap its
// neighbours — end up referenced from several small leaves instead of inflating one.  The number of extra references is capped
// (kSpatialBudget x the primitive count); only nodes that hold static triangles exclusively are split spatially (a duplicated
// instance would be traversed twice).
struct Builder {
    const std::vector<BuildPrim>& prims;
    const miro_gpu_tri* tri_verts;   // vertices of the static triangles (indexed by BuildPrim::index), or NULL: no spatial splits
    // Node and leaf-slot storage is allocated up front and handed out by atomic counters, so concurrently built sub-trees never
    // move each other's nodes.  (Where a node ends up in `bn` depends on timing; the tree it belongs to does not — the collapse
    // walks child links and assigns the final order.)
    std::vector<BinNode> bn;
    std::vector<uint32_t> leaf_prims;     // primitive (index into prims) of every leaf slot, a leaf's slots contiguous
    std::atomic<uint32_t> bn_next{0}, leaf_next{0};
    std::atomic<uint32_t> max_depth{0};
    std::atomic<size_t> extra_refs{0};
    size_t extra_budget = 0;
    float root_area = 1.f;

    Builder(const std::vector<BuildPrim>& p, const miro_gpu_tri* tv) : prims(p), tri_verts(tv) {
        extra_budget = tv ? (size_t)(kSpatialBudget * (double)p.size()) : 0;
        const size_t max_refs = p.size() + extra_budget + 16;      // every split that adds references is paid from the budget
        bn.resize(2 * max_refs);
        leaf_prims.resize(max_refs);
    }
    void note_depth(uint32_t d) { uint32_t cur = max_depth.load(std::memory_order_relaxed); while (d > cur && !max_depth.compare_exchange_weak(cur, d, std::memory_order_relaxed)) {} }

    static float centroid(const Ref& r, int axis) { return 0.5f * (r.box.lo[axis] + r.box.hi[axis]); }

    bool homogeneous(const std::vector<Ref>& refs) const {
        for (size_t i = 1; i < refs.size(); ++i) if (prims[refs[i].prim].kind != prims[refs[0].prim].kind) return false;
        return true;
    }
    bool all_static_triangles(const std::vector<Ref>& refs) const {
        for (const Ref& r : refs) if (prims[r.prim].kind != MIRO_GPU_KIND_TRI) return false;
        return true;
    }

    int32_t make_leaf(int32_t me, const std::vector<Ref>& refs) {
        const uint32_t first = leaf_next.fetch_add((uint32_t)refs.size(), std::memory_order_relaxed);
        if ((size_t)first + refs.size() > leaf_prims.size()) { fprintf(stderr, "miro_bvh: leaf storage exhausted (%u + %zu > %zu)\n", first, refs.size(), leaf_prims.size()); abort(); }
        bn[me].first = first; bn[me].count = (uint32_t)refs.size();
        for (size_t i = 0; i < refs.size(); ++i) leaf_prims[first + i] = refs[i].prim;
        return me;
    }

    // Box of the part of reference r that lies in the slab lo <= x[axis] <= hi.  Static triangles are clipped exactly (the box
    // of the clipped polygon, intersected with the reference's own box); anything else keeps its box cut by the slab.
    Box clip(const Ref& r, int axis, float lo, float hi) const {
        Box out;
        const BuildPrim& bp = prims[r.prim];
        if (tri_verts && bp.kind == MIRO_GPU_KIND_TRI) {
            const miro_gpu_tri& t = tri_verts[bp.index];
            const float* v[3] = {t.v0, t.v1, t.v2};
            out.reset();
            for (int e = 0; e < 3; ++e) {
                const float* a = v[e]; const float* b = v[(e + 1) % 3];
                if (a[axis] >= lo && a[axis] <= hi) out.growPoint(a);
                const float planes[2] = {lo, hi};
                for (float pl : planes) {
                    if ((a[axis] < pl && b[axis] > pl) || (a[axis] > pl && b[axis] < pl)) {
                        const float w = (pl - a[axis]) / (b[axis] - a[axis]);
                        float q[3];
                        for (int k = 0; k < 3; ++k) q[k] = a[k] + w * (b[k] - a[k]);
                        q[axis] = pl;
                        out.growPoint(q);
                    }
                }
            }
            // conservative against rounding of the interpolated points: one ulp-ish pad, then never outside the reference's own box
            for (int k = 0; k < 3; ++k) {
                const float pad = 1e-6f * std::max(fabsf(out.lo[k]), fabsf(out.hi[k]));
                out.lo[k] = std::max(out.lo[k] - pad, r.box.lo[k]); out.hi[k] = std::min(out.hi[k] + pad, r.box.hi[k]);
            }
        } else out = r.box;
        out.lo[axis] = std::max(out.lo[axis], lo); out.hi[axis] = std::min(out.hi[axis], hi);
        return out;
    }

    // `budget`: extra references this sub-tree may still create (in: its share; out: what it did not use).  The share is handed
    // down in proportion to the children's reference counts and what the left child leaves goes to the right one, so the depth-
    // first recursion does not spend the whole allowance in the first corner of the scene it visits.  Children of a node with
    // at least kParallelMin references are built concurrently, each with its own share (nothing is handed across).
    int32_t build(std::vector<Ref>& refs, uint32_t depth, size_t& budget) {
        note_depth(depth);
        const int32_t me = (int32_t)bn_next.fetch_add(1u, std::memory_order_relaxed);
        if ((size_t)me >= bn.size()) { fprintf(stderr, "miro_bvh: node storage exhausted (internal invariant)\n"); abort(); }
        const uint32_t count = (uint32_t)refs.size();
        Box box; box.reset();
        Box cbox; cbox.reset();
        for (const Ref& r : refs) {
            box.grow(r.box);
            const float c[3] = {centroid(r, 0), centroid(r, 1), centroid(r, 2)};
            cbox.growPoint(c);
        }
        bn[me].box = box;
        if (depth == 0) root_area = std::max(box.area(), 1e-30f);
        const bool homog = homogeneous(refs);
        if (count == 1 || (count <= 2 && homog)) return make_leaf(me, refs);

        std::vector<Ref> left, right;
        bool have_split = false;
        if (!homog && count <= kMaxLeaf) {
            // a would-be leaf with mixed primitive kinds: separate the kinds (leaves are homogeneous)
            const uint32_t k0 = prims[refs[0].prim].kind;
            for (const Ref& r : refs) (prims[r.prim].kind == k0 ? left : right).push_back(r);
            have_split = !left.empty() && !right.empty();
        }
        if (!have_split) {
            const float parent_area = std::max(box.area(), 1e-30f);
            // ---- object split: binned SAH over the three axes
            float best_cost = FLT_MAX; int best_axis = -1, best_bin = -1; uint32_t best_nl = 0, best_nr = 0;
            Box best_lbox, best_rbox; best_lbox.reset(); best_rbox.reset();
            for (int axis = 0; axis < 3; ++axis) {
                const float lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
                if (!(ext > 0.f)) continue;
                Box bb[kMaxBins]; uint32_t bc[kMaxBins];
                for (int b = 0; b < kBins; ++b) { bb[b].reset(); bc[b] = 0; }
                const float scale = kBins / ext;
                for (const Ref& r : refs) {
                    int b = (int)((centroid(r, axis) - lo) * scale);
                    b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                    bb[b].grow(r.box); bc[b]++;
                }
                Box right_box[kMaxBins]; uint32_t right_cnt[kMaxBins];
                Box acc; acc.reset(); uint32_t cnt = 0;
                for (int b = kBins - 1; b > 0; --b) { acc.grow(bb[b]); cnt += bc[b]; right_box[b] = acc; right_cnt[b] = cnt; }
                acc.reset(); cnt = 0;
                for (int b = 0; b < kBins - 1; ++b) {
                    acc.grow(bb[b]); cnt += bc[b];
                    if (cnt == 0 || right_cnt[b + 1] == 0) continue;
                    const float cost = kTraversalCost + kPrimCost * (acc.area() * cnt + right_box[b + 1].area() * right_cnt[b + 1]) / parent_area;
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; best_lbox = acc; best_rbox = right_box[b + 1]; best_nl = cnt; best_nr = right_cnt[b + 1]; }
                }
            }
            // ---- spatial split: only where the object split's children overlap noticeably (relative to the ROOT's area)
            float sp_cost = FLT_MAX; int sp_axis = -1; float sp_pos = 0.f; uint32_t sp_nl = 0, sp_nr = 0;
            if (budget > 0 && count > kMaxLeaf && all_static_triangles(refs)) {
                Box ov;
                for (int k = 0; k < 3; ++k) { ov.lo[k] = std::max(best_lbox.lo[k], best_rbox.lo[k]); ov.hi[k] = std::min(best_lbox.hi[k], best_rbox.hi[k]); }
                const bool overlapping = best_axis < 0 || ov.area() / root_area > kSpatialAlpha;
                if (overlapping) for (int axis = 0; axis < 3; ++axis) {
                    const float lo = box.lo[axis], ext = box.hi[axis] - box.lo[axis];
                    if (!(ext > 0.f)) continue;
                    Box bb[kMaxBins]; uint32_t n_in[kMaxBins], n_out[kMaxBins];
                    for (int b = 0; b < kBins; ++b) { bb[b].reset(); n_in[b] = n_out[b] = 0; }
                    const float scale = kBins / ext, width = ext / kBins;
                    for (const Ref& r : refs) {
                        int b0 = (int)((r.box.lo[axis] - lo) * scale), b1 = (int)((r.box.hi[axis] - lo) * scale);
                        b0 = b0 < 0 ? 0 : (b0 >= kBins ? kBins - 1 : b0); b1 = b1 < b0 ? b0 : (b1 >= kBins ? kBins - 1 : b1);
                        if (b0 == b1) bb[b0].grow(r.box);
                        else for (int b = b0; b <= b1; ++b) {
                            const Box c = clip(r, axis, lo + width * b, b == kBins - 1 ? box.hi[axis] : lo + width * (b + 1));
                            if (c.lo[0] <= c.hi[0] && c.lo[1] <= c.hi[1] && c.lo[2] <= c.hi[2]) bb[b].grow(c);
                        }
                        n_in[b0]++; n_out[b1]++;
                    }
                    Box right_box[kMaxBins]; uint32_t right_cnt[kMaxBins];
                    Box acc; acc.reset(); uint32_t cnt = 0;
                    for (int b = kBins - 1; b > 0; --b) { acc.grow(bb[b]); cnt += n_out[b]; right_box[b] = acc; right_cnt[b] = cnt; }
                    acc.reset(); cnt = 0;
                    for (int b = 0; b < kBins - 1; ++b) {
                        acc.grow(bb[b]); cnt += n_in[b];
                        if (cnt == 0 || right_cnt[b + 1] == 0 || cnt == count || right_cnt[b + 1] == count) continue;
                        const float cost = kTraversalCost + kPrimCost * (acc.area() * cnt + right_box[b + 1].area() * right_cnt[b + 1]) / parent_area;
                        if (cost < sp_cost) { sp_cost = cost; sp_axis = axis; sp_pos = lo + width * (b + 1); sp_nl = cnt; sp_nr = right_cnt[b + 1]; }
                    }
                }
            }
            const float split_cost = std::min(best_cost, sp_cost);
            if (count <= kMaxLeaf && homog && !(split_cost < kPrimCost * count)) return make_leaf(me, refs);   // a leaf is cheaper
            if (sp_axis >= 0 && sp_cost < best_cost && depth < 40 && (size_t)(sp_nl + sp_nr - count) <= budget) {
                left.reserve(sp_nl); right.reserve(sp_nr);
                for (const Ref& r : refs) {
                    if (r.box.hi[sp_axis] <= sp_pos) left.push_back(r);
                    else if (r.box.lo[sp_axis] >= sp_pos) right.push_back(r);
                    else {
                        Ref a = r, b = r;
                        a.box = clip(r, sp_axis, -FLT_MAX, sp_pos); b.box = clip(r, sp_axis, sp_pos, FLT_MAX);
                        const bool ok_a = a.box.lo[0] <= a.box.hi[0] && a.box.lo[1] <= a.box.hi[1] && a.box.lo[2] <= a.box.hi[2];
                        const bool ok_b = b.box.lo[0] <= b.box.hi[0] && b.box.lo[1] <= b.box.hi[1] && b.box.lo[2] <= b.box.hi[2];
                        if (ok_a) left.push_back(a);
                        if (ok_b) right.push_back(b);
                        if (!ok_a && !ok_b) left.push_back(r);
                    }
                }
                have_split = !left.empty() && !right.empty() && left.size() < count && right.size() < count;
                // the bins predicted sp_nl + sp_nr references; the partition compares the boxes with sp_pos directly and can find a
                // few more straddlers — a split that would overdraw the allowance is dropped for the object split (the storage
                // is sized from the allowance)
                if (have_split && left.size() + right.size() - count > budget) have_split = false;
                if (have_split) { const size_t used = left.size() + right.size() - count; extra_refs.fetch_add(used, std::memory_order_relaxed); budget -= used; }
                else { left.clear(); right.clear(); }
            }
            if (!have_split && best_axis >= 0 && depth < 40) {   // beyond 40 levels fall through to balanced median splits
                const float lo = cbox.lo[best_axis], scale = kBins / (cbox.hi[best_axis] - cbox.lo[best_axis]);
                left.reserve(best_nl); right.reserve(best_nr);      // the winning bin boundary's counts: no regrowth while partitioning
                for (const Ref& r : refs) {
                    int b = (int)((centroid(r, best_axis) - lo) * scale);
                    b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                    (b <= best_bin ? left : right).push_back(r);
                }
                have_split = !left.empty() && !right.empty();
                if (!have_split) { left.clear(); right.clear(); }
            }
            if (!have_split) {
                // all centroids coincide (or a degenerate partition): median split along the longest axis
                if (count <= kMaxLeaf && homog) return make_leaf(me, refs);
                int axis = 0; float e = -1.f;
                for (int k = 0; k < 3; ++k) if (box.hi[k] - box.lo[k] > e) { e = box.hi[k] - box.lo[k]; axis = k; }
                const uint32_t mid = count / 2;
                std::nth_element(refs.begin(), refs.begin() + mid, refs.end(), [&](const Ref& a, const Ref& b) { return centroid(a, axis) < centroid(b, axis); });
                left.assign(refs.begin(), refs.begin() + mid); right.assign(refs.begin() + mid, refs.end());
            }
        }
        std::vector<Ref>().swap(refs);       // the parent's list is not needed below this point
        const size_t nl = left.size(), nr = right.size();
        size_t b_left = (size_t)((double)budget * (double)nl / (double)(nl + nr));
        const size_t rest = budget - b_left;
        if (kParallelMin > 0 && count >= kParallelMin && depth < 7) {      // at most 2^7 tasks; the rule depends on the node alone
            size_t b_right = rest;
            std::future<int32_t> lf = std::async(std::launch::async, [&]() { return build(left, depth + 1, b_left); });
            const int32_t r = build(right, depth + 1, b_right);
            const int32_t l = lf.get();
            budget = b_left + b_right;
            bn[me].left = l; bn[me].right = r;
            return me;
        }
        const int32_t l = build(left, depth + 1, b_left);
        std::vector<Ref>().swap(left);
        size_t b_right = rest + b_left;          // its own share plus what the left sub-tree did not use
        const int32_t r = build(right, depth + 1, b_right);
        budget = b_right;
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
        const uint32_t kind = b.prims[b.leaf_prims[n.first]].kind;
        const uint32_t first = (uint32_t)order[kind].size();
        for (uint32_t i = 0; i < n.count; ++i) order[kind].push_back(b.prims[b.leaf_prims[n.first + i]].index);
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
                       std::vector<uint32_t> order[3], BvhStats* stats, const miro_gpu_tri* tri_verts) {
    if (prims.empty()) return MIRO_GPU_CHILD_EMPTY;
    if (const char* e = getenv("MIRO_BVH_MAX_LEAF")) { const int v = atoi(e); if (v >= 1 && v <= 8) kMaxLeaf = (uint32_t)v; }
    if (const char* e = getenv("MIRO_BVH_TRAVERSAL_COST")) { const float v = (float)atof(e); if (v > 0.f) kTraversalCost = v; }
    if (const char* e = getenv("MIRO_BVH_SPATIAL_MIN")) { const long v = atol(e); if (v >= 0) kSpatialMinPrims = (size_t)v; }
    if (const char* e = getenv("MIRO_BVH_BINS")) { const int v = atoi(e); if (v >= 4 && v <= kMaxBins) kBins = v; }
    if (const char* e = getenv("MIRO_BVH_ALPHA")) { const float v = (float)atof(e); if (v >= 0.f) kSpatialAlpha = v; }
    if (const char* e = getenv("MIRO_BVH_SPATIAL")) { const double v = atof(e); if (v >= 0.0 && v <= 4.0) kSpatialBudget = v; }
    if (const char* e = getenv("MIRO_BVH_PARALLEL_MIN")) { const long v = atol(e); if (v >= 0) kParallelMin = (size_t)v; }
    // scenes of a few hundred triangles gain nothing from spatial splits (measured: the Cornell box renders 3 % slower with them)
    Builder b(prims, (kSpatialBudget > 0.0 && prims.size() >= kSpatialMinPrims) ? tri_verts : nullptr);
    std::vector<Ref> refs(prims.size());
    for (size_t i = 0; i < prims.size(); ++i) {
        refs[i].prim = (uint32_t)i;
        for (int k = 0; k < 3; ++k) { refs[i].box.lo[k] = prims[i].lo[k]; refs[i].box.hi[k] = prims[i].hi[k]; }
    }
    size_t budget = b.extra_budget;
    static const bool timing = getenv("MIRO_HOST_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    const int32_t root = b.build(refs, 0, budget);
    const auto t1 = std::chrono::steady_clock::now();
    Collapser c{b, nodes, order, BvhStats()};
    const int32_t ref = c.emit(root, 0);
    if (timing) fprintf(stderr, "[miro_host] build_wide_bvh %zu prims: binary build %.1f ms, collapse %.1f ms\n", prims.size(),
                        std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    c.st.references = b.leaf_next.load();
    if (stats) *stats = c.st;
    return ref;
}

}  // namespace miro
