// miro_bvh.h — host-side acceleration-structure build of the product (kept on the host, as the
// reference keeps BVH::build, src/BVH.cpp:457-575).  A binned-SAH binary BVH is built top-down and
// collapsed into the 4-wide, 128-byte node layout of include/miro_gpu.h (the GPU counterpart of
// the reference's QBVH_Node collapse, src/BVH.cpp:100-389); object splits compete with spatial splits (SBVH).  The tree topology is NOT required to
// equal the reference's: closest-hit results do not depend on it.  Unlike the reference's builder
// it has no degenerate-axis NaN bin (src/BVH.cpp:714-730): zero-extent sets fall back to a median split.
#pragma once
#include <stdint.h>
#include <vector>
#include "../../include/miro_gpu.h"

namespace miro {

struct BuildPrim {
    float lo[3], hi[3];
    uint32_t kind;     // MIRO_GPU_KIND_*
    uint32_t index;    // caller's index of this primitive within its kind
};

struct BvhStats {
    uint32_t nodes = 0, leaves = 0, max_depth = 0;
    uint32_t references = 0;         // leaf slots: > the primitive count when spatial splits duplicated references
    double sah_cost = 0.0;
};

// Builds one BVH over `prims`, APPENDS its nodes to `nodes`, and appends, per kind, the caller's
// primitive indices in leaf order to `order[kind]`: a leaf reference (kind, first, count) produced
// here means order[kind][first .. first+count).  Returns the child-style root reference (a node
// index, or a leaf reference when prims.size() <= MIRO_GPU_MAX_LEAF, or MIRO_GPU_CHILD_EMPTY).
// tri_verts (optional): the static triangles' vertices, indexed by BuildPrim::index — enables spatial splits, after which a
// triangle may appear in SEVERAL leaves (order[MIRO_GPU_KIND_TRI] then lists it more than once; the caller's gather by `order`
// duplicates it, and a hit on either copy reports the same caller identity).
int32_t build_wide_bvh(const std::vector<BuildPrim>& prims, std::vector<miro_gpu_node>& nodes,
                       std::vector<uint32_t> order[3], BvhStats* stats = nullptr, const miro_gpu_tri* tri_verts = nullptr);

}  // namespace miro
