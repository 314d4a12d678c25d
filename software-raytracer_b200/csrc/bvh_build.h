// bvh_build.h - host-built BVH over the scene's object list (north star: "a host-built BVH for
// large primitive counts, verified hit-for-hit against the reference's brute-force traversal").
//
// The BVH never decides a hit. It only produces a CONSERVATIVE candidate set: every object the
// reference's own intersectors (Object.hpp:104-141,173-200) could report as a valid hit closer
// than the current best is guaranteed to be visited, and each candidate is then tested with the
// same strict-IEEE code the brute-force loop uses, with the reference's tie rule (lowest object
// id wins on equal distance, Raytracer.cpp:127-137). Conservativeness comes from inflating every
// box by kInflate * (largest coordinate magnitude of the scene bounds and the ray origins), two
// orders of magnitude above the float rounding error of either computation.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/rt_b200.h"
#include "mesh.h"

namespace rtb {

// 64-byte node, both children's boxes stored in the parent:
//   f[0..5]  child0 lo.x hi.x lo.y hi.y lo.z hi.z      f[6..11] child1 (same order)
//   c[0], c[1] child links: >= 0 inner node index; < 0 leaf: ~link = first | (count << 24)
struct BvhNode {
    float f[12];
    int32_t c[2];
    int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

struct HostBvh {
    std::vector<BvhNode> nodes;      // nodes[0] is the root (always an inner node)
    std::vector<int32_t> refs;       // leaf entries: [0, kTriRefBase) sphere slot, >= kTriRefBase triangle, < 0 ~cube slot
    int max_depth = 0;
    int n_prims = 0;
    float inflate_abs = 0.f;         // the absolute inflation that was applied
    float extent = 0.f;              // largest |coordinate| covered by the inflation bound
    // traversal statistics hooks are on the device side
};

constexpr float kInflate = 4e-5f;    // relative box inflation
constexpr int32_t kTriRefBase = 0x40000000;   // leaf ref of triangle i = kTriRefBase + i
constexpr int kMaxBvhDepth = 40;     // traversal stack entries per thread

// objects: the scene in list order. origin_extent: largest |coordinate| of any ray origin that will be
// traced from outside the scene bounds (the camera position); secondary origins lie on surfaces.
void build_bvh(const std::vector<rt_object>& objects, float origin_extent, HostBvh& out, int max_leaf = 4,
               const TriRecords* tris = nullptr);

}  // namespace rtb
