// bvh_build.h - host-built BVH over the scene's object list (north star: "a host-built BVH for
// large primitive counts, verified hit-for-hit against the reference's brute-force traversal").
//
// The BVH never decides a hit. It only produces a CONSERVATIVE candidate set: every object the
// reference's own intersectors (Object.hpp:104-141,173-200) could report as a valid hit closer
// than the current best is guaranteed to be visited, and each candidate is then tested with the
// same strict-IEEE code the brute-force loop uses, with the reference's tie rule (lowest object
// id wins on equal distance, Raytracer.cpp:127-137). Conservativeness comes from inflating every
// box by kInflate * (largest coordinate magnitude of the scene bounds and the ray origins), several
// times the float rounding error of either computation (derivation at kInflate below).
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

// Relative box inflation. What it has to cover (slab test of rt_device.cuh, per axis): the computed plane parameter
// a = fma(plane, i, -o*i) with i = 1/d from MUFU.RCP differs from (plane - o)/d by at most 2.4e-7 |t| (reciprocal,
// 2 ulp) + u (|o| + |plane|) |i| (the two roundings, u = 2^-24); moving a plane outwards by eps shifts its parameter by
// eps |i|, and |t| / |i| is the distance travelled along that axis <= |o| + |plane| <= 2 extent. So eps >= 6e-7 x
// extent is enough; 4e-6 keeps a factor 6 on top (and the reference's own sphere test accepts points at most
// ~1e-7 x extent outside the sphere, flat_build.h). The first version used 4e-5, which made the boxes of the
// 0.2-radius spheres of Scene1 (extent 2001 because of the ground sphere) 0.08 larger per side than needed.
constexpr float kInflate = 4e-6f;
constexpr int32_t kTriRefBase = 0x40000000;   // leaf ref of triangle i = kTriRefBase + i
constexpr int kMaxBvhDepth = 40;     // traversal stack entries per thread

// objects: the scene in list order. origin_extent: largest |coordinate| of any ray origin that will be
// traced from outside the scene bounds (the camera position); secondary origins lie on surfaces.
// threads: large subtrees are built concurrently (0 = all hardware threads, 1 = sequential; RTB200_BVH_THREADS overrides). The tree
// is byte-identical for every thread count (bvh_build.cpp Subtree).
void build_bvh(const std::vector<rt_object>& objects, float origin_extent, HostBvh& out, int max_leaf = 4,
               const TriRecords* tris = nullptr, float origin_offset = 0.f, int threads = 0);   // origin_offset: |rt_params.eps|

// Leaf-ordered primitive SLOTS for BVHs traversed from global memory: one 64-byte record per leaf reference, in `refs` order,
// so that a leaf test is a load at (first + i) instead of refs[first + i] -> geometry array -> id array (two dependent scattered
// loads per primitive, three for the id of a hit). 16 words per slot:
//   w0..w3  sphere (cx, cy, cz, r^2) | cube (px, py, pz, hx) | triangle (n.xyz, dn)
//   w4      the leaf reference itself (>= 0 sphere slot, < 0 ~cube slot, >= kTriRefBase triangle): type tag + tie-break key
//   w5      object id
//   w6, w7  cube (hy, hz) | triangle (m1.x, m1.y)
//   w8..w13 triangle (m1.z, k1, m2.x, m2.y, m2.z, k2)          (second half: only triangles read it)
// Spheres and cubes are decided by ONE 256-bit load, triangles by two. The values are the same floats the unordered arrays hold,
// so hits are bit-identical. sph / box: the packed geometry lists of pack_scene() (rt_host_pack.h), ids likewise.
void build_leaf_slots(const HostBvh& bvh, const float* sph4, const int* sph_id, const float* box4, const int* box_id,
                      const TriRecords* tris, std::vector<float>& slots);

// 32-byte nodes with 16-bit child planes on ONE grid per tree: the form the persistent kernels traverse when the tree lives in
// global memory. Why: the lanes of a warp sit in different nodes, so every load instruction of a node fetch costs one L1 wavefront
// PER LANE whatever its width, and the L1 data pipe is what bounds those kernels - a 64-byte node is two
// 256-bit loads, a 32-byte node one.
//   word k (k = 0..2)   child 0, axis k:  lo | hi << 16          word 3 + k   child 1, axis k
//   word 6, 7           the two child links (as in BvhNode)
// plane = org[axis] + q * step[axis]. Encoding (double arithmetic on the float org / step the device gets): lo is rounded DOWN and
// hi UP to the grid and both are moved one more unit outwards, which covers the decode's own rounding: the device evaluates
// t = fma(float(2^23 + q), step/d, (org - o)/d - 2^23 step/d), whose error is below 0.6 step/d (derivation in bvh_build.cpp), i.e.
// the decoded slab always CONTAINS the slab of the float box, which itself is conservative (kInflate). Still only a candidate
// filter: hits are decided by the strict tests. usable = false when the grid would be too coarse for the scene's primitives
// (step above 1/8 of the median primitive box extent on some axis): the float nodes are traversed then.
struct HostQNodes {
    std::vector<uint32_t> words;       // 8 per node, same node indices as HostBvh::nodes
    float org[3] = {0.f, 0.f, 0.f}, step[3] = {1.f, 1.f, 1.f};
    bool usable = false;
};
void build_qnodes(const HostBvh& bvh, HostQNodes& out);

}  // namespace rtb
