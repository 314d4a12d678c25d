// bvh_wide.h - 8-wide BVH with quantised child boxes, collapsed on the host from the binned-SAH BVH2 of bvh_build.h.
//
// Why: scenes whose BVH does not fit shared memory (10 000 spheres, 1 M triangles) are bound by DEPENDENT NODE FETCHES
// through L1/L2 (profiles/r1l_summary_c3_bvh.txt, r1z_summary_c3_wavefront_intersect_v2.txt: long-scoreboard stalls
// 4.2-4.4 per issue, L1 data pipe at 75 % of its peak with four scattered 16-byte loads per BVH2 visit). A wide node
// holds eight children in 80 bytes (10 B per child instead of 32 B), so a ray makes about a third of the fetches and
// moves about a third of the bytes. The layout follows the published compressed wide BVH idea (Ylitie, Karras, Laine
// 2017: per-node origin + per-axis power-of-two scale, 8-bit child planes, children stored in octant order so that
// a front-to-back order falls out of an XOR with the ray's octant) restated for this tracer's contract:
//
// * The BVH still never decides a hit (bvh_build.h): a wide node only yields CANDIDATES, which go through the strict
//   reference intersectors with the reference's tie rule. So the only requirement is conservativeness.
// * Child boxes are the already inflated BVH2 boxes, rounded OUTWARDS onto the node's 8-bit grid (checked in double:
//   origin + scale * qlo <= lo and origin + scale * qhi >= hi hold exactly, build_wide_bvh asserts it per child).
// * The device evaluates a plane as t = fma(32768 + q, s * i, (origin - o) * i - 32768 * s * i) with i = 1/d (MUFU.RCP),
//   where 32768 + q is the float whose bits are 0x47000000 | q << 8 (one PRMT, no integer conversion). Rounding errors:
//   reciprocal 2.4e-7 |t|, (origin - o) and its product with i 1.8e-7 extent |i|, the final fma 1.2e-7 extent |i| - all
//   below the 4e-6 extent the boxes were inflated by (bvh_build.h kInflate, same budget as the BVH2 slab test) - plus
//   the rounding of the constant term, up to 2^-9 grid steps when 32768 s i dominates it; that one is removed explicitly:
//   near planes use the constant minus 2^-8 grid steps, far planes plus 2^-8 (rt_device.cuh wide_node_hits).
// * Zero direction components use i = +-1e18 (the BVH2 loop uses 1e30): with scene extents limited to 1e12 (usable
//   below) every product stays finite, so no NaN can hide a child.
//
// Node, 80 bytes = 5 x uint4:
//   w0  origin x, y, z (float bits); bytes: biased exponent of the x, y, z scale, imask (bit s: slot s is an inner child)
//   w1  first inner child (node index; inner children are consecutive, in slot order), first leaf ref (index into
//       refs; the node's leaf children are consecutive), meta[0..3], meta[4..7]
//       meta of slot s: 0 empty; inner: 0x20 | (24 + s); leaf with n <= 3 refs at offset f < 24: (2^n - 1) << 5 | f
//   w2  qlo.x[0..7] qlo.y[0..7]     w3  qlo.z[0..7] qhi.x[0..7]     w4  qhi.y[0..7] qhi.z[0..7]
// Slot s holds the child lying towards (s&1 ? +x : -x, s&2 ? +y : -y, s&4 ? +z : -z) of the node centre (greedy
// assignment on the projected centroid offsets); a ray visits slots in decreasing (s ^ octinv) with octinv = 7 ^
// (dx<0 | dy<0 << 1 | dz<0 << 2).
#pragma once
#include <cstdint>
#include <vector>

#include "bvh_build.h"

namespace rtb {

struct WideNode { uint32_t w[20]; };
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");

struct HostWideBvh {
    std::vector<WideNode> nodes;     // nodes[0] is the root
    std::vector<int32_t> refs;       // leaf refs in node order (same encoding as HostBvh::refs)
    int depth = 0;                   // levels of wide nodes (root = 1): bound of the traversal stack
    bool usable = false;             // false: a leaf with more than 3 refs or an extent beyond 1e12 - callers keep the BVH2
};

constexpr float kWideMaxExtent = 1e12f;
constexpr int kWideMaxLeaf = 3;

// b2 must have been built with max_leaf <= kWideMaxLeaf (build_bvh); its boxes are already inflated.
void build_wide_bvh(const HostBvh& b2, HostWideBvh& out);

}  // namespace rtb
