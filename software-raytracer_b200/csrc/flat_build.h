// flat_build.h - host-built FLAT two-level accelerator for small scenes (<= 255 primitives).
//
// Why: for the bundled scenes (58-68 objects) a per-ray BVH loop spends most of its issue slots with
// 13 of 32 lanes active (profiles/r1b_summary_regen_bvh.txt): every lane waits for the longest
// traversal in its warp. The flat accelerator keeps the control flow of a warp uniform:
//
//   level 1  every lane runs the SAME loop over K cluster boxes + the cubes' boxes (slab tests, FMA)
//            and over the S "single" spheres (conservative sphere cull, FMA) -> a bit mask of hit
//            clusters and a small per-lane queue of candidate primitives;
//   level 2  loop over the lane's hit clusters: 8 conservative sphere culls each -> more candidates;
//   level 3  loop over the lane's candidates: the STRICT reference arithmetic (rt_device.cuh
//            sphere_d2 / box_hit, -fmad=false) with the reference's tie rule.
//
// Like the BVH, levels 1 and 2 never decide a hit: they only produce a conservative candidate set.
//
// CONSERVATIVE SPHERE CULL (proof sketch; u = 2^-24, all norms Euclidean).
// The reference (Object.hpp:115-127) reports a hit iff NOT(d2_ref > r2) with
//   L = fl(C - O), tc = |fl(L.d)|, P = fl(fl(d*tc) + O), Q = fl(P - C), d2_ref = fl(Q.Q).
// Rounding: |Q - Q_real| <= e := u (2 tc |d| + |O| + |Q|), so a reported hit implies
//   |Q_real|^2 <= r2 + D,  D = 4u r2 + 2 r e + e^2,  e <= u (3|O|max + 2|C| + r)(1 + 16u)
// (tc <= |O| + |C|; |Q| ~ r at the decision boundary). In real arithmetic, for either sign of L.d,
//   |Q_real|^2 >= |L|^2 - b^2 - 26u |L|^2        (b = any FMA evaluation of L.d; |d|^2 = 1 + O(u)).
// (with the centre behind the origin, b < 0, the abs(tc) quirk gives |Q_real|^2 = |L|^2 + 3 b^2, hence
// |Q_real|^2 >= |L|^2 + b^2 - 26u |L|^2 there: the device uses the signed square b|b| for both cases).
// The device evaluates v = b|b| + fma(-kappa, LL, R') with LL = FMA chain of L.L, whose own
// rounding is below 5u (LL + R'). With kappa = 1 - 64u and R' = (r2 + 4 D)(1 + 8u) a reported hit
// therefore implies v >= 0; the sphere is skipped only when v < 0. |O|max is the largest norm of any
// ray origin: the camera, or a point on a surface (within 2r of a sphere centre, inside a cube's
// bounds) displaced by eps along the normal - computed here from the scene and the caller's origin
// extent. Preconditions (checked by the builder, else `usable` is false): every coordinate finite
// and below 1e15 in magnitude, at most 255 primitives.
//
// BOXES (cluster bounds and cubes) use the BVH's inflated slab test with the same inflation rule
// (bvh_build.h kInflate): 4e-6 x the largest coordinate magnitude of scene bounds and origins.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/rt_b200.h"

namespace rtb {

constexpr int kFlatClusterSize = 8;      // sphere slots per cluster (padded with never-hit dummies)
constexpr int kFlatClusterStride = 9;    // cull records per cluster in memory: 8 + 1 pad, so that lanes reading the same slot of
                                         // DIFFERENT clusters hit different shared-memory banks (stride 144 B instead of 128 B)
constexpr int kFlatMaxPrims = 255;       // candidate codes are bytes
constexpr int kFlatMaxClusters = 32;     // level-1 cluster boxes: one mask word
constexpr int kFlatMaxLevel1 = 56;       // cubes + single spheres: they enter the 64-entry candidate queue before any cluster

struct HostFlat {
    bool usable = false;
    // level-1 boxes: 2 float4 each: (lo.xyz, 0) (hi.xyz, 0). Entries [0, n_clusters) are sphere clusters
    // (cluster k owns cull slots [8k, 8k+8)), entries [n_clusters, n_clusters + n_cubes) are the cubes in
    // cube-slot order.
    std::vector<float> boxes;            // 8 floats per box
    // conservative sphere records (cx, cy, cz, R'): kFlatClusterStride records per cluster (8 slots + pad), then the singles
    std::vector<float> cull;             // 4 floats per record
    std::vector<uint8_t> cull_slot;      // sphere slot (index into the exact sphere list): 8 per cluster (no pad; 255 = dummy), then the singles
    std::vector<int32_t> prim_id;        // candidate code (sphere slot, or n_spheres + cube slot) -> object id
    int n_clusters = 0, n_cubes = 0, n_singles = 0;
    float kappa = 1.f;
    float inflate_abs = 0.f, extent = 0.f;
};

// objects: the scene in list order; origin_extent: largest |coordinate| of any ray origin outside the scene.
// origin_offset: |rt_params.eps|, how far a secondary origin may sit off its surface (Raytracer.cpp:177).
void build_flat(const std::vector<rt_object>& objects, float origin_extent, HostFlat& out, float origin_offset = 0.f);

}  // namespace rtb
