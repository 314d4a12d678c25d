// flat_build.cpp - host builder of the flat two-level accelerator. See flat_build.h for the contract.
#include "flat_build.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

#include "bvh_build.h"     // kInflate

namespace rtb {
namespace {

struct Sph { float c[3]; float r2f; double r; int slot; };

inline float up(double v) {              // smallest float >= v
    float f = (float)v;
    return (double)f < v ? std::nextafterf(f, INFINITY) : f;
}

// Recursive median split along the longest centroid axis until a group has <= kFlatClusterSize members.
void split_groups(std::vector<Sph>& s, int first, int count, std::vector<std::pair<int, int>>& groups) {
    if (count <= kFlatClusterSize) { groups.emplace_back(first, count); return; }
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = first; i < first + count; ++i)
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], s[(size_t)i].c[k]); hi[k] = std::max(hi[k], s[(size_t)i].c[k]); }
    int axis = 0;
    for (int k = 1; k < 3; ++k) if (hi[k] - lo[k] > hi[axis] - lo[axis]) axis = k;
    // keep the left half a multiple of the cluster size so clusters come out full
    int half = ((count / 2 + kFlatClusterSize - 1) / kFlatClusterSize) * kFlatClusterSize;
    if (half >= count) half = count / 2;
    std::nth_element(s.begin() + first, s.begin() + first + half, s.begin() + first + count,
                     [axis](const Sph& a, const Sph& b) { return a.c[axis] < b.c[axis] || (a.c[axis] == b.c[axis] && a.slot < b.slot); });
    split_groups(s, first, half, groups);
    split_groups(s, first + half, count - half, groups);
}

}  // namespace

void build_flat(const std::vector<rt_object>& objects, float origin_extent, HostFlat& out, float origin_offset) {
    out = HostFlat();
    const double u = std::ldexp(1.0, -24);
    std::vector<Sph> sph;
    struct Cube { float lo[3], hi[3]; };
    std::vector<Cube> cubes;
    double coord_max = std::fabs((double)origin_extent);      // largest |coordinate| (box inflation, as in the BVH)
    double omax = std::sqrt(3.0) * std::fabs((double)origin_extent);   // largest origin NORM
    bool finite = std::isfinite(origin_extent);
    int n_sph = 0, n_box = 0;
    std::vector<int32_t> sph_ids, box_ids;
    for (size_t oi = 0; oi < objects.size(); ++oi) {
        const rt_object& o = objects[oi];
        if (o.type == RT_OBJ_SPHERE) sph_ids.push_back((int32_t)oi);
        else if (o.type == RT_OBJ_CUBE) box_ids.push_back((int32_t)oi);
        if (o.type == RT_OBJ_SPHERE) {
            Sph s;
            const float r2f = o.radius * o.radius;                     // squaredRadius as the exact test sees it
            for (int k = 0; k < 3; ++k) { s.c[k] = o.pos[k]; finite = finite && std::isfinite(o.pos[k]); }
            finite = finite && std::isfinite(r2f);
            s.r2f = r2f; s.r = std::sqrt((double)r2f); s.slot = n_sph++;
            const double cn = std::sqrt((double)s.c[0] * s.c[0] + (double)s.c[1] * s.c[1] + (double)s.c[2] * s.c[2]);
            omax = std::max(omax, cn + 2.0 * s.r);
            for (int k = 0; k < 3; ++k) coord_max = std::max(coord_max, std::fabs((double)s.c[k]) + s.r);
            sph.push_back(s);
        } else if (o.type == RT_OBJ_CUBE) {
            Cube c;
            double n2 = 0;
            for (int k = 0; k < 3; ++k) {
                const float h = std::fabs(o.half[k]);
                c.lo[k] = o.pos[k] - h; c.hi[k] = o.pos[k] + h;
                finite = finite && std::isfinite(c.lo[k]) && std::isfinite(c.hi[k]);
                const double m = std::fabs((double)o.pos[k]) + h;
                n2 += m * m; coord_max = std::max(coord_max, m);
            }
            omax = std::max(omax, std::sqrt(n2));
            cubes.push_back(c); ++n_box;
        }
    }
    if (!finite || !(std::fabs((double)origin_offset) < 1e15) || coord_max > 1e15 || n_sph + n_box > kFlatMaxPrims || n_sph + n_box == 0) return;
    const double off = std::isfinite(origin_offset) ? std::fabs((double)origin_offset) : 1e30;
    omax = omax * (1.0 + 1e-4) + 1e-2 + off;                           // eps offset along the normal, slack
    coord_max += off;
    out.extent = (float)coord_max;
    out.inflate_abs = kInflate * std::max((float)coord_max, 1e-3f);
    out.kappa = (float)(1.0 - 64.0 * u);                               // 1 - 2^-18, exact in float

    // big spheres (the ground, lights) would blow up any cluster box: level-1 singles
    std::vector<double> radii;
    for (const Sph& s : sph) radii.push_back(s.r);
    double r_med = 0;
    if (!radii.empty()) { std::nth_element(radii.begin(), radii.begin() + radii.size() / 2, radii.end()); r_med = radii[radii.size() / 2]; }
    std::vector<Sph> small, singles;
    for (const Sph& s : sph) (s.r > 4.0 * r_med ? singles : small).push_back(s);
    std::vector<std::pair<int, int>> groups;
    if (!small.empty()) split_groups(small, 0, (int)small.size(), groups);
    std::vector<std::pair<int, int>> clusters;
    for (auto g : groups) {
        if (g.second <= 2) for (int i = 0; i < g.second; ++i) singles.push_back(small[(size_t)g.first + i]);
        else clusters.push_back(g);
    }
    if ((int)clusters.size() > kFlatMaxClusters || (int)singles.size() + n_box > kFlatMaxLevel1) return;
    std::sort(singles.begin(), singles.end(), [](const Sph& a, const Sph& b) { return a.slot < b.slot; });

    auto cull_record = [&](const Sph& s) {
        const double cn = std::sqrt((double)s.c[0] * s.c[0] + (double)s.c[1] * s.c[1] + (double)s.c[2] * s.c[2]);
        const double e = u * (3.0 * omax + 2.0 * cn + s.r) * (1.0 + 16.0 * u);
        const double D = 4.0 * u * (double)s.r2f + 2.0 * s.r * e + e * e;
        out.cull.push_back(s.c[0]); out.cull.push_back(s.c[1]); out.cull.push_back(s.c[2]);
        out.cull.push_back(up(((double)s.r2f + 4.0 * D) * (1.0 + 8.0 * u)));
        out.cull_slot.push_back((uint8_t)s.slot);
    };
    const float inf = out.inflate_abs;
    auto push_box = [&](const float* lo, const float* hi) {
        for (int k = 0; k < 3; ++k) out.boxes.push_back(lo[k] - inf);
        out.boxes.push_back(0.f);
        for (int k = 0; k < 3; ++k) out.boxes.push_back(hi[k] + inf);
        out.boxes.push_back(0.f);
    };
    for (auto g : clusters) {
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = 0; i < kFlatClusterSize; ++i) {
            if (i < g.second) {
                const Sph& s = small[(size_t)g.first + i];
                const float r = (float)std::fabs(std::sqrt((double)s.r2f)) * (1.f + 1e-6f);
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], s.c[k] - r); hi[k] = std::max(hi[k], s.c[k] + r); }
                cull_record(s);
            } else {                                                   // dummy: v = b*b - kappa*LL - 1e30 < 0 always
                out.cull.push_back(0.f); out.cull.push_back(0.f); out.cull.push_back(0.f); out.cull.push_back(-1e30f);
                out.cull_slot.push_back(255);
            }
        }
        for (int pad = kFlatClusterSize; pad < kFlatClusterStride; ++pad) {   // bank-conflict padding, never read as a slot
            out.cull.push_back(0.f); out.cull.push_back(0.f); out.cull.push_back(0.f); out.cull.push_back(-1e30f);
        }
        push_box(lo, hi);
    }
    for (const Cube& c : cubes) push_box(c.lo, c.hi);
    for (const Sph& s : singles) cull_record(s);
    out.prim_id = sph_ids;
    out.prim_id.insert(out.prim_id.end(), box_ids.begin(), box_ids.end());
    out.n_clusters = (int)clusters.size(); out.n_cubes = n_box; out.n_singles = (int)singles.size();
    out.usable = true;
}

}  // namespace rtb
