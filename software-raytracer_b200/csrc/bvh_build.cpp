// bvh_build.cpp - binned-SAH BVH2 builder (host). See bvh_build.h for the exactness contract.
#include "bvh_build.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <system_error>
#include <thread>

namespace rtb {
namespace {

struct Box3 {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const Box3& b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    void grow_pt(const float* p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
    bool valid() const { return lo[0] <= hi[0]; }
};

struct Prim {
    Box3 box;
    float centroid[3];
    int32_t ref;          // >= 0 sphere slot, < 0 ~cube slot
};

// Nodes and refs of one subtree with subtree-local indices. The sequential builder allocates a node before it descends and
// appends leaf refs as it meets them, i.e. nodes are laid out in pre-order and refs in leaf order; a subtree built on another
// thread is appended to its parent's arrays with its links shifted (append_subtree), which gives exactly that layout again -
// the tree is byte-identical whatever the number of threads.
struct Subtree {
    std::vector<BvhNode> nodes;
    std::vector<int32_t> refs;
    int max_depth = 0;
};

struct Builder {
    std::vector<Prim> prims;
    float eps;

    static constexpr int kBins = 32;           // upper bound; `bins` is what split() uses
    int bins = 16;
    int kMaxLeaf = 4;
    float kTravCost = 1.0f, kPrimCost = 0.7f;

    int32_t make_leaf(int first, int count, Subtree& out) const {
        int start = (int)out.refs.size();
        for (int i = 0; i < count; ++i) out.refs.push_back(prims[(size_t)first + i].ref);
        return ~(int32_t)((uint32_t)start | ((uint32_t)count << 24));
    }

    // appends subtree `s` (whose root link is `link`) to `out`; returns the link as seen from `out`
    static int32_t append_subtree(Subtree& out, const Subtree& s, int32_t link) {
        const int32_t node_off = (int32_t)out.nodes.size();
        const uint32_t ref_off = (uint32_t)out.refs.size();
        auto shift = [&](int32_t l) -> int32_t {
            if (l >= 0) return l + node_off;
            const uint32_t v = (uint32_t)(~l);
            return ~(int32_t)(((v & 0xffffffu) + ref_off) | (v & 0xff000000u));
        };
        out.nodes.reserve(out.nodes.size() + s.nodes.size());
        for (BvhNode n : s.nodes) { n.c[0] = shift(n.c[0]); n.c[1] = shift(n.c[1]); out.nodes.push_back(n); }
        out.refs.insert(out.refs.end(), s.refs.begin(), s.refs.end());
        out.max_depth = std::max(out.max_depth, s.max_depth);
        return shift(link);
    }

    Box3 bounds(int first, int count) const {
        Box3 b;
        for (int i = 0; i < count; ++i) b.grow(prims[(size_t)first + i].box);
        return b;
    }

    void store_child(BvhNode& n, int which, const Box3& b) const {
        float* f = n.f + 6 * which;
        if (!b.valid()) {                       // empty child: NaN planes fail every slab comparison
            for (int k = 0; k < 6; ++k) f[k] = std::nanf("");
            return;
        }
        for (int k = 0; k < 3; ++k) { f[2 * k] = b.lo[k] - eps; f[2 * k + 1] = b.hi[k] + eps; }
    }

    // splits prims[first, first+count) and returns the split position (first < mid < first+count), or -1 for "make a leaf"
    int split(int first, int count, const Box3& nb, bool force) {
        Box3 cb;
        for (int i = 0; i < count; ++i) cb.grow_pt(prims[(size_t)first + i].centroid);
        float best_cost = FLT_MAX; int best_axis = -1, best_bin = -1;
        for (int axis = 0; axis < 3; ++axis) {
            float lo = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0)) continue;
            Box3 bb[kBins]; int bc[kBins] = {0};
            float scale = bins / ext;
            for (int i = 0; i < count; ++i) {
                const Prim& p = prims[(size_t)first + i];
                int b = std::min(bins - 1, std::max(0, (int)((p.centroid[axis] - lo) * scale)));
                bb[b].grow(p.box); bc[b]++;
            }
            float right_area[kBins]; int right_cnt[kBins];
            Box3 acc; int cnt = 0;
            for (int b = bins - 1; b > 0; --b) { acc.grow(bb[b]); cnt += bc[b]; right_area[b] = acc.area(); right_cnt[b] = cnt; }
            acc = Box3(); cnt = 0;
            for (int b = 0; b < bins - 1; ++b) {
                acc.grow(bb[b]); cnt += bc[b];
                if (cnt == 0 || right_cnt[b + 1] == 0) continue;
                float cost = acc.area() * cnt + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        if (best_axis < 0) {                    // all centroids coincide: median split by index if forced
            return (force || count > kMaxLeaf) ? first + count / 2 : -1;
        }
        float leaf_cost = kPrimCost * count;
        float split_cost = kTravCost + kPrimCost * best_cost / std::max(nb.area(), 1e-30f);
        if (!force && count <= kMaxLeaf && leaf_cost <= split_cost) return -1;
        float lo = cb.lo[best_axis], scale = bins / (cb.hi[best_axis] - cb.lo[best_axis]);
        auto mid_it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const Prim& p) {
            int b = std::min(bins - 1, std::max(0, (int)((p.centroid[best_axis] - lo) * scale)));
            return b <= best_bin;
        });
        int mid = (int)(mid_it - prims.begin());
        if (mid == first || mid == first + count) mid = first + count / 2;
        return mid;
    }

    static constexpr int kParallelMinPrims = 1 << 14;      // smaller subtrees are not worth a thread

    // Builds the subtree over prims[first, first+count) into `out` and returns its link (inner index or encoded leaf).
    // par_levels > 0: the two children of a large node are built concurrently (they own disjoint slices of `prims`).
    int32_t build(int first, int count, int depth, Subtree& out, int par_levels) {
        out.max_depth = std::max(out.max_depth, depth);
        if (count <= 0) return make_leaf(first, 0, out);
        const bool depth_left = depth < kMaxBvhDepth - 2;
        if (count == 1 || !depth_left) {
            // depth cap: emit (possibly several) leaves of <= 127 prims chained is not needed in practice;
            // a leaf holds up to 127 refs
            if (count <= 127) return make_leaf(first, count, out);
        }
        Box3 nb = bounds(first, count);
        int mid = split(first, count, nb, count > 127);
        if (mid < 0) return make_leaf(first, count, out);
        int32_t idx = (int32_t)out.nodes.size();
        out.nodes.emplace_back();
        int32_t l, r; Box3 lb, rb;
        build_children(first, mid, first + count, depth, out, par_levels, lb, rb, l, r);
        BvhNode& n = out.nodes[(size_t)idx];
        memset(&n, 0, sizeof n);
        store_child(n, 0, lb); store_child(n, 1, rb);
        n.c[0] = l; n.c[1] = r;
        return idx;
    }

    // children [first, mid) and [mid, end) of a node that already has its slot in out.nodes
    void build_children(int first, int mid, int end, int depth, Subtree& out, int par_levels, Box3& lb, Box3& rb, int32_t& l, int32_t& r) {
        if (par_levels > 0 && end - first >= kParallelMinPrims) {
            Subtree ls, rs;
            int32_t ll = 0, rl = 0;
            auto build_left = [&] { lb = bounds(first, mid - first); ll = build(first, mid - first, depth + 1, ls, par_levels - 1); };
            std::future<void> left;
            try { left = std::async(std::launch::async, build_left); } catch (const std::system_error&) {}   // no thread to be had: build it here
            rb = bounds(mid, end - mid);
            rl = build(mid, end - mid, depth + 1, rs, par_levels - 1);
            if (left.valid()) left.get(); else build_left();
            l = append_subtree(out, ls, ll);
            r = append_subtree(out, rs, rl);
        } else {
            lb = bounds(first, mid - first); rb = bounds(mid, end - mid);
            l = build(first, mid - first, depth + 1, out, 0);
            r = build(mid, end - mid, depth + 1, out, 0);
        }
    }
};

}  // namespace

void build_bvh(const std::vector<rt_object>& objects, float origin_extent, HostBvh& out, int max_leaf, const TriRecords* tris,
               float origin_offset, int threads) {
    out = HostBvh();
    Builder b;
    b.kMaxLeaf = max_leaf < 1 ? 1 : (max_leaf > 64 ? 64 : max_leaf);
    if (const char* v = getenv("RTB200_BVH_BINS")) { const int n = atoi(v); if (n >= 4 && n <= Builder::kBins) b.bins = n; }          // experiments
    if (const char* v = getenv("RTB200_BVH_PRIMCOST")) { const float c = (float)atof(v); if (c > 0.f && c < 100.f) b.kPrimCost = c; }
    int sph_slot = 0, box_slot = 0;
    Box3 scene;
    for (const rt_object& o : objects) {
        Prim p;
        if (o.type == RT_OBJ_SPHERE) {
            float r = std::fabs(o.radius);
            if (!(r == r)) r = 0.f;
            for (int k = 0; k < 3; ++k) { p.box.lo[k] = o.pos[k] - r; p.box.hi[k] = o.pos[k] + r; p.centroid[k] = o.pos[k]; }
            p.ref = sph_slot++;
        } else if (o.type == RT_OBJ_CUBE) {
            for (int k = 0; k < 3; ++k) {
                float h = std::fabs(o.half[k]);
                p.box.lo[k] = o.pos[k] - h; p.box.hi[k] = o.pos[k] + h; p.centroid[k] = o.pos[k];
            }
            p.ref = ~(box_slot++);
        } else continue;
        bool finite = true;
        for (int k = 0; k < 3; ++k) finite = finite && std::isfinite(p.box.lo[k]) && std::isfinite(p.box.hi[k]);
        if (!finite) {                          // non-finite geometry: keep it a candidate for every ray
            for (int k = 0; k < 3; ++k) { p.box.lo[k] = -1e30f; p.box.hi[k] = 1e30f; p.centroid[k] = 0.f; }
        }
        scene.grow(p.box);
        b.prims.push_back(p);
    }
    if (tris) {
        b.prims.reserve(b.prims.size() + (size_t)tris->count());
        for (int t = 0; t < tris->count(); ++t) {
            Prim p;
            const float* bb = tris->bounds.data() + (size_t)6 * t;
            bool finite = true;
            for (int k = 0; k < 3; ++k) {
                p.box.lo[k] = bb[k]; p.box.hi[k] = bb[3 + k]; p.centroid[k] = 0.5f * (bb[k] + bb[3 + k]);
                finite = finite && std::isfinite(bb[k]) && std::isfinite(bb[3 + k]);
            }
            if (!finite) for (int k = 0; k < 3; ++k) { p.box.lo[k] = -1e30f; p.box.hi[k] = 1e30f; p.centroid[k] = 0.f; }
            p.ref = kTriRefBase + t;
            scene.grow(p.box);
            b.prims.push_back(p);
        }
    }
    out.n_prims = (int)b.prims.size();
    float extent = std::fabs(origin_extent);
    if (scene.valid())
        for (int k = 0; k < 3; ++k) extent = std::max(extent, std::max(std::fabs(scene.lo[k]), std::fabs(scene.hi[k])));
    extent += std::isfinite(origin_offset) ? std::fabs(origin_offset) : 1e29f;     // secondary origins sit eps off their surface
    if (!(extent <= 1e29f)) extent = 1e29f;
    out.extent = extent;
    out.inflate_abs = kInflate * std::max(extent, 1e-3f);
    b.eps = out.inflate_abs;

    // the root is always an inner node so the traversal loop has a single entry shape
    Subtree t;
    t.nodes.emplace_back();
    int n = (int)b.prims.size();
    if (n == 0) {
        BvhNode& r = t.nodes[0];
        memset(&r, 0, sizeof r);
        b.store_child(r, 0, Box3()); b.store_child(r, 1, Box3());
        r.c[0] = b.make_leaf(0, 0, t); r.c[1] = r.c[0];
    } else {
        // levels of the tree whose two children are built concurrently: 2^levels tasks at most (0 = sequential; same tree)
        if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("RTB200_BVH_THREADS")) { const int v = atoi(e); if (v >= 1) threads = v; }
        int par_levels = 0;
        while ((1 << par_levels) < threads && par_levels < 6) ++par_levels;
        Box3 nb = b.bounds(0, n);
        int mid = n >= 2 ? b.split(0, n, nb, true) : -1;
        int32_t l, rgt; Box3 lb, rb;
        if (mid < 0) { lb = nb; l = b.build(0, n, 1, t, par_levels); rgt = b.make_leaf(0, 0, t); }
        else b.build_children(0, mid, n, 0, t, par_levels, lb, rb, l, rgt);
        BvhNode& r = t.nodes[0];
        memset(&r, 0, sizeof r);
        b.store_child(r, 0, lb); b.store_child(r, 1, rb);
        r.c[0] = l; r.c[1] = rgt;
    }
    out.nodes.swap(t.nodes); out.refs.swap(t.refs); out.max_depth = t.max_depth;
}

void build_leaf_slots(const HostBvh& bvh, const float* sph4, const int* sph_id, const float* box4, const int* box_id,
                      const TriRecords* tris, std::vector<float>& slots) {
    slots.assign(bvh.refs.size() * 16, 0.f);
    auto put_int = [](float* dst, int32_t v) { memcpy(dst, &v, 4); };
    for (size_t j = 0; j < bvh.refs.size(); ++j) {
        float* w = slots.data() + 16 * j;
        const int32_t r = bvh.refs[j];
        put_int(w + 4, r);
        if (r >= kTriRefBase) {
            const size_t k = (size_t)(r - kTriRefBase);
            const float* t = tris->rec.data() + 12 * k;
            w[0] = t[0]; w[1] = t[1]; w[2] = t[2]; w[3] = t[3];
            put_int(w + 5, tris->obj[k]);
            w[6] = t[4]; w[7] = t[5];
            for (int q = 0; q < 6; ++q) w[8 + q] = t[6 + q];
        } else if (r >= 0) {
            const float* sp = sph4 + 4 * (size_t)r;
            w[0] = sp[0]; w[1] = sp[1]; w[2] = sp[2]; w[3] = sp[3];
            put_int(w + 5, sph_id[r]);
        } else {
            const size_t c = (size_t)(~r);
            const float* bp = box4 + 8 * c;                  // (px, py, pz, 0) (hx, hy, hz, 0)
            w[0] = bp[0]; w[1] = bp[1]; w[2] = bp[2]; w[3] = bp[4];
            put_int(w + 5, box_id[c]);
            w[6] = bp[5]; w[7] = bp[6];
        }
    }
}

// Error budget of the device's decode (rt_bvh_lane.cuh node_step, Q16 branch), per axis, in units of step / |d| (the parameter
// distance of one grid step): a = step * inv with inv = rcp(d) (2 ulp) and one rounding: relative 3.6e-7, times q + 2^23 <= 8.5e6
// would be 3 units - but the same a enters b2 = fma(-2^23, a, b) with the opposite sign, so what remains is the error on q * a
// (q <= 65535: 0.024 units) plus the rounding of b2 itself, |b2| <= 2^23 |a| + |b|: half an ulp = 0.5 |a| + 6e-8 |b| (|b| <= 2 extent
// |inv|, i.e. 1.2e-7 extent / step = 0.008 units for a 65530-step grid) plus the final rounding of t (|t| <= 65536 |a| + ...: 0.004
// units) plus the rounding of b = (org - o) * inv (two roundings and the reciprocal: 4.8e-7 x 2 extent = 0.06 units). Sum < 0.6 units:
// ONE extra unit on each side is enough.
void build_qnodes(const HostBvh& bvh, HostQNodes& out) {
    out = HostQNodes();
    const size_t n = bvh.nodes.size();
    if (n == 0) return;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (const BvhNode& nd : bvh.nodes)
        for (int c = 0; c < 2; ++c)
            for (int k = 0; k < 3; ++k) {
                const double a = nd.f[6 * c + 2 * k], b = nd.f[6 * c + 2 * k + 1];
                if (!(a <= b)) continue;                       // empty child box (inverted): encoded as an empty slab below
                lo[k] = std::min(lo[k], a); hi[k] = std::max(hi[k], b);
            }
    std::vector<float> ext[3];
    for (int k = 0; k < 3; ++k) {
        if (!(lo[k] <= hi[k]) || !std::isfinite(lo[k]) || !std::isfinite(hi[k])) return;
        const double range = std::max(hi[k] - lo[k], 1e-30);
        out.step[k] = (float)(range / 65528.0);
        if (!(out.step[k] > 0.f) || !std::isfinite(out.step[k])) return;
        out.org[k] = (float)(lo[k] - 3.0 * (double)out.step[k]);
        // the float org / step must still cover the range with the margins: q in [2, 65533] before the extra unit
        if ((hi[k] - (double)out.org[k]) / (double)out.step[k] > 65533.0 || (lo[k] - (double)out.org[k]) / (double)out.step[k] < 2.0) return;
    }
    // coarseness check against the leaf-level boxes (children that are leaves): median extent per axis
    for (const BvhNode& nd : bvh.nodes)
        for (int c = 0; c < 2; ++c)
            if (nd.c[c] < 0)
                for (int k = 0; k < 3; ++k) { const float e = nd.f[6 * c + 2 * k + 1] - nd.f[6 * c + 2 * k]; if (e >= 0.f) ext[k].push_back(e); }
    for (int k = 0; k < 3; ++k) {
        if (ext[k].empty()) continue;
        std::nth_element(ext[k].begin(), ext[k].begin() + ext[k].size() / 2, ext[k].end());
        const float med = ext[k][ext[k].size() / 2];
        if (out.step[k] > 0.125f * med) return;               // too coarse: the quantised boxes would be visibly larger than the float ones
    }
    out.words.assign(n * 8, 0u);
    for (size_t i = 0; i < n; ++i) {
        const BvhNode& nd = bvh.nodes[i];
        uint32_t* w = out.words.data() + 8 * i;
        for (int c = 0; c < 2; ++c)
            for (int k = 0; k < 3; ++k) {
                const double a = nd.f[6 * c + 2 * k], b = nd.f[6 * c + 2 * k + 1];
                uint32_t qlo, qhi;
                if (!(a <= b)) { qlo = 65535u; qhi = 0u; }     // empty: lo > hi on every axis -> the slab test fails
                else {
                    const double fl = std::floor((a - (double)out.org[k]) / (double)out.step[k]) - 1.0;
                    const double ce = std::ceil((b - (double)out.org[k]) / (double)out.step[k]) + 1.0;
                    qlo = (uint32_t)std::min(65535.0, std::max(0.0, fl));
                    qhi = (uint32_t)std::min(65535.0, std::max(0.0, ce));
                    // containment with the budget of the decode: a full unit of slack on both sides, checked in double
                    if ((double)out.org[k] + ((double)qlo + 0.7) * (double)out.step[k] > a || (double)out.org[k] + ((double)qhi - 0.7) * (double)out.step[k] < b) { out.words.clear(); return; }
                }
                w[3 * c + k] = qlo | (qhi << 16);
            }
        w[6] = (uint32_t)nd.c[0]; w[7] = (uint32_t)nd.c[1];
    }
    out.usable = true;
}

}  // namespace rtb
