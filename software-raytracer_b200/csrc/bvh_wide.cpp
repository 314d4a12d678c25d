// bvh_wide.cpp - collapse of the BVH2 into 8-wide nodes with quantised child boxes (host). Contract in bvh_wide.h.
#include "bvh_wide.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace rtb {
namespace {

struct Item {
    int32_t link;                    // BVH2 link: >= 0 inner node, < 0 encoded leaf
    float lo[3], hi[3];              // the (inflated) box the BVH2 parent stores for it
    float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    int leaf_count() const { return (int)((uint32_t)(~link) >> 24); }
    int leaf_first() const { return (int)((uint32_t)(~link) & 0xffffffu); }
};

struct Collapser {
    const HostBvh& b2;
    HostWideBvh& out;
    bool ok = true;

    // children of BVH2 node n that can hold anything (empty leaves and NaN boxes are dropped)
    void children_of(int32_t n, std::vector<Item>& items) const {
        const BvhNode& nd = b2.nodes[(size_t)n];
        for (int c = 0; c < 2; ++c) {
            Item it;
            it.link = nd.c[c];
            bool valid = true;
            for (int k = 0; k < 3; ++k) {
                it.lo[k] = nd.f[6 * c + 2 * k]; it.hi[k] = nd.f[6 * c + 2 * k + 1];
                valid = valid && it.lo[k] <= it.hi[k];          // false for NaN
            }
            if (!valid) continue;
            if (it.link < 0 && it.leaf_count() == 0) continue;
            items.push_back(it);
        }
    }

    void fill(size_t widx, int32_t n2, int level) {
        out.depth = std::max(out.depth, level);
        std::vector<Item> items;
        children_of(n2, items);
        // greedy collapse: open the inner child with the largest surface until eight slots are used
        for (;;) {
            if (items.size() >= 8) break;
            int best = -1; float best_area = -1.f;
            for (size_t i = 0; i < items.size(); ++i)
                if (items[i].link >= 0 && items[i].area() > best_area) { best_area = items[i].area(); best = (int)i; }
            if (best < 0) break;
            std::vector<Item> sub;
            children_of(items[(size_t)best].link, sub);
            if (items.size() - 1 + sub.size() > 8) break;
            items.erase(items.begin() + best);
            items.insert(items.end(), sub.begin(), sub.end());
        }
        WideNode node;
        memset(&node, 0, sizeof node);
        uint8_t* qb = reinterpret_cast<uint8_t*>(&node.w[8]);   // 48 plane bytes: qlo.x qlo.y qlo.z qhi.x qhi.y qhi.z, 8 each
        for (int k = 0; k < 3; ++k) for (int s = 0; s < 8; ++s) { qb[8 * k + s] = 255; qb[24 + 8 * k + s] = 0; }   // empty slots: inverted
        if (items.empty()) { out.nodes[widx] = node; return; }

        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) { lo[k] = items[0].lo[k]; hi[k] = items[0].hi[k]; }
        for (const Item& it : items) for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], it.lo[k]); hi[k] = std::max(hi[k], it.hi[k]); }
        for (int k = 0; k < 3; ++k) if (!(std::fabs(lo[k]) <= kWideMaxExtent && std::fabs(hi[k]) <= kWideMaxExtent)) ok = false;
        if (!ok) { out.nodes[widx] = node; return; }

        // slot assignment: slot bits say on which side of the node centre the child lies
        int slot_of[8]; bool slot_used[8] = {false}; bool item_done[8] = {false};
        double cost[8][8];
        for (size_t i = 0; i < items.size(); ++i)
            for (int s = 0; s < 8; ++s) {
                double c = 0.0;
                for (int k = 0; k < 3; ++k) {
                    const double off = 0.5 * ((double)items[i].lo[k] + items[i].hi[k]) - 0.5 * ((double)lo[k] + hi[k]);
                    c += ((s >> k) & 1) ? off : -off;
                }
                cost[i][s] = c;
            }
        for (size_t round = 0; round < items.size(); ++round) {
            int bi = -1, bs = -1; double bc = -1e300;
            for (size_t i = 0; i < items.size(); ++i) {
                if (item_done[i]) continue;
                for (int s = 0; s < 8; ++s) if (!slot_used[s] && cost[i][s] > bc) { bc = cost[i][s]; bi = (int)i; bs = s; }
            }
            slot_of[bi] = bs; slot_used[bs] = true; item_done[bi] = true;
        }
        int item_at[8];
        for (int s = 0; s < 8; ++s) item_at[s] = -1;
        for (size_t i = 0; i < items.size(); ++i) item_at[slot_of[i]] = (int)i;

        // quantisation frame: origin = lo (float), per-axis scale 2^ex with (hi - lo) <= 255 * 2^ex
        int ex[3];
        for (int k = 0; k < 3; ++k) {
            const double ext = (double)hi[k] - (double)lo[k];
            int e = -100;
            if (ext > 0.0) { int fe; std::frexp(ext / 255.0, &fe); e = std::max(fe, -100); }   // ext/255 <= 2^fe
            for (;;) {                                       // every child's hi plane must land on the grid
                const double sc = std::ldexp(1.0, e);
                if (std::ceil(ext / sc) <= 255.0) break;
                ++e;
            }
            ex[k] = e;
        }
        uint8_t* hdr = reinterpret_cast<uint8_t*>(&node.w[3]);
        for (int k = 0; k < 3; ++k) { memcpy(&node.w[k], &lo[k], 4); hdr[k] = (uint8_t)(ex[k] + 127); }
        uint8_t* meta = reinterpret_cast<uint8_t*>(&node.w[6]);

        uint8_t imask = 0;
        int n_inner = 0;
        const size_t prim_base = out.refs.size();
        int prim_off = 0;
        for (int s = 0; s < 8; ++s) {
            const int i = item_at[s];
            if (i < 0) continue;
            const Item& it = items[(size_t)i];
            for (int k = 0; k < 3; ++k) {
                const double sc = std::ldexp(1.0, ex[k]);
                double ql = std::floor(((double)it.lo[k] - (double)lo[k]) / sc);
                double qh = std::ceil(((double)it.hi[k] - (double)lo[k]) / sc);
                ql = std::min(std::max(ql, 0.0), 255.0); qh = std::min(std::max(qh, 0.0), 255.0);
                // conservativeness, exact in double: the decoded box contains the inflated BVH2 box
                if (!((double)lo[k] + sc * ql <= (double)it.lo[k] && (double)lo[k] + sc * qh >= (double)it.hi[k])) ok = false;
                qb[8 * k + s] = (uint8_t)ql; qb[24 + 8 * k + s] = (uint8_t)qh;
            }
            if (it.link >= 0) {
                imask |= (uint8_t)(1u << s);
                meta[s] = (uint8_t)(0x20 | (24 + s));
                ++n_inner;
            } else {
                const int cnt = it.leaf_count();
                if (cnt > kWideMaxLeaf) { ok = false; continue; }
                meta[s] = (uint8_t)((((1u << cnt) - 1u) << 5) | (uint32_t)prim_off);
                for (int j = 0; j < cnt; ++j) out.refs.push_back(b2.refs[(size_t)it.leaf_first() + j]);
                prim_off += cnt;
            }
        }
        hdr[3] = imask;
        const size_t child_base = out.nodes.size();
        node.w[4] = (uint32_t)child_base;
        node.w[5] = (uint32_t)prim_base;
        out.nodes[widx] = node;
        if (!ok) return;
        out.nodes.resize(child_base + (size_t)n_inner);      // inner children: consecutive, in slot order
        size_t next = child_base;
        for (int s = 0; s < 8; ++s) {
            const int i = item_at[s];
            if (i < 0 || items[(size_t)i].link < 0) continue;
            fill(next++, items[(size_t)i].link, level + 1);
            if (!ok) return;
        }
    }
};

}  // namespace

void build_wide_bvh(const HostBvh& b2, HostWideBvh& out) {
    out = HostWideBvh();
    if (b2.nodes.empty() || !(b2.extent <= kWideMaxExtent)) return;
    Collapser c{b2, out};
    out.nodes.resize(1);
    c.fill(0, 0, 1);
    out.usable = c.ok && out.nodes.size() < (size_t)0x7fffffff && out.refs.size() < (size_t)0x7fffffff;
}

}  // namespace rtb
