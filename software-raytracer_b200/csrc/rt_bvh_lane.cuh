// rt_bvh_lane.cuh - the per-lane BVH traversal state machine of the persistent kernels (rt_wavefront.cu: k_wf_intersect_bvh,
// k_wf_stream) and its shared-memory stack. Depends only on rt_device.cuh, so the CPU tests can compile it for the host
// (tests/host_emu) and drive it with arbitrary schedules.
#pragma once
#include "rt_device.cuh"

namespace rtb {

constexpr int kLaneThreads = 128;      // threads per CTA of the kernels that use LaneStack (= rt_trace.cuh kThreads)

// ---- per-lane BVH traversal state machine of the persistent kernels (rt_wavefront.cu) -------------------------------
// Per-thread traversal stack in the shared-memory region setup_trace() reserves for it (2 x stack_entries x kLaneThreads words),
// addressed with 32-bit shared-window addresses and laid out [entry][link | entry distance][thread]: a push is two stores at
// immediate offsets and one add, and the pointer itself is the only register. The bottom entry is a SENTINEL (no link,
// distance -inf: never stale), so the pop loop needs no emptiness test:
//     do { sp -= entry; t = [sp + kLaneThreads * 4]; } while (t > best_t);   link = [sp];
// popping the sentinel returns kNone and leaves it in place. (The generic-pointer form of the same stack cost ~10 address
// instructions per push or pop and a stack-pointer compare per skipped stale entry.)
#ifdef RTB_HOST_EMULATION
static uint32_t g_emu_lane_stack[64 * 2 * kLaneThreads];       // CPU tests: one lane's stack
#endif
struct LaneStack {
    static constexpr uint32_t kEntry = 8u * kLaneThreads, kDist = 4u * kLaneThreads;     // bytes
    static constexpr int kNone = (int)0x80000000;            // "no link": never a real leaf (a leaf's count is <= 127)
    uint32_t sp;                                             // shared-window byte address of the next free entry
#ifdef RTB_HOST_EMULATION
    static void sts(uint32_t a, uint32_t v) { g_emu_lane_stack[a / 4] = v; }
    static uint32_t lds(uint32_t a) { return g_emu_lane_stack[a / 4]; }
    void init(const void*) { sp = 0; push(kNone, __int_as_float((int)0xff800000)); }
#else
    __device__ __forceinline__ static void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
    __device__ __forceinline__ static uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
    __device__ __forceinline__ void init(const void* thread_slot) {
        sp = (uint32_t)__cvta_generic_to_shared(thread_slot);
        push(kNone, __int_as_float((int)0xff800000));        // sentinel
    }
#endif
    __device__ __forceinline__ void push(int link, float t_entry) {
        sts(sp, (uint32_t)link); sts(sp + kDist, __float_as_uint(t_entry));
        sp += kEntry;
    }
    // next subtree that the line does not enter behind the best hit; kNone when the stack is empty
    __device__ __forceinline__ int pop(float best_t) {
        float t;
        do { sp -= kEntry; t = __uint_as_float(lds(sp + kDist)); } while (t > best_t);
        const int link = (int)lds(sp);
        if (link == kNone) sp += kEntry;                     // the sentinel stays
        return link;
    }
};

// One ray's traversal, advanced in steps so that a WARP can decide what its lanes do next ("speculative while-while"):
//   node_step()  one inner-node visit: both children's slab tests, far child pushed with its entry distance. A lane that
//                reaches a leaf STASHES it (leaf0) and keeps descending / popping - culling against a best distance that is
//                merely not yet as tight as it could be, so the candidate set stays conservative;
//   leaf_step()  the stashed leaf (and a second one waiting in `cur`) tested with the strict reference arithmetic.
// Any interleaving of the two ends with the same hit as closest_hit_bvh(): same candidate semantics, strict tests and tie
// rule (tests/test_device_logic_cpu.py drives this struct on the CPU with random schedules).
struct BvhLane {
    enum { IDLE = 0, ACTIVE = 1, DONE = 2 };
    static constexpr int NONE = LaneStack::kNone;
    LaneStack ls;
    int state, cur, leaf0;
    float ix, iy, iz, ox, oy, oz;
    uint32_t snx, sny, snz, sfx, sfy, sfz;                   // quantised nodes: PRMT selectors of the NEAR / FAR plane per axis (by the direction's sign)
    float best_t; int best_id, best_ref; bool have;
    float3 bn;

    __device__ __forceinline__ void init(const void* stack_slot) {
        ls.init(stack_slot);
        state = IDLE; cur = NONE; leaf0 = NONE;
        ix = iy = iz = ox = oy = oz = 0.f; best_t = 0.f; best_id = 0; best_ref = 0; have = false; bn = f3(0.f, 0.f, 0.f);
        snx = sny = snz = 0x5410u; sfx = sfy = sfz = 0x5432u;
    }
    // start at the root; the stack is empty (the previous traversal ended by popping the sentinel).
    // Plane parameters are t = fma(P, ix, ox). Float nodes: P = the plane's coordinate, ix = 1/d, ox = -o/d. Quantised nodes (q16,
    // bvh_build.h HostQNodes): P = the float 2^23 + q built from the 16-bit plane by one PRMT, ix = step/d, ox = (org - o)/d - 2^23 step/d.
    __device__ __forceinline__ void begin(float3 o, float3 d, bool q16 = false, float3 q_org = f3(0.f, 0.f, 0.f), float3 q_step = f3(1.f, 1.f, 1.f)) {
        const float big = 1e30f;
        ix = fabsf(d.x) > 1e-30f ? RTB_FAST_RCP(d.x) : copysignf(big, d.x);
        iy = fabsf(d.y) > 1e-30f ? RTB_FAST_RCP(d.y) : copysignf(big, d.y);
        iz = fabsf(d.z) > 1e-30f ? RTB_FAST_RCP(d.z) : copysignf(big, d.z);
        if (q16) {
            ox = (q_org.x - o.x) * ix; oy = (q_org.y - o.y) * iy; oz = (q_org.z - o.z) * iz;
            ix *= q_step.x; iy *= q_step.y; iz *= q_step.z;
            ox = fmaf(-8388608.f, ix, ox); oy = fmaf(-8388608.f, iy, oy); oz = fmaf(-8388608.f, iz, oz);
            // a word holds lo | hi << 16: the line enters a slab through lo when it runs in +axis direction, through hi otherwise. The
            // PRMT that builds the float picks the half, so the per-axis min / max of the float-node test disappear.
            snx = ix < 0.f ? 0x5432u : 0x5410u; sfx = snx ^ 0x0022u;
            sny = iy < 0.f ? 0x5432u : 0x5410u; sfy = sny ^ 0x0022u;
            snz = iz < 0.f ? 0x5432u : 0x5410u; sfz = snz ^ 0x0022u;
        } else {
            ox = -o.x * ix; oy = -o.y * iy; oz = -o.z * iz;
        }
        best_t = __int_as_float(0x7f800000); best_id = 0x7fffffff; best_ref = 0; have = false;
        cur = 0; leaf0 = NONE; state = ACTIVE;
    }
    __device__ __forceinline__ bool in_node() const { return state == ACTIVE && cur >= 0; }
    // keep at most one stashed leaf and, if there is more work, an inner node (or a second leaf) in `cur`
    __device__ __forceinline__ void settle() {
        if (cur == NONE) cur = ls.pop(best_t);
        if (cur < 0 && cur != NONE && leaf0 == NONE) { leaf0 = cur; cur = ls.pop(best_t); }
        if (cur == NONE && leaf0 == NONE) state = DONE;
    }
    // GLOBAL_NODES: the node array lives in global memory (not staged into shared memory): the 64-byte node is fetched with TWO
    // 256-bit loads (LDG.E.ENL2.256, sm_100) instead of three 128-bit and one 64-bit load. The lanes of a warp sit in different
    // nodes, so every load instruction costs one L1 wavefront PER LANE whatever its width - and ncu shows the L1 data pipe at
    // 94 % on the 10 000-sphere scene with four loads per visit (profiles/r2c_summary_stream_c3.txt).
    // qnodes != nullptr (warp-uniform): the 32-byte quantised node - ONE 256-bit load, twelve PRMTs - instead of the 64-byte float node.
    // QMODE: -1 the node form is decided at run time (qnodes != nullptr), 0 float nodes, 1 quantised nodes - the persistent kernels
    // are instantiated per form, which takes the pointer test (four uniform instructions) and the dead form's code out of every step.
    template <bool COUNT, bool GLOBAL_NODES = false, int QMODE = -1>
    __device__ __forceinline__ void node_step(const float4* __restrict__ nodes, TravCount& cnt, const uint4* __restrict__ qnodes = nullptr,
                                              uint32_t q2f16 = 0x4B00u) {   // requires in_node()
        if (COUNT) ++cnt.nodes;
        float4 n0, n1, n2;
        int2 ch;
        if (GLOBAL_NODES && (QMODE == 1 || (QMODE == -1 && qnodes))) {
            uint4 q0, q1;
#ifdef RTB_HOST_EMULATION
            q0 = qnodes[2 * cur]; q1 = qnodes[2 * cur + 1];
#else
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(qnodes + 2 * cur));
#endif
            // float 2^23 + q: bits 0x4B00 : q, the half chosen by the ray's near / far selector of that axis (begin())
#define RTB_QP(w, sel) __uint_as_float(__byte_perm((w), q2f16, (sel)))
            const float lo0 = fmaxf(fmaxf(fmaf(RTB_QP(q0.x, snx), ix, ox), fmaf(RTB_QP(q0.y, sny), iy, oy)), fmaf(RTB_QP(q0.z, snz), iz, oz));
            const float hi0 = fminf(fminf(fmaf(RTB_QP(q0.x, sfx), ix, ox), fmaf(RTB_QP(q0.y, sfy), iy, oy)), fmaf(RTB_QP(q0.z, sfz), iz, oz));
            const float lo1 = fmaxf(fmaxf(fmaf(RTB_QP(q0.w, snx), ix, ox), fmaf(RTB_QP(q1.x, sny), iy, oy)), fmaf(RTB_QP(q1.y, snz), iz, oz));
            const float hi1 = fminf(fminf(fmaf(RTB_QP(q0.w, sfx), ix, ox), fmaf(RTB_QP(q1.x, sfy), iy, oy)), fmaf(RTB_QP(q1.y, sfz), iz, oz));
#undef RTB_QP
            descend(lo0, hi0, lo1, hi1, (int)q1.z, (int)q1.w);
            return;
        } else
#ifndef RTB_HOST_EMULATION
        if (GLOBAL_NODES) {
            const float4* np = nodes + 4 * cur;
            float pad0, pad1;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w) : "l"(np));
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=r"(ch.x), "=r"(ch.y), "=f"(pad0), "=f"(pad1) : "l"(np + 2));
        } else
#endif
        {
            n0 = nodes[4 * cur]; n1 = nodes[4 * cur + 1]; n2 = nodes[4 * cur + 2];
            ch = *reinterpret_cast<const int2*>(nodes + 4 * cur + 3);
        }
        const float ax0 = fmaf(n0.x, ix, ox), bx0 = fmaf(n0.y, ix, ox), ay0 = fmaf(n0.z, iy, oy), by0 = fmaf(n0.w, iy, oy);
        const float az0 = fmaf(n1.x, iz, oz), bz0 = fmaf(n1.y, iz, oz);
        const float ax1 = fmaf(n1.z, ix, ox), bx1 = fmaf(n1.w, ix, ox), ay1 = fmaf(n2.x, iy, oy), by1 = fmaf(n2.y, iy, oy);
        const float az1 = fmaf(n2.z, iz, oz), bz1 = fmaf(n2.w, iz, oz);
        const float lo0 = fmaxf(fmaxf(fminf(ax0, bx0), fminf(ay0, by0)), fminf(az0, bz0));
        const float hi0 = fminf(fminf(fmaxf(ax0, bx0), fmaxf(ay0, by0)), fmaxf(az0, bz0));
        const float lo1 = fmaxf(fmaxf(fminf(ax1, bx1), fminf(ay1, by1)), fminf(az1, bz1));
        const float hi1 = fminf(fminf(fmaxf(ax1, bx1), fmaxf(ay1, by1)), fmaxf(az1, bz1));
        descend(lo0, hi0, lo1, hi1, ch.x, ch.y);
    }
    // the children's parameter intervals [lo, hi] decide where the lane goes: nearer child first, the other one waits on the stack
    __device__ __forceinline__ void descend(float lo0, float hi0, float lo1, float hi1, int c0, int c1) {
        const bool h0 = lo0 <= hi0 && hi0 >= 0.f && lo0 <= best_t;
        const bool h1 = lo1 <= hi1 && hi1 >= 0.f && lo1 <= best_t;
        if (h0 && h1) {
            const bool swap = lo1 < lo0;
            ls.push(swap ? c0 : c1, swap ? lo0 : lo1);       // the far child waits with its entry distance
            cur = swap ? c1 : c0;
        } else if (h0) cur = c0;
        else if (h1) cur = c1;
        else cur = NONE;
        settle();
    }
    // 32 bytes of a leaf slot (bvh_build.h build_leaf_slots): one 256-bit load
    __device__ __forceinline__ static void load_half_slot(const float4* p, float4& a, float4& b) {
#ifdef RTB_HOST_EMULATION
        a = p[0]; b = p[1];
#else
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#endif
    }
    // slots != nullptr: the leaf's primitives come from the leaf-ordered 64-byte slots (one 256-bit load decides a sphere or a
    // cube, two a triangle; the reference key and the object id ride along) instead of refs[] -> geometry array -> id array.
    // Same floats, same tests, same tie rule.
    template <bool COUNT>
    __device__ __forceinline__ void test_leaf(const SceneView& sc, const float4* __restrict__ sph, const float4* __restrict__ box,
                                              const int* __restrict__ refs, const float4* __restrict__ slots, int link, float3 o, float3 d,
                                              TravCount& cnt) {
        const unsigned int v = (unsigned int)(~link);
        const int first = (int)(v & 0xffffffu), n_refs = (int)(v >> 24);
        if (slots) {
            for (int i = 0; i < n_refs; ++i) {
                const float4* sp = slots + 4 * (size_t)(first + i);
                float4 a, b;
                load_half_slot(sp, a, b);
                const int r = __float_as_int(b.x), oid = __float_as_int(b.y);
                if (COUNT) { if (r >= kTriRef) ++cnt.tri; else if (r >= 0) ++cnt.sph; else ++cnt.box; }
                if (r >= kTriRef) {
                    float4 c2, d2;
                    load_half_slot(sp + 2, c2, d2);
                    float t; float3 nrm;
                    if (tri_hit(a, make_float4(b.z, b.w, c2.x, c2.y), make_float4(c2.z, c2.w, d2.x, d2.y), o, d, t, nrm)) {
                        if (t < best_t || (t == best_t && (oid < best_id || (oid == best_id && r < best_ref)))) {
                            best_t = t; best_id = oid; best_ref = r; bn = nrm; have = true;
                        }
                    }
                } else if (r >= 0) {
                    float t;
                    if (sphere_t(a, o, d, t)) {
                        if (t < best_t || (t == best_t && oid < best_id)) { best_t = t; best_id = oid; best_ref = r; have = true; }
                    }
                } else {
                    float dist; float3 nrm;
                    if (box_hit(make_float4(a.x, a.y, a.z, 0.f), make_float4(a.w, b.z, b.w, 0.f), o, d, dist, nrm)) {
                        if (dist < best_t || (dist == best_t && oid < best_id)) { best_t = dist; best_id = oid; best_ref = r; bn = nrm; have = true; }
                    }
                }
            }
            return;
        }
        for (int i = 0; i < n_refs; ++i) {
            const int r = refs[first + i];
            if (COUNT) { if (r >= kTriRef) ++cnt.tri; else if (r >= 0) ++cnt.sph; else ++cnt.box; }
            if (r >= kTriRef) {
                const int k = r - kTriRef;
                float t; float3 nrm;
                if (tri_hit(__ldg(sc.tri + 3 * k), __ldg(sc.tri + 3 * k + 1), __ldg(sc.tri + 3 * k + 2), o, d, t, nrm)) {
                    const int oid = __ldg(sc.tri_obj + k);
                    if (t < best_t || (t == best_t && (oid < best_id || (oid == best_id && r < best_ref)))) {
                        best_t = t; best_id = oid; best_ref = r; bn = nrm; have = true;
                    }
                }
            } else if (r >= 0) {
                float t;
                if (sphere_t(sph[r], o, d, t)) {
                    const int oid = sc.sph_id[r];
                    if (t < best_t || (t == best_t && oid < best_id)) { best_t = t; best_id = oid; best_ref = r; have = true; }
                }
            } else {
                const int j = ~r;
                float dist; float3 nrm;
                if (box_hit(box[2 * j], box[2 * j + 1], o, d, dist, nrm)) {
                    const int oid = sc.box_id[j];
                    if (dist < best_t || (dist == best_t && oid < best_id)) { best_t = dist; best_id = oid; best_ref = r; bn = nrm; have = true; }
                }
            }
        }
    }
    template <bool COUNT>
    __device__ __forceinline__ void leaf_step(const SceneView& sc, const float4* __restrict__ sph, const float4* __restrict__ box,
                                              const int* __restrict__ refs, const float4* __restrict__ slots, float3 o, float3 d,
                                              TravCount& cnt) {                                    // requires state == ACTIVE
        if (leaf0 != NONE) { test_leaf<COUNT>(sc, sph, box, refs, slots, leaf0, o, d, cnt); leaf0 = NONE; }
        if (cur < 0 && cur != NONE) { test_leaf<COUNT>(sc, sph, box, refs, slots, cur, o, d, cnt); cur = NONE; }
        settle();
    }
    // the closest hit as closest_hit_bvh() reports it
    __device__ __forceinline__ Hit finish(const float4* __restrict__ sph, float3 o, float3 d) const {
        Hit h;
        h.id = -1; h.t = 0.f; h.n = f3(0.f, 0.f, 0.f); h.p = f3(0.f, 0.f, 0.f);
        if (have) {
            h.id = best_id; h.t = best_t;
            h.p = f3(o.x + d.x * best_t, o.y + d.y * best_t, o.z + d.z * best_t);                          // Object.hpp:136 / :229
            if (best_ref >= 0 && best_ref < kTriRef) {
                const float4 s4 = sph[best_ref];
                h.n = normalized3(f3(h.p.x - s4.x, h.p.y - s4.y, h.p.z - s4.z));                           // Object.hpp:137
            } else h.n = bn;
        }
        return h;
    }
};


}  // namespace rtb
