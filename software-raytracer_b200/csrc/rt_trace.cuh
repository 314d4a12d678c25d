// rt_trace.cuh - closest-hit back ends shared by the render kernels (rt_kernels.cu, rt_wavefront.cu):
// shared-memory staging of the scene per CTA, the MODE dispatch, and the host-side choice of variant.
#pragma once
#include "rt_device.cuh"
#include "rt_kernels.h"

namespace rtb {

constexpr int kTileW = 16, kTileH = 8;          // CTA tile: 4 warps, each an 8x4 pixel block
constexpr int kThreads = 128;
#ifndef RTB_BVH_POP_CULL
#define RTB_BVH_POP_CULL 1             // big (global-memory) BVHs: keep the far child's entry distance on the stack, skip stale pops
#endif
#ifndef RTB_REGEN_MIN_BLOCKS
#define RTB_REGEN_MIN_BLOCKS 6      // register budget of the render kernel: 65536 / (128 * 6) -> 80 registers
#endif

// ---- closest-hit back ends -------------------------------------------------------------------
// MODE 0: brute force, geometry staged in shared memory      MODE 1: brute force from global (L1)
// MODE 2: BVH, nodes + refs + geometry staged in shared mem  MODE 3: BVH from global (L1/L2)
// MODE 4: flat two-level accelerator (flat_build.h), everything staged in shared memory
// Shared memory layout: [BVH stack: stack_entries x blockDim ints][spheres][cubes][nodes][refs]
//                       MODE 4: [candidate queues: kFlatQueue x blockDim bytes][spheres][cubes][level-1 boxes][cull records][prim ids][cull slots]
struct TraceCtx {
    const float4* sph; const float4* box; const float4* nodes; const int* refs;
    int* stack; float* stack_t; int stride;
    FlatView fl;
    unsigned char* q;          // MODE 4: this thread's candidate queue, entries `stride` bytes apart
};

template <int MODE>
__device__ __forceinline__ TraceCtx setup_trace(const SceneView& sc, const BvhView& bv, const FlatView& fl, float4* smem) {
    TraceCtx t;
    t.sph = sc.sph; t.box = sc.box; t.nodes = bv.nodes; t.refs = bv.refs; t.stack = nullptr; t.stack_t = nullptr; t.stride = blockDim.x;
    t.fl = fl; t.q = nullptr;
    float4* p = smem;
    if (MODE == 4) {
        t.q = reinterpret_cast<unsigned char*>(p) + threadIdx.x;
        p += (kFlatQueue * blockDim.x + 15) / 16;
        const int ng = sc.n_sph + 2 * sc.n_box, nb = 2 * (fl.n_clusters + fl.n_cubes), ncull = 8 * fl.n_clusters + fl.n_singles;
        const int np = sc.n_sph + sc.n_box;
        for (int i = threadIdx.x; i < ng; i += blockDim.x) p[i] = i < sc.n_sph ? __ldg(sc.sph + i) : __ldg(sc.box + (i - sc.n_sph));
        t.sph = p; t.box = p + sc.n_sph; p += ng;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) p[i] = __ldg(fl.boxes + i);
        t.fl.boxes = p; p += nb;
        for (int i = threadIdx.x; i < ncull; i += blockDim.x) p[i] = __ldg(fl.cull + i);
        t.fl.cull = p; p += ncull;
        int* ids = reinterpret_cast<int*>(p);
        for (int i = threadIdx.x; i < np; i += blockDim.x) ids[i] = __ldg(fl.prim_id + i);
        t.fl.prim_id = ids;
        unsigned char* slots = reinterpret_cast<unsigned char*>(ids + np);
        for (int i = threadIdx.x; i < ncull; i += blockDim.x) slots[i] = __ldg(fl.cull_slot + i);
        t.fl.cull_slot = slots;
        __syncthreads();
        return t;
    }
    if (MODE >= 2) {
        t.stack = reinterpret_cast<int*>(p) + threadIdx.x;
        t.stack_t = reinterpret_cast<float*>(t.stack + bv.stack_entries * blockDim.x);
        p += (2 * bv.stack_entries * blockDim.x + 3) / 4;      // links + entry distances
    }
    if (MODE == 0 || MODE == 2) {
        const int ng = sc.n_sph + 2 * sc.n_box;
        for (int i = threadIdx.x; i < ng; i += blockDim.x)
            p[i] = i < sc.n_sph ? __ldg(sc.sph + i) : __ldg(sc.box + (i - sc.n_sph));
        t.sph = p; t.box = p + sc.n_sph;
        p += ng;
        if (MODE == 2) {
            for (int i = threadIdx.x; i < 4 * bv.n_nodes; i += blockDim.x) p[i] = __ldg(bv.nodes + i);
            t.nodes = p;
            p += 4 * bv.n_nodes;
            int* r = reinterpret_cast<int*>(p);
            for (int i = threadIdx.x; i < bv.n_refs; i += blockDim.x) r[i] = __ldg(bv.refs + i);
            t.refs = r;
        }
        __syncthreads();
    }
    return t;
}

template <int MODE>
__device__ __forceinline__ Hit trace(const SceneView& sc, const TraceCtx& t, float3 o, float3 d) {
    if (MODE == 4) return closest_hit_flat(sc, t.fl, t.sph, t.box, t.q, t.stride, o, d);
    if (MODE >= 2) return closest_hit_bvh(sc, t.sph, t.box, t.nodes, t.refs, t.stack, t.stride, o, d, (RTB_BVH_POP_CULL && MODE == 3) ? t.stack_t : nullptr);
    return closest_hit(sc, t.sph, t.box, o, d);
}


// ---- host: variant + dynamic shared memory size for a back end -----------------------------------
inline size_t staged_bytes(const SceneView& sc) { return (size_t)(sc.n_sph + 2 * sc.n_box) * sizeof(float4); }
inline size_t flat_staged_bytes(const SceneView& sc, const FlatView& fl, int threads) {
    const size_t ncull = (size_t)8 * fl.n_clusters + fl.n_singles;
    return (size_t)kFlatQueue * threads + staged_bytes(sc) + (size_t)2 * (fl.n_clusters + fl.n_cubes) * 16 + ncull * 16 + (size_t)(sc.n_sph + sc.n_box) * 4 + ncull + 16;
}
inline int pick_mode(const SceneView& sc, const AccelSel& ac, size_t& smem, int threads = kThreads) {
    const size_t geo = staged_bytes(sc);
    if (ac.kind == kAccelFlat) { smem = flat_staged_bytes(sc, ac.flat, threads); return 4; }
    if (ac.kind == kAccelBrute) {
        if (geo <= kMaxStagedBytes) { smem = geo; return 0; }
        smem = 0; return 1;
    }
    const BvhView& bv = ac.bvh;
    const size_t stack = ((size_t)2 * bv.stack_entries * threads * sizeof(int) + 15) / 16 * 16;
    const size_t all = stack + geo + (size_t)bv.n_nodes * 64 + (size_t)bv.n_refs * 4 + 16;
    if (all <= kMaxBvhStagedBytes) { smem = all; return 2; }
    smem = stack; return 3;
}

}  // namespace rtb
