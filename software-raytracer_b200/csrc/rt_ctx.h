// rt_ctx.h - the context behind the C-ABI handle (include/rt_b200.h), shared by rt_capi.cu (single-device entry points)
// and rt_group.cu (library-owned multi-GPU). Internal: nothing here is part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_device.cuh"
#include "rt_kernels.h"
#include "bvh_build.h"
#include "bvh_wide.h"
#include "flat_build.h"
#include "mesh.h"
#include "scene_json.h"

using namespace rtb;       // internal header: only rt_capi.cu and rt_group.cu include it

// One-process-per-GPU exchange (rt_exchange_*): what a context knows about its peers.
struct ExchangeState {
    bool ready = false;
    int rank = 0, world = 1;
    const float4* accum[RT_MAX_PEERS] = {};
    ExchFlags* flags[RT_MAX_PEERS] = {};
    uint32_t* dst = nullptr;
    uint32_t epoch = 0;
};

struct rt_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;         // around the last render
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;         // around the last resolve
    cudaEvent_t ev_tune[5] = {};                      // autotuners

    HostScene scene;
    rt_camera cam;
    rt_params par;
    FrameView frame;
    SceneView view;
    bool frame_dirty = true;

    // device scene
    float4* d_sph = nullptr; int* d_sph_id = nullptr;
    float4* d_box = nullptr; int* d_box_id = nullptr;
    float4* d_mat = nullptr;
    size_t cap_sph = 0, cap_sph_id = 0, cap_box = 0, cap_box_id = 0, cap_mat = 0;
    // mesh extension: triangle records (mesh.h)
    TriRecords tris;
    float4* d_tri = nullptr; int* d_tri_obj = nullptr;
    size_t cap_tri = 0, cap_tri_obj = 0;

    // BVH (built lazily; see bvh_build.h)
    HostBvh bvh;
    BvhView bview;
    float4* d_bvh_nodes = nullptr; int* d_bvh_refs = nullptr;
    float4* d_bvh_slots = nullptr;     // leaf-ordered primitive slots (BVHs too large to stage in shared memory)
    uint4* d_bvh_qnodes = nullptr;     // 32-byte quantised nodes of the same tree (bvh_build.h HostQNodes)
    size_t cap_bvh_nodes = 0, cap_bvh_refs = 0, cap_bvh_slots = 0, cap_bvh_qnodes = 0;
    int opt_bvh_quant = 1;             // RT_OPT_BVH_QUANT
    HostWideBvh wide;                  // 8-wide quantised form for BVHs read from global memory (bvh_wide.h)
    uint4* d_wide_nodes = nullptr; int* d_wide_refs = nullptr;
    size_t cap_wide_nodes = 0, cap_wide_refs = 0;
    bool bvh_valid = false;

    // flat two-level accelerator for small scenes (built lazily; see flat_build.h)
    HostFlat flat;
    FlatView fview;
    float4* d_flat_boxes = nullptr; float4* d_flat_cull = nullptr; unsigned char* d_flat_slots = nullptr; int* d_flat_ids = nullptr;
    size_t cap_flat_boxes = 0, cap_flat_cull = 0, cap_flat_slots = 0, cap_flat_ids = 0;
    bool flat_valid = false;

    WavefrontBuffers* wf = nullptr;    // RT_PIPELINE_WAVEFRONT state (rt_wavefront.cu), allocated on first use

    // frame buffers
    float4* d_accum = nullptr;
    uint32_t* d_argb = nullptr;
    size_t cap_pixels = 0;
    unsigned long long* d_counters = nullptr;     // kCounters words, see rt_reset_accumulation
    // per-pixel primary-hit cache (RT_OPT_PRIMARY_REUSE; rt_kernels.cu k_primary_cache): valid until the camera, the scene or
    // the resolution changes
    float4* d_prim_nt = nullptr; int* d_prim_id = nullptr;
    size_t cap_prim = 0;
    bool prim_valid = false;
    ExchFlags* d_flags = nullptr;                 // this context's exchange flags (rt_exchange_*), zeroed at creation
    ExchangeState exch;
    uint32_t* d_scratch = nullptr;                // small device scratch (philox / pick)

    // accumulation state
    uint32_t samples = 0;          // samples per pixel in the buffer (global, after any external reduce)
    uint32_t next_sample = 0;      // next global sample index
    uint64_t paths = 0, total_paths = 0, total_segments_base = 0;
    int rank = 0, world = 1;
    int pixel_step = 1, strip_columns = 0;   // block-filled frames (rt_set_pixel_step)
    int opt_pipeline = RT_PIPELINE_AUTO, opt_accel = RT_ACCEL_AUTO, opt_bvh_threshold = 512;
    int opt_bvh_sched = 0, opt_bvh_wait_k = 20, opt_bvh_leaf = 4, opt_primary_reuse = 1;
    int opt_trav_stats = 0;            // RT_OPT_TRAVERSAL_STATS
    int opt_bvh_wide = 0;              // 0 (default) binary nodes, 1 wide nodes for BVHs of kWideMinPrims+ primitives, 2 always (tests)
    int opt_wf_refill = 8, opt_wf_node_min = 8, opt_wf_wave_mpaths = 0, opt_pool_tiles = 0, opt_flat_coop = 2;   // flat_coop: 0 off, 1 on, 2 measured per scene
    int tuned_flat_coop = 1;
    int tuned_accel = -1;          // RT_ACCEL_AUTO decision for the current scene/camera/params (-1: not measured yet)
    int tuned_pipeline = -1;       // RT_PIPELINE_AUTO decision for large BVH scenes (-1: not measured yet)
    float tune_pipe_ms[3] = {0.f, 0.f, 0.f};
    float4* d_tune = nullptr; size_t cap_tune = 0;
    float tune_ms[4] = {0.f, 0.f, 0.f, 0.f};
    int used_pipeline = RT_PIPELINE_REGEN, used_accel = RT_ACCEL_BRUTE;
    float last_render_ms = 0.f, last_resolve_ms = 0.f;
    bool render_timed = false, resolve_timed = false;
};
constexpr int kCounters = 32;      // the autotuners use d_counters + 4 as a scratch block of the same layout (rt_kernels.h kTileCursorSlot)


// helpers of rt_capi.cu that rt_group.cu uses
namespace rtb_capi {
int fail(rt_ctx* c, int code, const std::string& msg);
int prepare(rt_ctx* c);
int enable_peer_access_to(rt_ctx* c, const void* ptr, const char* what);   // no-op for the context's own device
int resolve_fused_unchecked(rt_ctx* c, const void* const* accum_ptrs, int world, uint32_t total_samples, int first_pixel, int n_pixels,
                            void* dst, int flip_y);
}
