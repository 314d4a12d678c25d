// rt_kernels.h - host-visible launchers of the CUDA kernels in rt_kernels.cu and rt_wavefront.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_device.cuh"

namespace rtb {

// Which closest-hit back end a launch uses, with the device views it needs.
enum { kAccelBrute = 0, kAccelBvh = 1, kAccelFlat = 2 };
struct AccelSel {
    int kind;
    BvhView bvh;
    FlatView flat;
};

// Geometry lists up to this size are staged into shared memory by every CTA
// (8192 spheres); larger scenes go through the BVH or read global memory.
constexpr size_t kMaxStagedBytes = 128 * 1024;
// A BVH (stack + geometry + nodes + refs) up to this size is staged whole; smaller keeps more CTAs per SM.
constexpr size_t kMaxBvhStagedBytes = 48 * 1024;


struct PeerPtrs { const float4* p[16]; };      // every rank's accumulation buffer, rank order (RT_MAX_PEERS)

cudaError_t launch_primary_aov(const SceneView& sc, const AccelSel& ac, const FrameView& fr, int* id, float* t,
                               float* n, float* p, cudaStream_t st);
cudaError_t launch_ray_dirs(const FrameView& fr, float* out, cudaStream_t st);
cudaError_t launch_trace_rays(const SceneView& sc, const AccelSel& ac, const float* org, const float* dir, int n,
                              int* id, float* t, float* nrm, float* pt, cudaStream_t st);
cudaError_t launch_env_color(const FrameView& fr, const float* dir, int n, float* out, cudaStream_t st);
cudaError_t launch_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t* dev_out4, cudaStream_t st);
cudaError_t launch_selftest_uniform(int* dev_failures, cudaStream_t st);
cudaError_t launch_pick(const SceneView& sc, const AccelSel& ac, const FrameView& fr, int px, int py, int* dev_id, cudaStream_t st);
cudaError_t launch_render_regen(const SceneView& sc, const AccelSel& ac, const FrameView& fr, float4* accum,
                                uint32_t s_begin, int n_samples, bool reuse_primary, unsigned long long* seg_counter, cudaStream_t st,
                                int pool_override = 0,    // 0: automatic; 1: always one pixel per lane; n >= 2: pool of n tiles per warp
                                bool flat_coop = true);   // flat accelerator: warp-cooperative levels 2/3 (rt_trace.cuh)
cudaError_t launch_render_bvh(const SceneView& sc, const AccelSel& ac, const FrameView& fr, float4* accum, uint32_t s_begin,
                              int n_samples, unsigned long long* seg_counter, int wait_k, cudaStream_t st);
cudaError_t launch_render_preview(const SceneView& sc, const AccelSel& ac, const FrameView& fr, float4* accum,
                                  unsigned long long* seg_counter, cudaStream_t st);
cudaError_t launch_render_blocks(const SceneView& sc, const AccelSel& ac, const FrameView& fr, float4* accum, uint32_t s_begin,
                                 int n_samples, int steps, int strip_w, unsigned long long* seg_counter, cudaStream_t st);
cudaError_t launch_resolve(const float4* accum, uint32_t samples, int width, int height, int first, int n, int flip_y,
                           uint32_t* out, int out_is_slice, cudaStream_t st);

cudaError_t launch_resolve_fused(const PeerPtrs& peers, int world, uint32_t samples, int width, int height, int first, int n,
                                 int flip_y, uint32_t* out, cudaStream_t st);

}  // namespace rtb

namespace rtb {
// ---- wavefront pipeline (rt_wavefront.cu) ----------------------------------------------------------
struct WavefrontBuffers;                        // device buffers of one context, owned by the C-ABI layer
WavefrontBuffers* wavefront_create();
void wavefront_destroy(WavefrontBuffers* wb);
// Adds samples [s_begin, s_begin + n_samples) of every pixel into accum, bit-identical to launch_render_regen.
cudaError_t launch_render_wavefront(WavefrontBuffers* wb, const SceneView& sc, const AccelSel& ac, const FrameView& fr,
                                    float4* accum, uint32_t s_begin, int n_samples, bool reuse_primary,
                                    unsigned long long* seg_counter, cudaStream_t st, bool bvh_refill = true,
                                    int k_refill = 8, int k_node_min = 8,
                                    int wave_mpaths = 0);   // paths per wave in units of 2^20 (0: default 128, capped by free memory)
}  // namespace rtb
