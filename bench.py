#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 path tracer (BASELINE.json metric:
Mpaths*bounces/s = path SEGMENTS per second, one segment = one closest-hit query + its shading).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): bundled Scene1 (67 spheres), 1920x1080, 1024 spp, depth 8,
default camera. One step = one rt_render_spp(1024) over the whole frame (2.12 G paths). At N > 1
every rank renders its own 1024 samples of the full frame (global spp = 1024*N, weak scaling) and
the float4 accumulation buffers are summed with one NCCL all-reduce inside the timed step.

`value`  : segments/s with the scene resident on the device (CUDA events around the kernels).
`e2e`    : the same metric through the C-ABI with HOST buffers each step: rt_set_scene (H2D) +
           rt_set_camera + rt_reset_accumulation + rt_render_spp + rt_resolve_rgba8 (D2H ARGB8).
`--impl reference`: the reference's own CPU code (oracle/_ref, else the validated C port) on the
           host cores, same metric, bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "software-raytracer_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "Mpaths*bounces/s (path segments per second), bundled Scene1 @1920x1080"
UNIT = "Msegments/s"
W, H, SPP, DEPTH = 1920, 1080, 1024, 8


def load_scene(name="Scene1"):
    return np.load(os.path.join(ROOT, "tests", "golden", "bundled_scenes.npz"))[name]


def flops_per_segment(objs):
    """23 per sphere test + 30 per cube test + 110 shading/scatter."""
    return 23 * int((objs["type"] == 1).sum()) + 30 * int((objs["type"] == 2).sum()) + 110


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 4 + i and r[4 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_reference_run(objs, frames, want_ref=True):
    """Time the reference's CPU implementation of the path on all host cores: `frames` 1-spp frames at
    WxH. Returns (segments/s, info). oracle/_ref (the reference's own code, 16 strip threads, per-thread
    MSVC rand) when its .so is present, else the bit-for-bit validated C port."""
    from oracle_py import Oracle, Reference, OrcCamera, REF_SO
    cores = os.cpu_count() or 1
    tmp_scene = "/tmp/_bench_scene1.json"
    if want_ref and os.path.exists(REF_SO):
        import rtb200
        rtb200.scene_file_write(tmp_scene, objs, None, "")        # host-only writer: the _ref loader needs a JSON file
        ref = Reference()
        assert ref.load_scene(tmp_scene) == len(objs)
        ref.setup(W, H, 55, DEPTH, False, None)
        sec, segs = ref.render_frames(frames, rng_mode=0, count_segments=True)
        return segs / sec, {"kind": "reference", "cores": cores, "threads": 16, "seconds": sec, "segments": int(segs),
                            "sample": "%d frames of 1 spp at %dx%d, depth %d, Scene1; reference's own renderArea loop, 16 threads, per-thread MSVC rand()" % (frames, W, H, DEPTH)}
    orc = Oracle()
    cam = OrcCamera(); cam.right[0] = 1; cam.up[1] = 1; cam.forward[2] = 1; cam.fov_deg = 55
    p = orc.default_params(width=W, height=H, max_bounces=DEPTH, mode=0)
    sec, segs = orc.time_render(objs, cam, p, frames, rng_mode=0, threads=cores)
    return segs / sec, {"kind": "port", "cores": cores, "threads": cores, "seconds": sec, "segments": int(segs),
                        "sample": "%d spp at %dx%d, depth %d, Scene1; validated C port, %d threads" % (frames, W, H, DEPTH, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    objs = load_scene()
    frames_per_step = 2
    for _ in range(args.warmup):
        cpu_reference_run(objs, 1)
    times, segs = [], 0
    info = None
    for _ in range(args.steps):
        rate, info = cpu_reference_run(objs, frames_per_step)
        times.append(info["seconds"]); segs += info["segments"]
    total = sum(times)
    value = segs / total / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic: bundled Scene1 fixture, default camera",
            "config": {"workload": "Scene1 (67 spheres) %dx%d depth %d; each step = %d frames of 1 spp on the host CPU" % (W, H, DEPTH, frames_per_step)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "paths_per_s_M": (W * H * frames_per_step * args.steps) / total / 1e6}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import rtb200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # keep stdout to the one JSON line: NCCL's version banner / debug output goes to stderr
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    global W, H, DEPTH
    objs = load_scene(args.scene)
    spp = args.spp
    cam0 = rtb200.default_camera()
    mesh = None
    workload = "Scene1 (67 spheres)" if args.scene == "Scene1" else args.scene
    scene_label = "bundled %s" % args.scene
    data_label = "synthetic: bundled %s fixture (tests/golden/bundled_scenes.npz), default camera, Philox seeds" % args.scene
    if args.config == "c3":                              # BASELINE.json configs[2]: 10k random spheres at 4K
        from rtb200.scenes import synthetic_spheres, config3_camera
        objs = synthetic_spheres(10000); cam0 = config3_camera(rtb200.default_camera); W, H = 3840, 2160
        workload = "config 3: 10 000 random spheres + ground + 8 lights"
        scene_label = "synthetic 10k spheres"
        data_label = "synthetic: rtb200.scenes.synthetic_spheres(10000) (seeded), config3_camera, Philox seeds"
    elif args.config == "c4":                            # configs[3]: ~1M-triangle mesh through the BVH
        from rtb200.scenes import heightfield_mesh, mesh_scene
        objs = mesh_scene(); mesh = heightfield_mesh(1024, 512)
        cam0.pos[1] = 1.5; cam0.pos[2] = -1.0
        workload = "config 4: 1 048 576-triangle heightfield mesh + 3 spheres"
        scene_label = "synthetic 1M-triangle mesh"
        data_label = "synthetic: rtb200.scenes.heightfield_mesh(1024, 512) (seeded) + mesh_scene(), Philox seeds"
    elif args.config == "c5":                            # configs[4]: interactive 1 spp frames at 720p
        W, H = 1280, 720
    elif args.config == "c1":                            # configs[0]: the reference's own CPU-runnable case
        W, H = 640, 480
        if spp == SPP:
            spp = 64
    stream = torch.cuda.Stream()
    tr = rtb200.PathTracer(local)
    tr.set_stream(stream.cuda_stream)
    tr.set_option(rtb200.RT_OPT_ACCEL, {"auto": rtb200.RT_ACCEL_AUTO, "brute": rtb200.RT_ACCEL_BRUTE, "bvh": rtb200.RT_ACCEL_BVH,
                                        "flat": rtb200.RT_ACCEL_FLAT}[args.accel])
    tr.set_option(rtb200.RT_OPT_PRIMARY_REUSE, 0 if args.no_primary_reuse else 1)
    tr.set_option(rtb200.RT_OPT_PIPELINE, {"auto": rtb200.RT_PIPELINE_AUTO, "regen": rtb200.RT_PIPELINE_REGEN, "wavefront": rtb200.RT_PIPELINE_WAVEFRONT}[args.pipeline])
    if args.bvh_sched >= 0:
        tr.set_option(rtb200.RT_OPT_BVH_SCHED, args.bvh_sched)
    if args.bvh_wide >= 0:
        tr.set_option(rtb200.RT_OPT_BVH_WIDE, args.bvh_wide)
    if args.wait_k >= 0:
        tr.set_option(rtb200.RT_OPT_BVH_WAIT_K, args.wait_k)
    if args.flat_coop >= 0:
        tr.set_option(rtb200.RT_OPT_FLAT_COOP, args.flat_coop)
    if args.wf_refill >= 0:
        tr.set_option(rtb200.RT_OPT_WF_REFILL, args.wf_refill)
    if args.wf_node_min >= 0:
        tr.set_option(rtb200.RT_OPT_WF_NODE_MIN, args.wf_node_min)
    tr.set_scene(objs)
    if mesh is not None:
        tr.set_mesh(0, mesh[0], mesh[1])
    tr.set_camera(cam0)
    tr.set_params(rtb200.default_params(width=W, height=H, mode=rtb200.RT_MODE_PATH, max_bounces=DEPTH,
                                        seed_lo=2026, seed_hi=rank))       # every rank: its own sample streams
    tr.reset_accumulation()
    if args.config == "c5":
        return run_interactive(args, tr, stream, torch)

    class DevBuf:                                   # wrap the library's accumulation buffer for NCCL
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    accum = torch.as_tensor(DevBuf(tr.accum_device_ptr(), W * H * 4), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: map every rank's accumulation buffer (and rank 0's surface) through CUDA IPC so the fused
    # reduce + resolve kernel can read / write them over NVLink
    peer_ptrs, dst_surface, tick = None, None, None
    fused = world > 1 and args.reduce == "fused"
    if fused:
        def gather_handles(which):
            mine = torch.tensor(list(tr.ipc_export(which)), dtype=torch.uint8, device="cuda")
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            return [bytes(h.cpu().tolist()) for h in allh]
        acc_h, srf_h = gather_handles(0), gather_handles(1)
        peer_ptrs = [tr.accum_device_ptr() if r == rank else tr.ipc_open(acc_h[r]) for r in range(world)]
        dst_surface = tr.argb_device_ptr() if rank == 0 else tr.ipc_open(srf_h[0])
        tick = torch.zeros(1, device="cuda")
    n_px = W * H
    my_first = n_px * rank // world
    my_count = n_px * (rank + 1) // world - my_first

    def exchange():
        """the path's one exchange step (N > 1), stream-ordered on every rank."""
        if fused:
            dist.all_reduce(tick)                    # device-side barrier: every rank's render has finished
            tr.resolve_fused(peer_ptrs, spp * world, my_first, my_count, dst_surface)
            dist.all_reduce(tick)                    # rank 0's surface is complete
        else:
            dist.all_reduce(accum)

    def step_resident():
        """device-resident step: samples into the accumulation buffer (+ the one exchange at N > 1)."""
        tr.render_spp(spp)
        if world > 1:
            exchange()

    # ---- device-timed: value ---------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            tr.reset_accumulation(); step_resident()
        barrier()
        seg0 = tr.stats().total_segments
        trc0 = tr.stats().total_traced_segments
        sampler = ClockSampler(local); sampler.start()
        evs = []
        tr.reset_accumulation()
        barrier()
        for _ in range(args.steps):
            flush.fill_(1)                                                  # L2 flush between timed steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); step_resident(); e1.record(stream)
            evs.append((e0, e1))
        barrier()
        clocks = sampler.summary()
        step_ms = [a.elapsed_time(b) for a, b in evs]
        segs_rank = tr.stats().total_segments - seg0
        traced_rank = tr.stats().total_traced_segments - trc0
        paths_rank = W * H * spp * args.steps

        # ---- end to end through the C-ABI with host buffers: e2e ------------------------------
        out = np.zeros((H, W), np.uint32)
        cam = cam0

        def step_e2e():
            if mesh is None:
                tr.set_scene(objs)                   # host rt_object[] -> device SoA (H2D)
            tr.set_camera(cam)
            tr.reset_accumulation()
            tr.render_spp(spp)
            if world > 1:
                exchange()
                if fused:
                    if rank == 0:
                        out[:] = tr.read_surface()   # D2H of the fused result
                    else:
                        tr.sync()
                    return
                tr.set_sample_count(spp * world)
            tr.resolve_rgba8(True, out)              # Reinhard + pack, D2H into the host surface; synchronises
        for _ in range(max(1, args.warmup - 1)):
            step_e2e()
        barrier()
        seg1 = tr.stats().total_segments
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        e1.record(stream)
        barrier()
        e2e_wall = time.perf_counter() - t0
        e2e_dev_ms = e0.elapsed_time(e1)
        e2e_segs = tr.stats().total_segments - seg1

    total_ms = sum(step_ms)
    t = torch.tensor([total_ms, e2e_wall * 1e3, float(segs_rank), float(e2e_segs)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms = tmax[0].item(), tmax[1].item()
        segs_all, e2e_segs_all = tsum[2].item(), tsum[3].item()
    else:
        e2e_ms, segs_all, e2e_segs_all = e2e_wall * 1e3, float(segs_rank), float(e2e_segs)

    if rank == 0:
        peaks, peaks_src = measured_peaks()
        st = tr.stats()
        fps = flops_per_segment(objs)
        value = segs_all / (total_ms * 1e-3) / 1e6
        # dominant kernel = k_render_regen, one launch per step; its duration = the step (N=1)
        kern_s = statistics.mean(step_ms) * 1e-3
        achieved_tf = (segs_rank / args.steps) * fps / kern_s / 1e12
        sm_mhz = peaks.get("sm_max_mhz", 1965.0)
        peak_tf = st.sm_count * 128 * 2 * sm_mhz * 1e6 / 1e12
        line = {
            "metric": METRIC.replace("bundled Scene1", scene_label).replace("1920x1080", "%dx%d" % (W, H)), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": data_label,
            "config": {"workload": "%s %dx%d, %d spp per GPU per step, depth %d, path mode" % (workload, W, H, spp, DEPTH),
                       "l2": "flushed between timed steps (256 MiB write)", "parallelism": ("spp-sharded x%d, " % world) + ("single GPU" if world == 1 else "fused reduce+resolve kernel over NVLink peer memory" if fused else "one NCCL all-reduce per step"),
                       "build": "strict IEEE, -fmad=false (bit-exact geometry vs the reference)",
                       "accel": {rtb200.RT_ACCEL_BRUTE: "brute-force object loop", rtb200.RT_ACCEL_BVH: "host-built BVH candidates + strict tests",
                                 rtb200.RT_ACCEL_FLAT: "flat two-level accelerator (conservative FMA culls, warp-uniform) + strict tests"}[st.accel],
                       "primary_reuse": ("off: every sample re-traces its primary ray" if args.no_primary_reuse else
                                         "on: the reference has no pixel jitter, so the primary closest-hit query of a pixel is identical for "
                                         "all samples; it runs once per pixel per launch and every sample shades/scatters from it "
                                         "(bit-identical radiance; `value` counts path segments delivered, traced_segments_per_s_M the queries executed)"),
                       "pipeline": {rtb200.RT_PIPELINE_REGEN: "regeneration megakernel", rtb200.RT_PIPELINE_WAVEFRONT: "wavefront (raygen / persistent intersect / shade + ballot compaction)"}[st.pipeline],
                       "scene": args.scene},
            "paths_per_s_M": paths_rank * world / (total_ms * 1e-3) / 1e6,
            "ms_per_1spp_frame": total_ms / args.steps / spp,
            "segments_per_path": segs_rank / paths_rank,
            "traced_segments_per_s_M": traced_rank * world / (total_ms * 1e-3) / 1e6,
            "traced_segments_per_path": traced_rank / paths_rank,
            "e2e": {"value": e2e_segs_all / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": int(objs.nbytes + 52), "d2h_bytes_per_step": int(W * H * 4),
                    "ms_per_step": e2e_ms / args.steps, "device_ms_per_step": e2e_dev_ms / args.steps,
                    "api": "rt_set_scene + rt_set_camera + rt_reset_accumulation + rt_render_spp + rt_resolve_rgba8(host)"},
            "gpu_launches": int(args.steps * (1 if world == 1 else 2) + args.steps * 2),
            "gpu_launches_detail": "per rank. timed value region: 1 k_render_regen per step (+ 1 k_resolve_fused at N > 1); e2e region: k_render_regen + k_resolve (N = 1) or k_resolve_fused (N > 1) per step",
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "traffic": 33.3e6, "traffic_note": "dram read+write per launch at 1080p x 1024 spp, ncu --set full (profiles/r1y_summary_final_1024spp.txt): the float4 accumulation buffer once (the write-back of the other 33 MB is still in L2 when the kernel ends); independent of spp",
                         "kernel": "k_render_regen", "kernel_ms": kern_s * 1e3,
                         "flop_per_segment": fps,
                         "peak_source": "%d SMs x 128 lanes x 2 (FMA) x %.0f MHz (%s MEASURED_PEAKS.json sm_max_mhz)" % (st.sm_count, sm_mhz, peaks_src),
                         "note": "path is FP32-CUDA-core issue bound, not HBM or tensor (SURVEY.md 8d). `achieved` is ALGORITHMIC: the reference's "
                                 "brute-force op count per segment (23/sphere + 30/cube + 110) x path segments delivered. The flat accelerator's conservative "
                                 "culls and the primary-hit reuse execute far fewer operations than that (ncu, profiles/r1y_summary_final_1024spp.txt: 25 warp "
                                 "instructions per delivered segment, issue slots 78% busy, 22.7 of 32 threads active per instruction), so frac says how much "
                                 "reference-equivalent work is delivered per peak FLOP, not how busy the FP32 pipe is",
                         "executed_warp_instr_per_segment_ncu": 25.1, "issue_active_pct_ncu": 78.4, "active_threads_per_inst_ncu": 22.65,
                         "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_algorithmic_gbs": (W * H * 32 / kern_s) / 1e9},
        }
        if st.accel == rtb200.RT_ACCEL_BVH:
            # BVH configs (3, 4): the brute-force FLOP count per segment (23 per sphere ...) is meaningless as a roofline numerator for
            # a tree traversal; these kernels are bound by instruction issue + dependent node fetches, evidenced by ncu, not by a live figure
            wf = st.pipeline == rtb200.RT_PIPELINE_WAVEFRONT
            line["roofline"] = {"bound": "issue", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                                "kernel": "k_wf_intersect_bvh" if wf else "k_render_regen<3>",
                                "note": ("wavefront pipeline: k_wf_intersect_bvh is 80 % of the step (ncu launch list profiles/r1B_launches_c3_c4.txt), issue slots "
                                         "61 % (10k spheres) / 51 % (1M triangles) busy at 15.3 of 32 threads active per instruction; k_wf_shade (11-17 %) is HBM-bound on "
                                         "the dense path state. BVH + primitives are L2-resident: neither HBM bandwidth nor FP32 peak bounds the traversal" if wf else
                                         "megakernel per-ray BVH loop: 5.4-5.5 of 32 threads active per instruction (profiles/r1l_summary_c3_bvh.txt, r1l_summary_c4_bvh.txt)")}
            line["gpu_launches_detail"] = ("per rank: the wavefront pipeline launches raygen + 2 kernels per bounce round per wave (about 20 per step) + accumulate/commit"
                                           if wf else line["gpu_launches_detail"])
            if wf:
                line["gpu_launches"] = int(args.steps * (3 + 2 * DEPTH) * 2)
        if world == 1 and not args.no_cpu and args.config in ("c1", "c2"):
            rate, info = cpu_reference_run(objs, args.cpu_frames)
            line["cpu_baseline"] = {"value": rate / 1e6, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                                    "sample": info["sample"], "threads": info["threads"]}
        print(json.dumps(line), flush=True)
    tr.close()
    if world > 1:
        dist.destroy_process_group()


def run_interactive(args, tr, stream, torch):
    """BASELINE.json configs[4]: progressive 1 spp frames at 1280x720, each frame = rt_render_spp(1) +
    rt_resolve_rgba8 into a HOST surface (D2H 3.7 MB), what a viewer's frame loop does. Frame latency p50/p99."""
    import rtb200
    out, _owner = rtb200.host_surface(W, H)          # page-locked surface (rt_host_alloc): one DMA per frame
    n = max(args.steps, 1) * 200
    with torch.cuda.stream(stream):
        for _ in range(20):
            tr.render_spp(1); tr.resolve_rgba8(True, out)
        tr.reset_accumulation(); tr.sync()
        seg0 = tr.stats().total_segments
        lat = []
        t_all = time.perf_counter()
        for _ in range(n):
            t0 = time.perf_counter()
            tr.render_spp(1)
            tr.resolve_rgba8(True, out)              # synchronises: the frame is on the host
            lat.append((time.perf_counter() - t0) * 1e3)
        total = time.perf_counter() - t_all
        segs = tr.stats().total_segments - seg0
    lat.sort()
    line = {"metric": "frame latency, progressive 1 spp/frame (BASELINE.json configs[4])", "value": lat[len(lat) // 2], "unit": "ms (p50)",
            "p99_ms": lat[int(len(lat) * 0.99) - 1], "mean_ms": 1e3 * total / n, "frames": n, "fps": n / total,
            "n_gpus": 1, "higher_is_better": False, "dtype": "f32", "data": "synthetic: bundled Scene1 fixture",
            "config": {"workload": "Scene1 %dx%d, 1 spp per frame, depth %d, render + resolve + D2H of %d bytes per frame into a page-locked host surface" % (W, H, DEPTH, W * H * 4)},
            "Msegments_per_s": segs / total / 1e6}
    print(json.dumps(line), flush=True)
    tr.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs[0..4]; the headline (and default) is c2, the others are report-only lines")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="samples per pixel per GPU per step (default: the config's 1024)")
    ap.add_argument("--cpu-frames", type=int, default=24, help="1-spp frames of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--accel", default="auto", choices=["auto", "brute", "bvh", "flat"], help="closest-hit back end (results are identical)")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "regen", "wavefront"],
                    help="regen: persistent-lane regeneration megakernel (default); wavefront: raygen/intersect/shade kernels over device queues (bit-identical)")
    ap.add_argument("--no-primary-reuse", action="store_true", help="re-trace the (identical) primary ray for every sample, like the reference")
    ap.add_argument("--reduce", default="fused", choices=["fused", "nccl"], help="N > 1 exchange: fused peer-memory reduce+resolve kernel, or NCCL all-reduce")
    ap.add_argument("--bvh-sched", type=int, default=-1, help="RT_OPT_BVH_SCHED override")
    ap.add_argument("--wait-k", type=int, default=-1, help="RT_OPT_BVH_WAIT_K override")
    ap.add_argument("--bvh-wide", type=int, default=-1, help="RT_OPT_BVH_WIDE: 0 binary BVH nodes (default), 1 8-wide quantised nodes for 1024+ primitives, 2 always")
    ap.add_argument("--flat-coop", type=int, default=-1, help="RT_OPT_FLAT_COOP: 0 per-lane levels 2/3, 1 warp-cooperative, 2 measured per scene (default)")
    ap.add_argument("--wf-refill", type=int, default=-1, help="RT_OPT_WF_REFILL override")
    ap.add_argument("--wf-node-min", type=int, default=-1, help="RT_OPT_WF_NODE_MIN override")
    ap.add_argument("--scene", default="Scene1", help="bundled scene fixture (the headline config is Scene1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
