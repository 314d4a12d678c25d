#!/usr/bin/env python
"""BASELINE.json configs[1] at ITS OWN SIZE: bundled Scene1, 1920x1080, 1024 spp, depth 8, rendered twice by the
REFERENCE'S OWN code (oracle/_ref/libref.so: renderArea + RaytraceScene + SetScreenPixel's running mean,
Raytracer.cpp:63-76,141-185,223-257) with its own rand() (per-thread MSVC LCG, 16 strip threads). Run 0 seeds every
worker with 1 (what the shipped Windows binary does), run 1 with 977: two independent estimates whose RMSE is the
reference's own Monte Carlo noise floor at this size.

    python oracle/make_goldens_c2.py         # ~2 x 6 min on 8 cores

Output: tests/golden/converged_c2_1080p.npz (float16 linear RGB means, y-up; values reach 515 = sun + sky, far inside
float16's range; its 2^-11 relative rounding is two orders of magnitude below the per-pixel noise) and the
"converged_c2" entry of tests/golden/meta.json. Committed because /root/reference does not exist on the GPU box.

TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle_py import Reference, build_ref, REFERENCE_DIR  # noqa: E402

GOLD = os.path.normpath(os.path.join(HERE, "..", "tests", "golden"))
W, H, SPP, DEPTH = 1920, 1080, 1024, 8


def main():
    build_ref()
    ref = Reference()
    assert ref.load_scene(os.path.join(REFERENCE_DIR, "Scenes", "Scene1.json")) == 67
    ref.setup(W, H, 55, DEPTH, False, None)
    imgs, info = [], {}
    for run in range(2):
        t0 = time.time()
        sec, segs = ref.render_frames(SPP, rng_mode=0, start_frame=1, count_segments=(run == 0), thread_seed=1 if run == 0 else 977)
        imgs.append(ref.color_buffer()[..., :3].astype(np.float32).copy())
        if run == 0:
            info = {"spp": SPP, "width": W, "height": H, "depth": DEPTH, "segments_per_path": segs / (W * H * SPP), "cpu_seconds": sec}
        print("run", run, "%.1fs" % (time.time() - t0), flush=True)
    a16, b16 = imgs[0].astype(np.float16), imgs[1].astype(np.float16)
    info["two_run_rmse_linear"] = float(np.sqrt(np.mean((imgs[0] - imgs[1]) ** 2)))
    info["two_run_rmse_linear_after_f16"] = float(np.sqrt(np.mean((a16.astype(np.float32) - b16.astype(np.float32)) ** 2)))
    ta, tb = imgs[0] / (1 + imgs[0]), imgs[1] / (1 + imgs[1])
    info["two_run_rmse_tonemapped"] = float(np.sqrt(np.mean((ta - tb) ** 2)))
    info["mean_rgb"] = imgs[0].mean(axis=(0, 1)).tolist()
    np.savez_compressed(os.path.join(GOLD, "converged_c2_1080p.npz"), a=a16, b=b16)
    mp = os.path.join(GOLD, "meta.json")
    with open(mp) as f:
        meta = json.load(f)
    meta["converged_c2"] = info
    with open(mp, "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(info))


if __name__ == "__main__":
    main()
