#!/bin/bash
# r4s, eight GPUs, final build: bench.py at N=8 (weak headline + strong leg + in-bench check) and rt_headless --gpus 8
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r4s_bench_n8.json 2> gpurun_out/r4s_bench_n8.err
python scratch/show_bench.py gpurun_out/r4s_bench_n8.json | head -3 || strings gpurun_out/r4s_bench_n8.err | tail -20
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
rtb200.scene_file_write("/tmp/Scene1.json", np.load("tests/golden/bundled_scenes.npz")["Scene1"], None, "Scene1")
PY
software-raytracer_b200/bin/rt_headless --scene /tmp/Scene1.json --width 1920 --height 1080 --spp 1024 --bounces 8 --gpus 8 2>&1 | tee gpurun_out/r4s_headless_gpus8.txt
