#!/bin/bash
# round 2, eight GPUs: bench.py at N=8 (weak headline + strong leg + in-bench check), N=4, rt_headless --gpus 8, C3 strong
nvidia-smi -L | wc -l
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2g_bench_n$n.json 2> gpurun_out/r2g_bench_n$n.err
python scratch/show_bench.py gpurun_out/r2g_bench_n$n.json || strings gpurun_out/r2g_bench_n$n.err | tail -20
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 2 --config c3 --scaling strong --spp 128 > gpurun_out/r2g_bench_c3_strong_n8.json 2> gpurun_out/r2g_bench_c3_strong_n8.err
python scratch/show_bench.py gpurun_out/r2g_bench_c3_strong_n8.json || strings gpurun_out/r2g_bench_c3_strong_n8.err | tail -20
python - <<'PY'
import numpy as np, sys, json
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
objs = np.load("tests/golden/bundled_scenes.npz")["Scene1"]
rtb200.scene_file_write("/tmp/Scene1.json", objs, None, "Scene1")
PY
software-raytracer_b200/bin/rt_headless --scene /tmp/Scene1.json --width 1920 --height 1080 --spp 1024 --bounces 8 --gpus 8
software-raytracer_b200/bin/rt_headless --scene /tmp/Scene1.json --width 1920 --height 1080 --spp 1024 --bounces 8 --gpus 1 2>/dev/null || true
