#!/bin/bash
# r3a: end-of-round checkpoint on one GPU: smoke, suite, both bench arms as the driver calls them, launch list, full captures of the four kernels
python -c "import __graft_entry__ as g; g.smoke()"
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r3a_bench_reference_arm.json 2> gpurun_out/r3a_bench_reference_arm.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r3a_bench_n1.json 2> gpurun_out/r3a_bench_n1.err
tail -c 300 gpurun_out/r3a_bench_n1.err
python scratch/show_bench.py gpurun_out/r3a_bench_n1.json
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, "software-raytracer_b200/python"); import rtb200
rtb200.scene_file_write("/tmp/Scene1.json", np.load("tests/golden/bundled_scenes.npz")["Scene1"], None, "Scene1")
PY
software-raytracer_b200/bin/rt_headless --scene /tmp/Scene1.json --width 1280 --height 720 --interactive 300 --scale 1 --bounces 8 2>&1 | tail -1 | tee gpurun_out/r3a_headless_interactive.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --accel flat"
$CMD > gpurun_out/r3a_plain.json 2> gpurun_out/r3a_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r3a_launches.csv $CMD > gpurun_out/r3a_ncu_launches.log 2>&1
tail -1 gpurun_out/r3a_ncu_launches.log | cut -c1-200
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --accel flat"
ncu --set full --clock-control none --import-source on -k regex:k_render_regen -s 4 -c 1 -o gpurun_out/r3a_regen_c2 -f $CMD2 > gpurun_out/r3a_ncu_regen.log 2>&1; tail -1 gpurun_out/r3a_ncu_regen.log | cut -c1-200
CMD3="python bench.py --config c5 --steps 1 --warmup 3 --no-cpu --no-configs"
ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 60 -c 1 -o gpurun_out/r3a_pool_c5 -f $CMD3 > gpurun_out/r3a_ncu_pool.log 2>&1; tail -1 gpurun_out/r3a_ncu_pool.log | cut -c1-200
for c in c3 c4; do
  CMD4="python bench.py --config $c --no-configs --no-cpu --steps 1 --warmup 1 --pipeline wavefront"
  ncu --set full --clock-control none --import-source on -k regex:"k_wf_intersect_bvh|k_wf_shade" -s 16 -c 2 -o gpurun_out/r3a_wf_$c -f $CMD4 > gpurun_out/r3a_ncu_$c.log 2>&1; tail -1 gpurun_out/r3a_ncu_$c.log | cut -c1-200
done
