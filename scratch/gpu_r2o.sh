#!/bin/bash
# round 2 checkpoint: smoke(), suite, both bench arms, launch list of the default command
python -c "import __graft_entry__ as g; g.smoke()"
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2o_bench_reference_arm.json 2> gpurun_out/r2o_bench_reference_arm.err
python -c "
import json; d=json.loads(open('gpurun_out/r2o_bench_reference_arm.json').read().strip().splitlines()[-1]); print('reference arm', d['value'], d['unit'], d['cpu_baseline']['cores'], d['config'])"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err
tail -c 600 gpurun_out/r2o_bench_n1.err
python scratch/show_bench.py gpurun_out/r2o_bench_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --accel flat"
$CMD > gpurun_out/r2o_plain.json 2> gpurun_out/r2o_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2o_launches.csv $CMD > gpurun_out/r2o_ncu_launches.log 2>&1
tail -1 gpurun_out/r2o_ncu_launches.log | cut -c1-200
grep -c "librt_b200" /proc/self/maps || true
