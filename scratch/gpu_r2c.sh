#!/bin/bash
# round 2: whole GPU suite (all failures listed), then ncu captures of the streaming kernel on C3 and C4
python -m pytest tests -m gpu -q 2>&1 | tail -40
for c in c3 c4; do
  CMD="python bench.py --config $c --no-configs --no-cpu --steps 1 --warmup 1 --pipeline stream"
  $CMD > gpurun_out/r2c_plain_$c.json 2> gpurun_out/r2c_plain_$c.err && \
  ncu --set full --clock-control none --import-source on -k regex:k_wf_stream -s 1 -c 1 -o gpurun_out/r2c_stream_$c -f $CMD > gpurun_out/r2c_ncu_$c.log 2>&1
  tail -3 gpurun_out/r2c_ncu_$c.log
done
ls -la gpurun_out/*.ncu-rep
