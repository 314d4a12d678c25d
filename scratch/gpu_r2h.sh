#!/bin/bash
# round 2: 32-byte quantised nodes: suite, pipelines on c3/c4 with and without them, ncu of the stream kernel
python -m pytest tests -m gpu -q 2>&1 | tail -8
cat > /tmp/ab.py <<'PY'
import sys, json, subprocess
for quant in (1, 0):
    for p in ("wavefront", "stream"):
        for c in ("c3", "c4"):
            import os
            env = dict(os.environ, RTB200_BVH_QUANT=str(quant))
            out = subprocess.run([sys.executable, "bench.py", "--config", c, "--no-configs", "--no-cpu", "--steps", "3", "--warmup", "2", "--pipeline", p],
                                 capture_output=True, text=True, env=env).stdout
            d = json.loads(out.strip().splitlines()[-1])
            print("quant", quant, c, p, round(d["value"]), round(d["traced_segments_per_s_M"]), round(d["ms_per_step"], 2), round(d["roofline"]["node_visits_per_query"], 2), flush=True)
PY
python /tmp/ab.py
for c in c3 c4; do
  CMD="python bench.py --config $c --no-configs --no-cpu --steps 1 --warmup 1 --pipeline stream"
  $CMD > gpurun_out/r2h_plain_$c.json 2> gpurun_out/r2h_plain_$c.err && \
  ncu --set full --clock-control none --import-source on -k regex:k_wf_stream -s 1 -c 1 -o gpurun_out/r2h_stream_$c -f $CMD > gpurun_out/r2h_ncu_$c.log 2>&1
  tail -2 gpurun_out/r2h_ncu_$c.log
done
