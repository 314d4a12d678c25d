"""A/B on one box: BVH pipelines on c3 / c4 with and without the 32-byte quantised nodes (RTB200_BVH_QUANT)."""
import json
import os
import subprocess
import sys

for rep in range(2):
    for quant in (1, 0):
        for p in ("wavefront", "stream"):
            for c in ("c3", "c4"):
                env = dict(os.environ, RTB200_BVH_QUANT=str(quant))
                out = subprocess.run([sys.executable, "bench.py", "--config", c, "--no-configs", "--no-cpu", "--steps", "3", "--warmup", "2", "--pipeline", p],
                                     capture_output=True, text=True, env=env).stdout
                d = json.loads(out.strip().splitlines()[-1])
                print("quant", quant, c, p, round(d["value"]), round(d["traced_segments_per_s_M"]), round(d["ms_per_step"], 2), round(d["roofline"]["node_visits_per_query"], 2), flush=True)
