"""Prints the figures of a bench.py JSON line that matter when iterating."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "traced_segments_per_s_M", "n_gpus", "scaling", "multi_gpu_check", "exchange_ms")})
print("e2e", d["e2e"]["value"], "roofline frac", d["roofline"]["frac"], d["roofline"].get("frac_delivered"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
if d.get("strong"):
    print("strong", d["strong"])
for l in d.get("configs", []):
    print({k: l.get(k) for k in ("config", "value", "traced_segments_per_s_M", "ms_per_step", "pipeline", "p99_ms", "render_kernel_ms_p50", "error") if l.get(k) is not None},
          "frac", (l.get("roofline") or {}).get("frac"), (l.get("roofline") or {}).get("basis", "")[:160])
