"""Key metrics of EVERY kernel launch in an ncu report (one block per launch). usage: summarize_ncu_multi.py <report.ncu-rep>"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
KEYS = ["launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second"]
for vals in rows[2:]:
    m = dict(zip(hdr, vals))
    print("== %s  %s ms" % (m.get("Kernel Name", "?")[:70], m.get("gpu__time_duration.sum")))
    for k in KEYS:
        print(k, m.get(k))
    for k, v in m.items():
        if "issue_stalled" in k and "per_issue_active" in k:
            try:
                if float(v) > 0.1: print(k.replace("smsp__average_warps_issue_stalled_", ""), v)
            except ValueError:
                pass
