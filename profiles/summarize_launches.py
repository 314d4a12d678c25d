"""Per-kernel summary of an `ncu --metrics ... --csv --log-file X.csv` launch list: launches, total time, and the
instruction-weighted average of active threads per warp instruction (warp execution efficiency)."""
import csv, sys, collections
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]; iK = hdr.index("Kernel Name"); iM = hdr.index("Metric Name"); iV = hdr.index("Metric Value"); iID = hdr.index("ID")
    per = collections.defaultdict(dict)
    name = {}
    for r in rows[1:]:
        per[r[iID]][r[iM]] = float(r[iV].replace(",", "")); name[r[iID]] = r[iK].split("(")[0].split("<")[0].split("::")[-1]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    for k, m in per.items():
        a = agg[name[k]]
        inst = m.get("smsp__inst_executed.sum", 0.0)
        a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0.0) / 1e6; a[2] += inst
        a[3] += inst * m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
        a[4] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a[5] += m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) * m.get("gpu__time_duration.sum", 0.0) / 1e6
    print(path)
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-18s launches %3d  time %8.3f ms (%4.1f%%)  warp-inst %8.1f M  avg active threads/inst %5.2f  dram %8.1f MB  issue-active %4.1f%%"
              % (k, a[0], a[1], 100 * a[1] / tot, a[2] / 1e6, a[3] / max(a[2], 1), a[4] / 1e6, a[5] / max(a[1], 1e-9)))
