"""Hot SOURCE lines of one kernel from an ncu report captured with --import-source on (cuda,sass view):
share of the warp-level instructions executed and the average active threads per instruction, per file:line.
usage: summarize_ncu_lines.py <report.ncu-rep> [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; out = []; tot = 0
for r in csv.reader(raw.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No", "Kernel Name") or r[0] == "": continue
    try: ie, te = int(r[7]), int(r[8])
    except (ValueError, IndexError): continue
    out.append((ie, te, cur, int(r[0]), r[1].strip()[:120])); tot += ie
out.sort(reverse=True)
print("warp instructions attributed to source lines: %.3f G" % (tot / 1e9))
for ie, te, f, l, s in out[:top]:
    print("%5.2f%% thr %4.1f  %s:%d  %s" % (100.0 * ie / tot, te / max(ie, 1), f, l, s))
