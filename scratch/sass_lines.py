#!/usr/bin/env python3
"""Static SASS statistics of one kernel: instructions per source line (nvdisasm -g line info).
usage: sass_lines.py <object.o> <kernel-name-substring> [--list] [--top N]"""
import re, subprocess, sys, tempfile, os, collections
obj, pat = sys.argv[1], sys.argv[2]
lst = "--list" in sys.argv
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
inside = False; cur = ("?", 0); per = collections.Counter(); n = 0
for l in dis:
    if l.startswith(".text."):
        inside = pat in l
        if inside: print("==", l[:160])
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        n += 1; per[cur] += 1
        if lst: print("%s:%d\t%s\t%s" % (cur[0], cur[1], m.group(1), m.group(2)))
print("instructions:", n)
if not lst:
    for (f, ln), c in per.most_common(top): print("%5d  %s:%d" % (c, f, ln))
