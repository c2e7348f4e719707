#!/usr/bin/env python3
"""SASS instructions of one kernel grouped by execution count (= by loop nest), with the source lines they belong to.
usage: ncu_hist.py report.ncu-rep libvmb200.so mangled_kernel [topN]"""
import csv, subprocess, re, os, tempfile, sys
rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 12
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.strip() == ".text.%s:" % kname)
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//--------------------- "):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
iI = hdr.index("Instructions Executed")
hist = {}
for k, r in enumerate(data):
    hist.setdefault(int(r[iI]), []).append(k)
tot = sum(int(r[iI]) for r in data)
print("sass %d (nvdisasm %d), warp instructions %d" % (len(data), len(lines), tot))
for t, n, ks in sorted(((n * len(v), n, v) for n, v in hist.items()), reverse=True)[:top]:
    ls = sorted(set(lines[k][1] for k in ks if k < len(lines) and lines[k][0] == "fused.cu"))
    print("exec %9d x %4d sass = %5.1f%%  fused.cu lines %s" % (n, len(ks), 100.0 * t / tot, (ls[:6] + ["..."] + ls[-6:]) if len(ls) > 12 else ls))
