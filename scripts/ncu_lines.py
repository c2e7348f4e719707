#!/usr/bin/env python3
"""Executed instructions / stall samples of one ncu capture aggregated per SOURCE line.
usage: ncu_lines.py report.ncu-rep libvmb200.so mangled_kernel_name [topN]
The SASS page of the report carries no line numbers; they come from `nvdisasm --print-line-info` on the cubin inside the .so
(same build): instruction k of the kernel in both listings is the same instruction."""
import csv, os, re, subprocess, sys, tempfile
rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.strip() == ".text.%s:" % kname)
lines = []  # (file, line) per instruction
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
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
iW, iWI = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
print("sass instructions: ncu %d, nvdisasm %d" % (len(data), len(lines)))
agg = {}
for k, r in enumerate(data):
    key = lines[k] if k < len(lines) else ("?", 0)
    a = agg.setdefault(key, [0, 0, 0, 0, 0])
    a[0] += int(r[iI]); a[1] += int(r[iS]); a[2] += 1
    a[3] += int(r[iW] or 0); a[4] += int(r[iWI] or 0)
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp instructions %d, samples %d" % (ti, ts))
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    text = ""
    for d in ("victoriametrics_b200/csrc",):
        pth = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        if os.path.exists(pth):
            if pth not in srcs:
                srcs[pth] = open(pth).read().splitlines()
            if 0 < ln <= len(srcs[pth]):
                text = srcs[pth][ln - 1].strip()[:90]
    print("%-16s %5d  inst %5.1f%%  samp %5.1f%%  sass %4d  smem wavefronts %9d (ideal %9d)  %s" % (f, ln, 100.0 * a[0] / ti, 100.0 * a[1] / max(ts, 1), a[2], a[3], a[4], text))

# ---- optional: aggregate by line ranges of one file: extra args "file:lo-hi=label" ...
regions = [a for a in sys.argv[5:] if "=" in a]
if regions:
    print("---- regions")
    rest = ti
    for spec in regions:
        rng, label = spec.split("=", 1)
        f, lr = rng.split(":")
        lo, hi = [int(x) for x in lr.split("-")]
        tot = sum(a[0] for (ff, ln), a in agg.items() if ff == f and lo <= ln <= hi)
        rest -= tot
        print("%-40s %5.1f%%  (%s)" % (label, 100.0 * tot / ti, rng))
    print("%-40s %5.1f%%" % ("everything else", 100.0 * rest / ti))
