#!/usr/bin/env python
"""Warp instructions and average active lanes per source function, from an ncu report (--page source --csv) joined with
the line table of the same cubin (nvdisasm -g -c).  Shows where a kernel spends its instructions at low lane counts.
usage: ncu_lanes.py report.ncu-rep kernel.sass path/to/render.cuh [topN]"""
import bisect
import collections
import csv
import io
import re
import subprocess
import sys



def kernel_section(path, pattern=r"render_kernel.*Lb0"):
    """The lines of the one function of an `nvdisasm -g -c` listing whose .text section name matches `pattern` (a cubin holds
    several kernels, each with addresses from 0)."""
    out, on = [], False
    for ln in open(path):
        if ln.startswith("//---------------------"):
            on = re.search(pattern, ln) is not None
            continue
        if on:
            out.append(ln)
    return out or list(open(path))

rep, sass, srcpath = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 16
starts = []
for i, l in enumerate(open(srcpath).read().split("\n"), 1):
    if l.startswith(("FTB_DEV", "__global__", "__device__")):
        m = re.search(r"(\w+)\(", l)
        if m:
            starts.append((i, m.group(1) if m.group(1) != "__launch_bounds__" else "render_kernel"))
keys = [s for s, _ in starts]
addr_line, cur = {}, None
for ln in kernel_section(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        addr_line.setdefault(int(m.group(1), 16), cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, ii, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg, base = collections.defaultdict(lambda: [0.0, 0.0]), None
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    base = a if base is None else base
    f, l = addr_line.get(a - base, ("?", 0))
    fn = f
    if f.endswith("render.cuh"):
        k = bisect.bisect_right(keys, l) - 1
        fn = starts[k][1] if k >= 0 else "?"
    agg[fn][0] += float(r[ii] or 0)
    agg[fn][1] += float(r[it] or 0)
tot = sum(v[0] for v in agg.values())
print("%-28s %8s %8s" % ("function", "instr %", "lanes"))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%-28s %7.1f%% %8.1f" % (k, 100 * v[0] / tot, v[1] / max(v[0], 1)))
print("%-28s %7.1f%% %8.1f" % ("all", 100.0, sum(v[1] for v in agg.values()) / tot))
