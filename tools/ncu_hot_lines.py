#!/usr/bin/env python
"""Joins the per-SASS-instruction samples of an ncu report (--page source --csv) with the line table of the
same kernel (nvdisasm -g of a cubin built from the same source) and prints the hottest source lines.
usage: ncu_hot_lines.py report.ncu-rep kernel.sass(from nvdisasm -g -c) [topN]"""
import csv
import io
import os
import re
import subprocess
import sys
import collections



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

rep, sass = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
addr_line = {}
cur = None
for ln in kernel_section(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        addr_line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isamp, iinst, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0])
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
stall = collections.defaultdict(lambda: collections.Counter())
tot_s = tot_i = 0.0
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    key = addr_line.get(a - base, ("?", 0))
    s, i = float(r[isamp] or 0), float(r[iinst] or 0)
    agg[key][0] += s; agg[key][1] += i
    tot_s += s; tot_i += i
    for c in stall_cols:
        v = float(r[c] or 0)
        if v:
            stall[key][hdr[c]] += v
src = {}
for f in set(k[0] for k in agg):
    try:
        # FTB_SRC_DIR: a directory holding the sources as they were when the profiled library was built
        path = subprocess.run(["find", os.environ.get("FTB_SRC_DIR", "."), "-name", f, "-not", "-path", "./.git/*"], stdout=subprocess.PIPE, text=True).stdout.split()[0]
        src[f] = open(path).read().splitlines()
    except Exception:
        src[f] = []
print("total samples %.0f, warp instructions %.0f" % (tot_s, tot_i))
for key, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src.get(key[0], [])
    line = text[key[1] - 1].strip()[:95] if 0 < key[1] <= len(text) else ""
    st = ",".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(s, 1)) for k, v in stall[key].most_common(2))
    print("%-16s L%4d  samples %5.1f%%  inst %5.1f%%  [%s]  %s" % (key[0], key[1], 100 * s / tot_s, 100 * i / tot_i, st, line))
