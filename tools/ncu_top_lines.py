#!/usr/bin/env python
"""Top stall source lines of an .ncu-rep (needs -lineinfo + --import-source on):
   python tools/ncu_top_lines.py gpurun_out/x.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, ix, agg, kname, first_kernel = None, {}, [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kname = r[1]
        if first_kernel is None:
            first_kernel = kname
        elif kname != first_kernel or agg and kname == first_kernel and cur_file is None:
            pass
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        ix = {}
        for i, h in enumerate(r):
            ix.setdefault(h, i)
        continue
    if ix and r[0].isdigit():
        try:
            s = int(r[ix["# Samples"]])
        except (ValueError, KeyError):
            s = 0
        agg.append((s, cur_file, int(r[0]), r[1].strip()[:88], r))
total = sum(a[0] for a in agg) or 1
print(f"{rep}: {total} samples (all captured launches)")
for s, f, ln, src, r in sorted(agg, key=lambda a: -a[0])[:topn]:
    g = lambda k: r[ix[k]] if k in ix else "?"
    print(f"{s:6d} {100 * s / total:5.1f}%  {f}:{ln:<4d} long_sb={g('stall_long_sb'):>5s} barrier={g('stall_barrier'):>5s} "
          f"lg={g('stall_lg'):>4s} membar={g('stall_membar'):>4s} | {src}")
