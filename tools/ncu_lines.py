#!/usr/bin/env python
"""Per-source-line hot spots from an ncu report captured with --import-source on (kernels built with -lineinfo).
usage: tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import collections
import csv
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{regex}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
fname, func, hdr = None, None, None
seen_funcs = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        func = r[1]
        if func not in seen_funcs:
            seen_funcs.append(func)
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or func != seen_funcs[0]:
        continue
    if r[0] != "" and r[0].isdigit() and r[2] == "-":
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()])
        a[0] += int(r[hdr["# Samples"]]); a[1] += int(r[hdr["Instructions Executed"]]); a[2] += int(r[hdr["Thread Instructions Executed"]])
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
tot_t = sum(a[2] for a in agg.values())
print(f"{seen_funcs[0]}: samples {tot_s}, warp-inst {tot_i}, avg active lanes {tot_t / tot_i:.1f}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / tot_s:5.1f}% smp {100 * a[1] / tot_i:5.1f}% inst  lanes {a[2] / max(a[1], 1):4.1f}  {f}:{l}  {a[3][:110]}")
