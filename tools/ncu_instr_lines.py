#!/usr/bin/env python
"""Top source lines by executed warp instructions for every kernel instance of an .ncu-rep (--import-source on).

    python tools/ncu_instr_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
inst, fp, name, hdr, first_file = -1, None, None, None, None
res = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fp = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if r[1] != name:
            first_file = None
        name = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_x, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
        if first_file is None or fp == first_file:
            first_file = fp
            inst += 1
        continue
    if hdr and r[0].isdigit() and r[2] == "-":
        v = res.setdefault(inst, {}).setdefault((fp, int(r[0])), [r[1].strip()[:95], 0, 0, name])
        v[1] += int(r[i_x] or 0)
        v[2] += int(r[i_s] or 0)
for k, lines in res.items():
    tot = sum(v[1] for v in lines.values()) or 1
    smp = sum(v[2] for v in lines.values()) or 1
    nm = next(iter(lines.values()))[3]
    print(f"\n## instance {k}: {nm[:60]}  warp instructions {tot}  samples {smp}")
    for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {f}:{l:4d} instr {100 * v[1] / tot:5.1f}%  samples {100 * v[2] / smp:5.1f}%  {v[0]}")
