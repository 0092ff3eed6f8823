#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of the kernels in an .ncu-rep (captured with --import-source on).

    python tools/ncu_source_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern, fpath, hdr, k = None, None, None, -1
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        name = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s, i_n, i_x = hdr.index("# Samples"), hdr.index("Warp Stall Sampling (Not-issued Samples)"), hdr.index("Instructions Executed")
        # a new kernel instance starts when the first file of the list repeats
        key = (name, fpath)
        if k < 0 or key in agg[k] or any(f[0] != name for f in agg[k]):
            k += 1
            agg[k] = {}
        agg[k][key] = []
        continue
    if hdr and r[0].isdigit() and r[2] == "-":  # a CUDA source line (SASS rows carry an address)
        try:
            agg[k][(name, fpath)].append((int(r[0]), r[1].strip(), int(r[i_s] or 0), int(r[i_n] or 0), int(r[i_x] or 0)))
        except ValueError:
            pass
for k, files in agg.items():
    lines = [(f[1],) + x for f, xs in files.items() for x in xs]
    tot_s = sum(x[3] for x in lines) or 1
    tot_x = sum(x[5] for x in lines) or 1
    print(f"\n## kernel instance {k}: {list(files)[0][0][:40]}  samples {tot_s}  warp instructions {tot_x}")
    print("| file:line | samples % | not-issued % | instr % | source |")
    print("|---|---:|---:|---:|---|")
    for x in sorted(lines, key=lambda x: -x[3])[:top]:
        print(f"| {x[0]}:{x[1]} | {100 * x[3] / tot_s:.1f} | {100 * x[4] / tot_s:.1f} | {100 * x[5] / tot_x:.1f} | `{x[2][:90]}` |")
