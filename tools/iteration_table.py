#!/usr/bin/env python
"""One graph-replayed calibration iteration out of an ncu launch list taken with
`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` (the multi-tensor launch sequence the CUDA
graph replays): per launch its duration, DRAM bytes and achieved DRAM GB/s; for the HBM-bound quantiser / optimiser
kernels also the ALGORITHMIC bytes over the duration against the measured copy peak.

    python tools/iteration_table.py gpurun_out/launches.csv [first_kernel_substring] > profiles/rNN_iteration.md
"""
import csv
import io
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# algorithmic bytes per iteration, HNeRV-Bunny-3M (2 646 219 quantised scalars): fake-quant reads x, alpha and
# writes codes, de-quantised value (16 B / element); Jacobian + Adam reads dW, x, alpha, m, v and writes alpha, m, v (32 B)
ELEMS = 2646219
ALGO = {"fakequant_fwd_multi_kernel": 16 * ELEMS, "adaround_step_multi_kernel": 32 * ELEMS}


def main():
    path = sys.argv[1]
    first = sys.argv[2] if len(sys.argv) > 2 else "fakequant_fwd_multi"
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    L = OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        d = L.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip(), "grid": r["Grid Size"]})
        v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        else:
            d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rows = list(L.values())
    starts = [k for k, d in enumerate(rows) if first in d["name"]]
    # the last complete pair of consecutive starts whose distance equals the most common distance = one replayed iteration
    gaps = [b - a for a, b in zip(starts, starts[1:])]
    # the graph replays the multi-tensor sequence (fewest launches); eager warm-up iterations are longer
    common = min(g for g in set(gaps) if g >= 20)
    a = [s for s, g in zip(starts, gaps) if g == common][-1]
    it = rows[a:a + common]
    peak = 6536.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError):
        pass
    tot = sum(d["ns"] for d in it)
    print(f"launches per iteration: {len(it)}; device time {tot / 1e3:.1f} us under ncu (serialised, cold L2: compare SHARES); "
          f"HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json)\n")
    print("| # | kernel | grid | us | share | DRAM MB (rd+wr) | DRAM GB/s | algorithmic MB | algorithmic GB/s | frac of HBM peak |")
    print("|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k, d in enumerate(it):
        by = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        short = d["name"].replace("nq::", "")
        algo = ALGO.get(short)
        extra = f"{algo / 1e6:.1f} | {algo / d['ns']:.0f} | {algo / d['ns'] / peak:.2f}" if algo else " | | "
        print(f"| {k} | `{short}` | {d['grid'].split(',')[0].strip('(')} | {d['ns'] / 1e3:.1f} | {100 * d['ns'] / tot:.1f}% | {by / 1e6:.1f} | "
              f"{by / d['ns']:.0f} | {extra} |")


if __name__ == "__main__":
    main()
