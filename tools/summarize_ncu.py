#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table
(count, total, share), and optionally pull the roofline-relevant raw metrics out of an .ncu-rep.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
    python tools/summarize_ncu.py raw gpurun_out/prof.ncu-rep        > profiles/rNN_kernel.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        rows.append((name, ns))
    agg = OrderedDict()
    for n, ns in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"launches: {len(rows)}   total device time: {tot / 1e6:.3f} ms (cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | total ms | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {ns / 1e6:.3f} | {ns / c / 1e3:.1f} | {100 * ns / tot:.1f}% |")


KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active", "smsp__inst_executed.sum", "l1tex__t_bytes", "sm__inst_executed_pipe_uniform")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for row in rd[2:]:
        rec = dict(zip(hdr, row))
        print(f"### {rec.get('Kernel Name', '?')}  (id {rec.get('ID')})\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for h, u in zip(hdr, units):
            if any(h.startswith(k) for k in KEYS):
                print(f"| {h} | {rec[h]} | {u} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
