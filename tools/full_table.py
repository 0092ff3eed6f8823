#!/usr/bin/env python
"""Table of an `ncu --set full` capture of one calibration iteration's tensor-core launches (tools/capture_round.sh):
per launch the duration, tensor-pipe activity, shared-memory wavefront shares, L2 share, DRAM bytes; also writes the DRAM
bytes per logical kernel name (the `roofline.traffic` figure of bench.py).

    python tools/full_table.py gpurun_out/<tag>_full.ncu-rep <tag> "<title>"   ->  profiles/<tag>_tc_full.md, profiles/<tag>_traffic.json
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, tag, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    g = lambda r, n: r[hdr.index(n)] if n in hdr else ""
    # logical names: launches in issue order of one eager iteration: fwd[0..L-1], head, then per stage from the last: wgrad, dgrad
    data = rows[2:]
    kern = [g(r, "Kernel Name") for r in data]
    first_head = next(i for i, k in enumerate(kern) if "head_tapexp_kernel" in k)
    start = first_head
    while start > 0 and "conv_tc_kernel" in kern[start - 1]:
        start -= 1
    nst = first_head - start
    names = {}
    for i in range(nst):
        names[start + i] = f"conv_fwd[{i}]"
    names[first_head] = "head_fwd_loss"
    j = first_head + 1
    for st in range(nst, -1, -1):
        if j < len(kern) and ("wgrad" in kern[j]):
            names[j] = f"conv_wgrad[{st}]"
            j += 1
        if st > 0 and j < len(kern) and "conv_tc_kernel" in kern[j]:
            names[j] = f"conv_dgrad[{st}]"
            j += 1
    traffic = {}
    out = [f"# {title}\n",
           "| launch | kernel | grid | us | tensor pipe active % | smem wavefronts: tensor-core operand reads % / LSU (cp.async, st.shared) % | L2 % | DRAM rd MB | DRAM wr MB | regs |",
           "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for i, r in enumerate(data):
        if i not in names:
            continue
        rd, wr = float(g(r, "dram__bytes_read.sum")) * 1e6, float(g(r, "dram__bytes_write.sum")) * 1e6
        traffic[names[i]] = rd + wr
        k = g(r, "Kernel Name").replace("void ", "").split("(")[0]
        out.append(f"| {names[i]} | `{k}` | {g(r, 'Grid Size').strip('()').split(',')[0]} | {float(g(r, 'gpu__time_duration.sum')):.1f} | "
                   f"{float(g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):.1f} | "
                   f"{float(g(r, 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed')):.1f} / "
                   f"{float(g(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed')):.1f} | "
                   f"{float(g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {g(r, 'launch__registers_per_thread')} |")
    open(os.path.join(ROOT, "profiles", f"{tag}_tc_full.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
