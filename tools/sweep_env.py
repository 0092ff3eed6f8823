#!/usr/bin/env python
"""Run bench.py under several values of one environment variable and print the per-kernel times.
    python tools/sweep_env.py NQ_WG_NKH default 1 2 3 5 [-- extra bench args]"""
import json
import os
import subprocess
import sys

var, vals = sys.argv[1], sys.argv[2:]
extra = []
if "--" in vals:
    i = vals.index("--")
    vals, extra = vals[:i], vals[i + 1:]
for v in vals:
    env = dict(os.environ)
    env.pop(var, None)
    if v != "default":
        env[var] = v
    out = subprocess.run([sys.executable, "bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline"] + extra,
                         env=env, capture_output=True, text=True, timeout=300)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        k = d["roofline"]["conv_kernels_ms"]
        print(f"{var}={v}: {d['ms_per_step']:.3f} ms/step  decode {d['decode']['frames_per_s']:.0f} fps  "
              + " ".join(f"{n}={t:.3f}" for n, t in k.items() if t > 0.08), flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{var}={v}: FAILED {e} {out.stderr[-400:]}", flush=True)
