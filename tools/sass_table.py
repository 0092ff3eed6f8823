#!/usr/bin/env python
"""Mnemonic counts per tensor-core kernel of the built library (cuobjdump -sass): the table of profiles/sass_tcgen05.md.
    python tools/sass_table.py > /tmp/table.md"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "neuroquant_b200", "libnq_sm100.so")], capture_output=True, text=True).stdout
cur, cnt = None, OrderedDict()
keys = ["UTCHMMA", "UTCHMMA.2CTA", "A_KEEP", "A_REUSE", "LDTM", "UTCBAR", "UTCBAR.MULTICAST", "UTCBAR.2CTA", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS"]
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        cnt[cur] = {k: 0 for k in keys}
        continue
    if cur is None:
        continue
    c = cnt[cur]
    if "UTCHMMA" in ln:
        c["UTCHMMA.2CTA" if ".2CTA" in ln else "UTCHMMA"] += 1
        c["A_KEEP"] += "A_KEEP" in ln
        c["A_REUSE"] += "A_REUSE" in ln
    elif "UTCBAR" in ln:
        c["UTCBAR.2CTA" if ".2CTA" in ln else "UTCBAR.MULTICAST" if "MULTICAST" in ln else "UTCBAR"] += 1
    else:
        for k in ("LDTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS"):
            if re.search(r"\b" + k, ln):
                c[k] += 1
names = subprocess.run(["c++filt"] + list(cnt), capture_output=True, text=True).stdout.split("\n")
print("| kernel | UTCHMMA | UTCHMMA.2CTA | A_KEEP / A_REUSE | LDTM | UTCBAR / .MULTICAST / .2CTA | UBLKCP | UTMALDG | LDGSTS | SYNCS |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for (k, c), n in zip(cnt.items(), names):
    if c["UTCHMMA"] + c["UTCHMMA.2CTA"] == 0:
        continue
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    print(f"| `{n}` | {c['UTCHMMA']} | {c['UTCHMMA.2CTA']} | {c['A_KEEP']} / {c['A_REUSE']} | {c['LDTM']} | {c['UTCBAR']} / {c['UTCBAR.MULTICAST']} / {c['UTCBAR.2CTA']} | "
          f"{c['UBLKCP']} | {c['UTMALDG']} | {c['LDGSTS']} | {c['SYNCS']} |")
