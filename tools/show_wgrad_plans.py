"""Print the tensor-core wgrad plan of every stage of a workload (host arithmetic only, no GPU needed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuroquant_b200 import _lib as L, workloads  # noqa: E402
from neuroquant_b200.engine import stage_descs  # noqa: E402

import ctypes as C  # noqa: E402

names = sys.argv[1:] or list(workloads.WORKLOADS)
for name in names:
    arch, cfg = workloads.WORKLOADS[name]
    geoms = workloads.geometry_from_cfg(cfg, arch)
    h0, w0 = workloads.embed_shape(cfg, arch)[-2:]
    descs = stage_descs(geoms, 2, h0, w0, True)
    print(name)
    for i, d in enumerate(descs):
        p = L.TcWgradPlan()
        st = L.lib.nq_tc_plan_wgrad(C.byref(d), 2, 2, C.byref(p))
        if st:
            print(f"  stage {i}: {d.h}x{d.w} cin_p={d.cin_p} ks={d.ksize}: status {st}")
            continue
        print(f"  stage {i}: {d.h}x{d.w} C={p.C} N={p.N} ks={d.ksize} msplit={p.msplit} ncg_c={p.ncg_c} MB={p.MB} nkh={p.nkh} "
              f"NC={p.NC} nsplits={p.nsplits} TR={p.TR} nbuf={p.nbuf} grid={p.psplits * p.msplit * p.nsplits * p.khg} "
              f"tiles/split={p.tiles_per_split}")
