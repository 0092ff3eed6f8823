"""Print the tensor-core forward / dgrad plans of every stage of a workload (host arithmetic only, no GPU needed)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuroquant_b200 import _lib as L, workloads  # noqa: E402
from neuroquant_b200.engine import stage_descs  # noqa: E402

names = sys.argv[1:] or list(workloads.WORKLOADS)
for name in names:
    arch, cfg = workloads.WORKLOADS[name]
    geoms = workloads.geometry_from_cfg(cfg, arch)
    h0, w0 = workloads.embed_shape(cfg, arch)[-2:]
    descs = stage_descs(geoms, 2, h0, w0, True)
    print(name)
    for i, d in enumerate(descs):
        for dr, tag in ((0, "fwd"), (1, "dgrad")):
            p = L.TcPlan()
            st = L.lib.nq_tc_plan_conv(C.byref(d), dr, 2, 2, C.byref(p))
            if st:
                print(f"  {tag}[{i}]: status {st}")
                continue
            print(f"  {tag}[{i}]: {d.h}x{d.w} ks={d.ksize} C={p.C} N={p.N} NT={p.NT} mt={p.mt} bcat={p.bcat} cg2={p.cg2} ks={p.ksplit} res={p.resident} epi={p.n_epi} abuf={p.n_abuf} acc={p.n_acc}x{p.acc_stride} KC={p.KC} SBC={p.SBC} PW={p.PW} "
                  f"a_buf={p.a_buf_bytes} stage={p.b_stage_bytes} x{p.gst} x{p.n_bstages} smem={p.smem_bytes} tiles={p.total_tiles}")
