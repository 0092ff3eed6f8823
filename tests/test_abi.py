"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/neuroquant_b200.h declares (no compute calls -- there is no GPU here), error strings exist,
and the product package never imports the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "neuroquant_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from neuroquant_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in the header but not exported"
    assert set(syms) == set(_lib.EXPORTS), set(syms) ^ set(_lib.EXPORTS)
    assert _lib.ABI_VERSION == raw.nq_abi_version()
    for code in (0, -1, -2, -3, -4, -5):
        assert _lib.lib.nq_status_string(code)


def test_cpu_tensors_are_rejected():
    import pytest
    import torch
    from neuroquant_b200 import _lib
    with pytest.raises(_lib.NqError):
        _lib.ptr(torch.zeros(3))
    with pytest.raises(_lib.NqError):
        _lib.uaq_init_max(torch.zeros(2, 2, 1, 1), 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "neuroquant_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def test_host_logic_geometry_and_schedule():
    import neuroquant_b200 as nq
    from neuroquant_b200.workloads import WORKLOADS, conv_flops, embed_shape
    arch, cfg = WORKLOADS["hnerv-bunny-3m"]
    geo = nq.geometry_from_cfg(cfg, arch)
    assert [(g.cout, g.cin, g.k) for g in geo] == [(92, 16, 1), (1925, 92, 1), (1024, 77, 3), (848, 64, 5), (176, 53, 5),
                                                   (148, 44, 5), (3, 37, 3)]
    assert abs(conv_flops(geo, *embed_shape(cfg, arch)[1:], n=2) / 1e9 - 202.34) < 0.01
    arch, cfg = WORKLOADS["nerv-bunny-3m"]
    geo = nq.geometry_from_cfg(cfg, arch)
    assert (geo[0].cout, geo[0].rh, geo[0].rw) == (1160, 2, 4)
    td = nq.LinearTempDecay(21000, rel_start_decay=0.2, start_b=20, end_b=2)
    assert f"{td(4500):.2f}" == "19.68" and f"{td(19500):.2f}" == "3.61"  # reference log lines


def test_tc_plans_fit_for_every_workload():
    """nq_tc_plan_conv is pure host arithmetic: every stage of every named workload (and the tiny
    golden nets) gets a tensor-core plan within the 227 KB shared-memory budget, >= 2 weight stages."""
    import ctypes as C
    import neuroquant_b200 as nq
    from neuroquant_b200 import _lib as L
    from neuroquant_b200.engine import stage_descs
    from neuroquant_b200.workloads import WORKLOADS, embed_shape
    from tests.helpers import CASES
    cases = [(a, c, embed_shape(c, a)) for a, c in WORKLOADS.values()]
    cases += [(a, c, (None, 2, 4) if a == "hnerv" else (None, 1, 1)) for a, c in CASES.values()]
    for arch, cfg, (_, h0, w0) in cases:
        geo = nq.geometry_from_cfg(cfg, arch)
        descs = stage_descs(geo, 2, h0, w0, True)
        for i, d in enumerate(descs[:-1]):
            assert d.cin_p % 16 == 0 and d.nout_p % 16 == 0
            for direction, ap, bp in ((0, 2, 1), (0, 2, 2), (1, 2, 2), (1, 1, 1)):
                if direction == 1 and i == 0:
                    continue
                pl = L.TcPlan()
                st = L.lib.nq_tc_plan_conv(C.byref(d), direction, ap, bp, C.byref(pl))
                assert st == 0, (arch, i, direction, st)
                assert pl.smem_bytes <= 227 * 1024 and pl.n_bstages >= 2
                assert pl.C % pl.SBC == 0 and pl.KC % pl.SBC == 0 and pl.SBC % 16 == 0
                assert pl.NT % 16 == 0 and pl.NT <= 256 and pl.N % 16 == 0
                assert pl.total_tiles == pl.tiles_x * pl.tiles_y * pl.tiles_n * 2
                assert (pl.CGS // 16) % 8 == 4 and pl.CGS >= pl.PW * pl.PH * 16
                # hardware shape limits of the variants the plan may pick: one MMA takes N <= 256, 512 TMEM columns
                cols = pl.NT * (2 if pl.bcat else 1)
                assert cols <= 256 and pl.mt in (1, 2) and pl.acc_stride >= cols * pl.mt and pl.n_acc * pl.acc_stride <= 512
                assert 2 <= pl.n_abuf <= 8 and 2 <= pl.n_acc <= 8 and pl.n_epi in (8, 12) and pl.n_bstages <= 16
                assert 1 <= pl.gst <= 8
                assert pl.smem_bytes >= 1024 + pl.n_abuf * pl.a_buf_bytes + pl.n_bstages * pl.gst * pl.b_stage_bytes
                if pl.gst > 1:  # several weight stages per ring slot: one bulk copy <= 32 KB, >= 3 slots in flight
                    assert pl.gst * pl.b_stage_bytes <= 32 * 1024 and pl.n_bstages >= 3 and not pl.resident
                if pl.cg2:  # CTA pairs: each CTA stages half of the columns; no side-by-side planes, no resident weights
                    assert not pl.resident and pl.NT % 16 == 0
                    # side-by-side planes in a pair plan: a full plane + half of the hi plane again per CTA
                    assert pl.b_stage_bytes == (pl.NT * pl.SBC * 3 if pl.bcat else pl.NT * pl.SBC * 2 * bp // 2)
                    assert pl.tiles_x * pl.tiles_y * 2 >= 32  # enough pixel tiles to hide the relay hop of a pair
                else:
                    assert pl.b_stage_bytes == pl.NT * pl.SBC * 2 * bp
                if pl.bcat:
                    assert ap == 2 and bp == 2
                if pl.resident:
                    assert pl.tiles_n == 1 and pl.mt == 1 and not pl.bcat and d.ksize * d.ksize <= pl.n_bstages
        assert descs[-1].cin_p % 4 == 0
        for i, d in enumerate(descs[:-1]):
            for ap, bp in ((1, 1), (2, 2)):
                wp = L.TcWgradPlan()
                st = L.lib.nq_tc_plan_wgrad(C.byref(d), ap, bp, C.byref(wp))
                if st == -3 and d.cin_p * d.ksize > 504:
                    continue  # wide 12M-class stages: more accumulator rows than one TMEM pass -> FFMA wgrad kernel
                assert st == 0, (arch, i, st)
                assert wp.nkh * wp.MB * wp.NC * (2 if wp.bcat else 1) <= 512 and wp.NC % 16 == 0
                assert wp.smem_bytes <= 227 * 1024 and wp.nbuf >= 2 and wp.NC <= 256
                if wp.bcat:  # both dZ planes as one MMA operand: N = 2 NC <= 256, single block, single column slice
                    assert wp.NC <= 128 and wp.MB == 1 and wp.nsplits == 1 and ap == 2 and bp == 2
                    assert wp.b_plane_bytes == (wp.NC // 8) * wp.CGS_B
                assert wp.khg * wp.nkh >= d.ksize and wp.AR == wp.TR + wp.nkh - 1
                assert d.ksize * wp.ncg_c + 1 <= wp.MB * 16 and wp.psplits * wp.tiles_per_split >= wp.tiles_total
                assert wp.msplit * wp.ncg_c >= wp.ncg and (wp.msplit - 1) * wp.ncg_c < wp.ncg
                assert wp.psplits * wp.nsplits * wp.khg * wp.msplit <= max(148, wp.nsplits * wp.khg * wp.msplit)


def test_packed_operand_depends_on_the_batch_size():
    """Host-only statement of what DecoderEngine._fit_packed / the geometry check of `reuse_weights` exist for: the plan of a
    stage, hence order and size of its packed weights, changes with the batch size.  Stage 3 of HNeRV-Bunny-3M (40 x 80
    pixels = 25 tiles per frame): one frame stays below the 32 tiles CTA pairs need, two frames pair up, and the pair plan
    of the data gradient holds the hi plane twice (6 instead of 4 bytes per weight).  The GPU counterpart is
    tests/test_gpu_fullsize.py::test_packed_operands_follow_the_batch_size."""
    import ctypes as C
    import neuroquant_b200 as nq
    from neuroquant_b200 import _lib as L
    from neuroquant_b200.engine import stage_descs
    from neuroquant_b200.workloads import WORKLOADS, embed_shape
    arch, cfg = WORKLOADS["hnerv-bunny-3m"]
    _, h0, w0 = embed_shape(cfg, arch)
    geo = nq.geometry_from_cfg(cfg, arch)
    plans = {}
    for n in (1, 2):
        d = stage_descs(geo, n, h0, w0, True)[3]
        for direction in (0, 1):
            pl = L.TcPlan()
            assert L.lib.nq_tc_plan_conv(C.byref(d), direction, 2, 2, C.byref(pl)) == 0
            plans[(n, direction)] = pl
    assert (d.h, d.w) == (40, 80)
    for direction in (0, 1):
        assert (plans[(1, direction)].cg2, plans[(2, direction)].cg2) == (0, 1)
    assert plans[(1, 1)].bcat == 1 and plans[(2, 1)].bcat == 1
    assert plans[(2, 1)].wpk_bytes * 2 == plans[(1, 1)].wpk_bytes * 3
    assert plans[(2, 0)].wpk_bytes == plans[(1, 0)].wpk_bytes  # same bytes, another order (each CTA of a pair streams its half)


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """Every struct the Python side mirrors has the size and trailing-field offset the C header gives it
    (compiled here with gcc: the header is plain C)."""
    import ctypes as C
    import subprocess
    from neuroquant_b200 import _lib as L
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "neuroquant_b200.h"\n'
        'int main(void) {\n'
        '  printf("%zu %zu\\n", sizeof(nq_conv_desc), offsetof(nq_conv_desc, act));\n'
        '  printf("%zu %zu\\n", sizeof(nq_tc_plan), offsetof(nq_tc_plan, wpk_bytes));\n'
        '  printf("%zu %zu\\n", sizeof(nq_tc_wgrad_plan), offsetof(nq_tc_wgrad_plan, workspace_floats));\n'
        '  printf("%zu %zu\\n", sizeof(nq_fq_task), offsetof(nq_fq_task, want_reg));\n'
        '  printf("%zu %zu\\n", sizeof(nq_ada_task), offsetof(nq_ada_task, use_reg));\n'
        '  printf("%zu %zu\\n", sizeof(nq_wgrad_finish_task), offsetof(nq_wgrad_finish_task, cin_dst));\n'
        '  printf("%zu %zu\\n", sizeof(nq_tc_pack_task), offsetof(nq_tc_pack_task, cin_src));\n'
        '  return 0;\n}\n')
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = [tuple(int(v) for v in ln.split()) for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()]
    want = [(C.sizeof(L.ConvDesc), L.ConvDesc.act.offset),
            (C.sizeof(L.TcPlan), L.TcPlan.wpk_bytes.offset),
            (C.sizeof(L.TcWgradPlan), L.TcWgradPlan.workspace_floats.offset),
            (C.sizeof(L.FqTask), L.FqTask.want_reg.offset),
            (C.sizeof(L.AdaTask), L.AdaTask.use_reg.offset),
            (C.sizeof(L.WgFinishTask), L.WgFinishTask.cin_dst.offset),
            (C.sizeof(L.TcPackTask), L.TcPackTask.cin_src.offset)]
    assert got == want


def test_tensor_core_kernels_compile_without_spills():
    """ptxas -v logs written by the build (neuroquant_b200/csrc/Makefile): no tensor-core kernel spills registers -- a
    spill in the single issuing warp or in an epilogue warp costs more than any tuning in DESIGN.md section 4 gains."""
    import os
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neuroquant_b200", "csrc")
    seen = 0
    for name in ("nq_conv_tc.o.ptxas.log", "nq_wgrad_tc.o.ptxas.log", "nq_head_tc.o.ptxas.log"):
        text = open(os.path.join(csrc, name)).read()
        for stores, loads in re.findall(r"(\d+) bytes spill stores, (\d+) bytes spill loads", text):
            assert (stores, loads) == ("0", "0"), (name, stores, loads)
            seen += 1
    assert seen >= 15
