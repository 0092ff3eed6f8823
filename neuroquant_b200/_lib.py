"""ctypes binding of libnq_sm100.so (C ABI declared in include/neuroquant_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every wrapper
refuses non-CUDA tensors.  PyTorch is only the owner of device memory and streams here; pointers and
sizes cross the boundary as plain integers.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NQ_LIB_PATH: load another build of the same library (A/B timing of kernel variants); default is the in-tree build
LIB_PATH = os.environ.get("NQ_LIB_PATH") or os.path.join(_HERE, "libnq_sm100.so")


class NqError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """Mirror of nq_conv_desc."""
    _fields_ = [(n, C.c_int32) for n in
                ("n", "h", "w", "cin", "cin_p", "ksize", "cout", "rh", "rw", "c_grp", "cg", "act")]

    @property
    def nout_p(self):
        return self.rh * self.rw * self.cg

    @property
    def kdim(self):
        return self.ksize * self.ksize * self.cin_p


class TcPlan(C.Structure):
    """Mirror of nq_tc_plan."""
    _fields_ = [(n, C.c_int32) for n in
                ("dir", "C", "N", "NT", "KC", "SBC", "a_planes", "b_planes", "PW", "PH", "CGS", "a_plane_bytes",
                 "a_buf_bytes", "b_stage_bytes", "n_bstages", "smem_bytes", "tiles_x", "tiles_y", "tiles_n",
                 "total_tiles", "cluster", "mt", "bcat", "n_abuf", "n_acc", "acc_stride", "n_epi", "resident", "ksplit", "cg2", "gst", "reserved")] + [("wpk_bytes", C.c_int64), ("workspace_floats", C.c_int64)]


class TcWgradPlan(C.Structure):
    """Mirror of nq_tc_wgrad_plan."""
    _fields_ = [(n, C.c_int32) for n in
                ("C", "N", "a_planes", "b_planes", "ncg", "G", "MB", "NC", "nsplits", "TR", "nkh", "khg", "AR", "msplit", "ncg_c", "bcat",
                 "CGS_A", "CGS_B",
                 "a_plane_bytes", "b_plane_bytes", "buf_bytes", "nbuf", "smem_bytes", "tiles_x", "tiles_y",
                 "tiles_total", "psplits", "tiles_per_split")] + [("workspace_floats", C.c_int64)]


class FqTask(C.Structure):
    """Mirror of nq_fq_task."""
    _fields_ = [(n, C.c_void_p) for n in ("x", "alpha", "delta", "zero_point", "codes", "deq")] + \
               [("rows", C.c_int64), ("row_len", C.c_int64)] + \
               [(n, C.c_int32) for n in ("channel_wise", "n_bits", "mode", "want_reg")]


class AdaTask(C.Structure):
    """Mirror of nq_ada_task."""
    _fields_ = [(n, C.c_void_p) for n in ("g", "x", "alpha", "delta", "zero_point", "exp_avg", "exp_avg_sq")] + \
               [("rows", C.c_int64), ("row_len", C.c_int64)] + \
               [(n, C.c_int32) for n in ("channel_wise", "n_bits", "use_reg", "reserved")]


class WgFinishTask(C.Structure):
    """Mirror of nq_wgrad_finish_task."""
    _fields_ = [("d", C.POINTER(ConvDesc)), ("workspace", C.c_void_p), ("dw_ref", C.c_void_p), ("db_ref", C.c_void_p),
                ("psplits", C.c_int32), ("n_cols", C.c_int32), ("cin_dst", C.c_int32), ("reserved", C.c_int32)]


class TcPackTask(C.Structure):
    """Mirror of nq_tc_pack_task."""
    _fields_ = [("d", C.POINTER(ConvDesc)), ("plan", C.POINTER(TcPlan)), ("w_ref", C.c_void_p), ("zero_point", C.c_void_p),
                ("wpk", C.c_void_p), ("delta", C.c_void_p), ("bias_ref", C.c_void_p), ("scale_packed", C.c_void_p),
                ("bias_packed", C.c_void_p)] + [(n, C.c_int32) for n in ("cin_src", "zp_stride", "d_stride", "reserved")]


MULTI_MAX = 16


def _load():
    if not os.path.exists(LIB_PATH):
        raise NqError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C neuroquant_b200/csrc`).  neuroquant_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    P, I, L, F, D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    DP = C.POINTER(ConvDesc)
    TP = C.POINTER(TcPlan)
    sigs = {
        "nq_status_string": (C.c_char_p, [I]),
        "nq_last_cuda_error": (I, []),
        "nq_abi_version": (I, []),
        "nq_sm_count": (I, []),
        "nq_uaq_init_max": (I, [P, L, L, I, P, P, P]),
        "nq_uaq_init_search": (I, [P, L, L, I, I, P, P, P]),
        "nq_fakequant_fwd": (I, [P, P, P, P, L, L, I, I, I, P, P, P, F, P]),
        "nq_fakequant_bwd": (I, [P, P, P, P, P, L, L, I, I, I, F, F, F, P, P, P]),
        "nq_adaround_init_alpha": (I, [P, P, L, L, I, P, P]),
        "nq_adam_step": (I, [P, P, P, P, L, D, D, D, D, I, P]),
        "nq_fakequant_bwd_soft_dev": (I, [P, P, P, P, P, L, L, I, I, F, I, P, P, P]),
        "nq_adam_step_dev": (I, [P, P, P, P, L, D, D, D, P, P]),
        "nq_fakequant_fwd_multi": (I, [C.POINTER(FqTask), I, P, F, P]),
        "nq_adaround_step_multi": (I, [C.POINTER(AdaTask), I, F, D, D, D, P, P]),
        "nq_fwht": (I, [P, P, L, I, L, L, P]),
        "nq_pack_weight": (I, [DP, P, I, P, P, P, P, P]),
        "nq_conv_fwd": (I, [DP, P, P, P, P, P, P]),
        "nq_conv_dgrad": (I, [DP, P, P, P, I, I, I, P, P]),
        "nq_conv_wgrad": (I, [DP, P, P, P, P, L, I, P]),
        "nq_unpack_wgrad": (I, [DP, P, I, P, P, P]),
        "nq_head_fwd_loss": (I, [DP, P, P, P, I, P, F, F, P, P, P, P]),
        "nq_head_wgrad_blocks": (I, [DP]),
        "nq_head_wgrad": (I, [DP, P, P, P, P, L, P]),
        "nq_tc_plan_conv": (I, [DP, I, I, I, TP]),
        "nq_tc_pack_weight": (I, [DP, TP, P, I, P, I, P, P]),
        "nq_tc_pack_epilogue": (I, [DP, P, I, P, P, P, P]),
        "nq_tc_pack_multi": (I, [C.POINTER(TcPackTask), I, P]),
        "nq_tc_conv_fwd": (I, [DP, TP, P, P, P, P, P, P, P]),
        "nq_tc_conv_dgrad": (I, [DP, TP, P, P, P, I, I, I, P, P, L, P]),
        "nq_tc_head_fwd_loss": (I, [DP, TP, P, P, P, P, I, P, F, F, P, P, P, P]),
        "nq_tc_plan_wgrad": (I, [DP, I, I, C.POINTER(TcWgradPlan)]),
        "nq_tc_conv_wgrad": (I, [DP, C.POINTER(TcWgradPlan), P, P, P, P, L, P]),
        "nq_tc_wgrad_finish_multi": (I, [C.POINTER(WgFinishTask), I, P]),
        "nq_head_wgrad_tapexp_splits": (I, [DP]),
        "nq_head_wgrad_tapexp": (I, [DP, P, P, P, L, P]),
        "nq_jet_act": (I, [P, P, P, P, P, L, I, P, P, P, I, P]),
        "nq_head_fwd_loss_split": (I, [DP, P, P, P, I, P, F, F, P, P, P, P]),
        "nq_head_fwd_loss_tapexp": (I, [DP, P, P, P, I, P, F, F, P, P, P, P]),
        "nq_head_fwd_loss_tapexp_u8": (I, [DP, P, P, P, I, P, F, F, P, P, P, P]),
        "nq_u8_to_f32": (I, [P, P, L, P]),
        "nq_jet_head": (I, [P, P, P, P, P, P, I, I, I, I, P, P]),
        "nq_nchw_to_split": (I, [P, P, I, I, I, I, I, P]),
        "nq_split_to_nchw": (I, [P, P, I, I, I, I, I, P]),
        "nq_f32_to_split": (I, [P, P, L, P]),
        "nq_split_to_f32": (I, [P, P, L, P]),
        "nq_packed_bytes": (L, [L, I]),
        "nq_pack_codes": (I, [P, L, I, P, P, P]),
        "nq_unpack_codes": (I, [P, L, I, P, P]),
        "nq_nchw_to_nhwc": (I, [P, P, I, I, I, I, I, P]),
        "nq_nhwc_to_nchw": (I, [P, P, I, I, I, I, I, P]),
        "nq_act_bwd_unshuffle": (I, [P, P, I, I, I, I, I, I, I, P, P]),
        "nq_lp_loss": (I, [P, P, L, F, F, P, P, P]),
        "nq_block_loss_bwd": (I, [P, P, P, P, I, I, I, I, I, I, F, F, P, P, P]),
        "nq_block_loss_bwd_fisher": (I, [P, P, P, P, P, I, I, I, I, I, I, I, F, P, P, P, P]),
        "nq_qdrop_gather": (I, [P, P, P, P, F, I, I, L, P, P]),
        "nq_multi_dot": (I, [P, P, P, I, I, P, P]),
        "nq_psnr": (I, [P, P, I, L, P, P]),
        "nq_omega_search_workspace": (I, [I, I, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "nq_omega_search": (I, [P, P, I, I, D, P, P, L, P, P, P]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here == header/library drift: fail loudly
        fn.restype, fn.argtypes = res, args
    return lib, tuple(sigs)


lib, EXPORTS = _load()
ABI_VERSION = lib.nq_abi_version()


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib.nq_status_string(status).decode()
        extra = f" (cudaError {lib.nq_last_cuda_error()})" if status == -5 else ""
        raise NqError(f"{what}: {msg}{extra}")


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NqError("neuroquant_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise NqError(f"expected a contiguous float32 tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# -------------------------------------------------------------------------------------------------
# thin typed wrappers (tensor in, tensor out); names follow the C ABI
# -------------------------------------------------------------------------------------------------
def rows_of(x: torch.Tensor, channel_wise: bool):
    """(rows, row_len, d_stride) of the reference's scale broadcasting: one scale per output channel
    for 4-D weights, one per tensor for 1-D biases (quantizer.py:131-152)."""
    if channel_wise and x.dim() == 4:
        return x.shape[0], x.numel() // x.shape[0], 1
    return 1, x.numel(), 0


def uaq_init_max(x, n_bits, channel_wise=True):
    rows, row_len, _ = rows_of(x, channel_wise)
    delta = torch.empty(rows, device=x.device, dtype=torch.float32)
    zp = torch.empty_like(delta)
    check(lib.nq_uaq_init_max(ptr(x), rows, row_len, n_bits, ptr(delta), ptr(zp), stream()), "nq_uaq_init_max")
    if x.dim() == 4 and channel_wise:
        return delta.view(-1, 1, 1, 1), zp.view(-1, 1, 1, 1)
    return delta.view(-1), zp.view(-1)


SCALE_METHODS = {"mse": 1, "l1": 2, "gaussian": 3}


def uaq_init_search(x, n_bits, channel_wise, method: str):
    """'mse' / 'l1' / 'gaussian' initialisers (quantizer.py:170-222), asymmetric; same shapes as uaq_init_max."""
    rows, row_len, _ = rows_of(x, channel_wise)
    delta = torch.empty(rows, device=x.device, dtype=torch.float32)
    zp = torch.empty_like(delta)
    check(lib.nq_uaq_init_search(ptr(x), rows, row_len, n_bits, SCALE_METHODS[method], ptr(delta), ptr(zp), stream()),
          "nq_uaq_init_search")
    if x.dim() == 4 and channel_wise:
        return delta.view(-1, 1, 1, 1), zp.view(-1, 1, 1, 1)
    return delta.view(-1), zp.view(-1)


def fakequant_fwd(x, alpha, delta, zp, n_bits, mode, want_codes=True, want_deq=True, reg_sum=None, reg_b=0.0):
    rows, row_len, d_stride = (delta.numel(), x.numel() // delta.numel(), 1) if delta.numel() > 1 else (1, x.numel(), 0)
    codes = torch.empty_like(x) if want_codes else None
    deq = torch.empty_like(x) if want_deq else None
    check(lib.nq_fakequant_fwd(ptr(x), ptr(alpha), ptr(delta), ptr(zp), rows, row_len, d_stride, n_bits, mode,
                               ptr(codes), ptr(deq), ptr(reg_sum), float(reg_b), stream()), "nq_fakequant_fwd")
    return codes, deq


def fakequant_bwd(g, x, alpha, delta, zp, n_bits, mode, grad_scale=1.0, reg_w=0.0, reg_b=0.0, out=None):
    rows, row_len, d_stride = (delta.numel(), x.numel() // delta.numel(), 1) if delta.numel() > 1 else (1, x.numel(), 0)
    if mode == 1:
        d_alpha = out if out is not None else torch.empty_like(x)
        d_delta = None
    else:
        d_alpha = None
        d_delta = out if out is not None else torch.empty_like(delta)
    check(lib.nq_fakequant_bwd(ptr(g), ptr(x), ptr(alpha), ptr(delta), ptr(zp), rows, row_len, d_stride, n_bits,
                               mode, float(grad_scale), float(reg_w), float(reg_b), ptr(d_alpha), ptr(d_delta),
                               stream()), "nq_fakequant_bwd")
    return d_alpha if mode == 1 else d_delta


def adaround_init_alpha(x, delta):
    rows, row_len, d_stride = (delta.numel(), x.numel() // delta.numel(), 1) if delta.numel() > 1 else (1, x.numel(), 0)
    alpha = torch.empty_like(x)
    check(lib.nq_adaround_init_alpha(ptr(x), ptr(delta), rows, row_len, d_stride, ptr(alpha), stream()),
          "nq_adaround_init_alpha")
    return alpha


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    check(lib.nq_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, beta1, beta2,
                           eps, step, stream()), "nq_adam_step")


def fwht_channel(w: torch.Tensor, out=None):
    """Orthonormal WHT along dim 1 of a contiguous (C_out, C, KH, KW) tensor."""
    co, c, kh, kw = w.shape
    out = torch.empty_like(w) if out is None else out
    inner = kh * kw
    check(lib.nq_fwht(ptr(w), ptr(out), co * inner, c, inner, c * inner, stream()), "nq_fwht")
    return out


def fwht_rows(x: torch.Tensor):
    n = x.shape[-1]
    out = torch.empty_like(x)
    check(lib.nq_fwht(ptr(x), ptr(out), x.numel() // n, n, 1, n, stream()), "nq_fwht")
    return out


def lp_loss_sum(pred, tgt, p=2.0, grad_scale=0.0, want_grad=False):
    loss = torch.zeros(1, device=pred.device, dtype=torch.float32)
    grad = torch.empty_like(pred) if want_grad else None
    check(lib.nq_lp_loss(ptr(pred), ptr(tgt), pred.numel(), float(p), float(grad_scale), ptr(loss), ptr(grad),
                         stream()), "nq_lp_loss")
    return loss, grad


def multi_dot(a_list, b_list, mode=0):
    n = len(a_list)
    a_arr = (C.c_void_p * n)(*[ptr(a) for a in a_list])
    b_arr = (C.c_void_p * n)(*[ptr(b) for b in b_list])
    s_arr = (C.c_int64 * n)(*[a.numel() for a in a_list])
    out = torch.empty(n, device=a_list[0].device, dtype=torch.float32)
    check(lib.nq_multi_dot(a_arr, b_arr, s_arr, n, mode, ptr(out), stream()), "nq_multi_dot")
    return out


def psnr(a, b):
    n = a.shape[0]
    out = torch.empty(n, device=a.device, dtype=torch.float32)
    check(lib.nq_psnr(ptr(a), ptr(b), n, a.numel() // n, ptr(out), stream()), "nq_psnr")
    return out
