"""Decoder engine: runs a quantised NeRV/HNeRV decoder (forward, loss, backward, Adam) as a fixed
sequence of libnq_sm100 kernel launches on one CUDA stream.

This is the execution layer under the reference-shaped API in `neuroquant_b200.quantization` /
`neuroquant_b200.models`.  It replaces, for the decoder only (the encoder is never quantised,
quant_model.py:28-29):

  * QuantModule.forward                       (quant_layer.py:67-81)
  * QuantNeRVBlock.forward                    (quant_block.py:31-35)
  * HNeRV.decode / NeRV.decode + OutImg       (HNeRV.py:49-71, NeRV.py:44-65, _layers.py:10-16)
  * lp_loss + LossFunction.collect_round_loss (quantizer.py:66-73, calib_model.py:39-47)
  * autograd of all of the above + torch.optim.Adam.step (calib_model.py:145-165, :206-226)

Data layout in HBM:
  activations  tensor-core engine (default): "split-bf16" NHWC -- two bf16 planes hi = bf16(v), lo = bf16(v - hi), channels
               padded to 16 (8 for the head's input), pads kept at zero; exact-fp32 FFMA engine (NQ_CONV=simt): fp32 NHWC,
               channels padded to 4
  targets      (n, 3, H, W) fp32 in [0, 1], or uint8 as the data set stores them (value / 255 evaluated on the device)
  weights      reference layout (C_out, C_in[pow2 if rotated], k, k) for everything the quantiser
               touches (weight, alpha, delta, zero_point, codes);  "packed" GEMM layouts
               (include/neuroquant_b200.h) for what the convolutions read
  gradients    one flat buffer holding every dW (reference layout) and db, so that data-parallel
               ranks exchange it with a single NCCL all-reduce

There is no PyTorch fallback: every step is a kernel of libnq_sm100.so.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

ROUND_NEAREST, ROUND_SOFT, ROUND_HARD = 0, 1, 2
_ACT = {"none": 0, "gelu": 1}
_ACT_SAVED_GRAD = 2  # nq_conv_desc.act: GELU whose z buffer holds the derivative
_HEAD = {"tanh": 0, "sigmoid": 1}


def _pad(c: int, m: int) -> int:
    return (c + m - 1) // m * m


def _pad4(c: int) -> int:
    return _pad(c, 4)


def next_pow2(n: int) -> int:
    """quant_layer.py:13-14."""
    return 1 if n == 0 else 1 << (n - 1).bit_length()


@dataclass
class StageGeom:
    """One decoder stage: conv(k, stride 1, same) -> up-shuffle(rh, rw) -> act."""
    cin: int
    cout: int
    k: int
    rh: int = 1
    rw: int = 1
    act: str = "none"  # 'none' | 'gelu' for hidden stages; 'tanh' | 'sigmoid' for the head

    @property
    def c_grp(self) -> int:
        return self.cout // (self.rh * self.rw)


def geometry_from_cfg(cfg: dict, arch: str) -> List[StageGeom]:
    """Stage list of a reference YAML config (HNeRV.py:12-43, NeRV.py:12-38)."""
    import numpy as np

    arch = arch.lower()
    strides = list(cfg["dec_strides"])
    geo: List[StageGeom] = []
    c = cfg["dec_in_channel"]
    if arch == "hnerv":
        fc = int(np.prod(cfg["enc_strides"]) // np.prod(strides))
        geo.append(StageGeom(cfg["enc_channel"][-1], c, 1, 1, 1, "none"))  # HNeRV.py:32 (fc folds below)
        if fc != 1:
            # HNeRV.py:57 folds (fc_h, fc_w) out of the channel axis of the stem output
            if c % (fc * fc):
                raise ValueError("dec_in_channel must be divisible by fc^2")
            geo[0] = StageGeom(cfg["enc_channel"][-1], c, 1, fc, fc, "none")
            c = c // (fc * fc)
    elif arch == "nerv":
        fch = cfg["crop_h"] // int(np.prod(strides))
        fcw = cfg["crop_w"] // int(np.prod(strides))
        geo.append(StageGeom(int(cfg["level"] * 2), c * fch * fcw, 1, fch, fcw, "none"))  # NeRV.py:26,51
    else:
        raise ValueError(f"unsupported arch {arch!r} (quant_model.py:13 supports NeRV and HNeRV)")
    if cfg.get("dec_norm", "none") != "none":
        raise NotImplementedError("QuantNeRVBlock drops the norm layer (quant_block.py:27-29); only dec_norm: none")
    if cfg["dec_acts"] not in _ACT:
        raise NotImplementedError(f"dec_acts={cfg['dec_acts']!r}: only 'gelu'/'none' have fused epilogues")
    for ks, s in zip(cfg["dec_kernels"], strides):
        co = int(max(round(c / cfg["channel_reduce"]), cfg["channel_lbound"]))
        geo.append(StageGeom(c, co * s * s, ks, s, s, cfg["dec_acts"]))
        c = co
    if cfg["out_bias"] not in _HEAD:
        raise NotImplementedError(f"out_bias={cfg['out_bias']!r}: only tanh / sigmoid heads")
    geo.append(StageGeom(c, 3, 3, 1, 1, cfg["out_bias"]))
    return geo


class QuantStage:
    """Device state of one QuantModule: weights, quantiser parameters, rounding variables.

    Tensors are shared by reference with the API objects (QuantModule / quantizers) that expose
    them under the reference's attribute names; the engine never copies them.
    """

    def __init__(self, geom: StageGeom, weight: torch.Tensor, bias: torch.Tensor, n_bits: int = 8,
                 hadamard: bool = False):
        assert tuple(weight.shape) == (geom.cout, geom.cin, geom.k, geom.k), (tuple(weight.shape), geom)
        self.geom = geom
        self.weight = weight.detach().contiguous().float()  # org_weight (quant_layer.py:41)
        self.bias = bias.detach().contiguous().float()
        self.hadamard = bool(hadamard)
        self.n_bits = int(n_bits)
        self.cin_src = next_pow2(geom.cin) if hadamard else geom.cin
        if hadamard:  # quant_layer.py:43-49
            padded = torch.zeros(geom.cout, self.cin_src, geom.k, geom.k, device=weight.device, dtype=torch.float32)
            padded[:, :geom.cin] = self.weight
            self.w_src = L.fwht_channel(padded)  # hadamard_weight
        else:
            self.w_src = self.weight
        self.delta_w = self.zp_w = self.delta_b = self.zp_b = None
        self.alpha_w = self.alpha_b = None
        self.codes_w = torch.empty_like(self.w_src)  # x_quant cache (quantizer.py:297)
        self.codes_b = torch.empty_like(self.bias)

    def set_bits(self, n_bits: int):
        if not 2 <= n_bits <= 8:
            raise AssertionError("bitwidth not supported")  # quantizer.py:96,237
        self.n_bits = int(n_bits)

    def init_scales(self):
        """UniformAffineQuantizer.init_quantization_scale('max', channel_wise) (quantizer.py:127-168)."""
        self.delta_w, self.zp_w = L.uaq_init_max(self.w_src, self.n_bits, True)
        self.delta_b, self.zp_b = L.uaq_init_max(self.bias, self.n_bits, True)

    def start_adaround(self):
        """AdaRoundQuantizer.__init__ (quantizer.py:259-319): fp16-rounded scales, alpha from the
        fractional part of w_src / delta."""
        self.delta_w = self.delta_w.detach().half().float().contiguous()
        self.zp_w = self.zp_w.detach().half().float().contiguous()
        self.delta_b = self.delta_b.detach().half().float().contiguous()
        self.zp_b = self.zp_b.detach().half().float().contiguous()
        self.alpha_w = L.adaround_init_alpha(self.w_src, self.delta_w)
        self.alpha_b = L.adaround_init_alpha(self.bias, self.delta_b)


def stage_descs(geoms: Sequence[StageGeom], n: int, h0: int, w0: int, use_tc: bool) -> List[L.ConvDesc]:
    """nq_conv_desc of every stage for a batch of n embeddings on an h0 x w0 grid.  Channel padding:
    16 for the input of a tensor-core stage (two 8-channel bf16 chunks per MMA), 8 for the head's input
    (pairs of fp32 float4), 4 on the FFMA path; a tensor-core stage's GEMM-N (rh*rw*cg) must be a
    multiple of 16 as well."""
    last = len(geoms) - 1
    in_pad = [(16 if use_tc else 4) if i < last else (8 if use_tc else 4) for i in range(last + 1)]
    out = []
    h, w = h0, w0
    cin_p = _pad(geoms[0].cin, in_pad[0])
    for i, g in enumerate(geoms):
        head = i == last
        cg = 4 if head else _pad(g.c_grp, in_pad[i + 1])
        if use_tc and not head and (g.rh * g.rw * cg) % 16:
            cg = _pad(g.c_grp, 16)
        out.append(L.ConvDesc(n, h, w, g.cin, cin_p, g.k, g.cout, g.rh, g.rw, g.c_grp, cg, 0 if head else _ACT[g.act]))
        h, w = h * g.rh, w * g.rw
        cin_p = cg
    return out


class _Plan:
    """Buffers of one (batch, grid) shape."""

    def __init__(self, eng: "DecoderEngine", n: int, h0: int, w0: int, train: bool):
        dev = eng.device
        self.n, self.h0, self.w0, self.train = n, h0, w0, train
        self.desc: List[L.ConvDesc] = []
        self.x: List[torch.Tensor] = []  # stage inputs (x[i+1] is stage i's activated output)
        self.z: List[Optional[torch.Tensor]] = []  # what backward needs of the pre-activation: GELU'(z) (train only)
        self.dz: List[Optional[torch.Tensor]] = []
        last = len(eng.geoms) - 1
        self.desc = stage_descs(eng.geoms, n, h0, w0, eng.use_tc)
        if train and os.environ.get("NQ_SAVE_ACT_GRAD", "1") != "0":
            # act 2: the forward epilogue evaluates GELU' next to GELU (one exponential for both) and keeps it in the
            # z buffer; dgrad then multiplies instead of re-evaluating erfc / exp per element.
            for d in self.desc:
                if d.act == _ACT["gelu"]:
                    d.act = _ACT_SAVED_GRAD

        def act_buf(*shape):
            # tensor-core engine: "split-bf16" storage (hi plane, lo plane; same bytes as fp32); FFMA engine: fp32
            if eng.use_tc:
                return torch.zeros((2,) + shape, device=dev, dtype=torch.bfloat16)
            return torch.zeros(shape, device=dev)

        self.x.append(act_buf(n, h0, w0, self.desc[0].cin_p))
        h, w = h0, w0
        for i, (g, d) in enumerate(zip(eng.geoms, self.desc)):
            if train:
                self.dz.append(act_buf(n, h, w, _pad(d.nout_p, 8) if eng.use_tc else d.nout_p))
            if i == last:
                break
            h, w = h * g.rh, w * g.rw
            self.x.append(act_buf(n, h, w, d.cg))
            self.z.append(torch.empty(n, h, w, d.cg, device=dev) if (train and g.act != "none") else None)
        self.H, self.W = h, w
        self.img = torch.empty(n, 3, h, w, device=dev)
        self.loss = torch.zeros(1, device=dev)
        # tensor-core plans: forward with 1 or 2 weight planes, data gradient
        self.tc_fwd = {}
        self.tc_dgrad = []
        if eng.use_tc:
            for i in range(last + 1):  # the head too: N = 3 real columns padded to 16
                for bpl in (1, 2):
                    pl = L.TcPlan()
                    L.check(L.lib.nq_tc_plan_conv(C.byref(self.desc[i]), 0, eng.fwd_a_planes, bpl, C.byref(pl)), "nq_tc_plan_conv")
                    pl.cluster = eng.cluster
                    self.tc_fwd[(i, bpl)] = pl
                pl = L.TcPlan()
                if train and i > 0:
                    L.check(L.lib.nq_tc_plan_conv(C.byref(self.desc[i]), 1, eng.bwd_a_planes, eng.bwd_b_planes, C.byref(pl)),
                            "nq_tc_plan_conv")
                    pl.cluster = eng.cluster
                self.tc_dgrad.append(pl)
            self.tc_head_dgrad = self.tc_dgrad[last] if train and last > 0 else None


class DecoderEngine:
    """Executes a list of QuantStage on one device.  mode: 'off' | 'uaq' | 'ada'."""

    def __init__(self, stages: Sequence[QuantStage]):
        self.stages = list(stages)
        self.geoms = [s.geom for s in self.stages]
        for g in self.geoms[:-1]:
            if g.act not in _ACT:
                raise NotImplementedError(f"activation {g.act!r}")
        hg = self.geoms[-1]
        if (hg.k, hg.cout, hg.rh, hg.rw) != (3, 3, 1, 1) or hg.act not in _HEAD:
            raise NotImplementedError("head must be a 3x3 conv to 3 channels with tanh/sigmoid OutImg")
        self.device = self.stages[0].weight.device
        if self.device.type != "cuda":
            raise L.NqError("DecoderEngine needs CUDA tensors (no CPU fallback)")
        self.mode = "uaq"
        self.stage_state = None  # optional per-stage (mode, soft_w, soft_b) override, see _round_modes
        self.soft_w = False
        self.soft_b = False
        # Convolution path: tcgen05 tensor cores (default) or the exact-fp32 FFMA kernels (NQ_CONV=simt).
        # Both are CUDA kernels of libnq_sm100.so; neither is a fallback for a missing library.
        self.use_tc = os.environ.get("NQ_CONV", "tc").lower() != "simt"
        # bf16 planes per operand (1 = hi only, 2 = hi + lo): forward activations, backward gradients/weights
        self.fwd_a_planes = int(os.environ.get("NQ_FWD_A_PLANES", "2"))
        self.bwd_a_planes = int(os.environ.get("NQ_BWD_A_PLANES", "2"))
        self.bwd_b_planes = int(os.environ.get("NQ_BWD_B_PLANES", "2"))
        self.wgrad_tc = os.environ.get("NQ_WGRAD", "tc").lower() != "simt"
        # weight-gradient operand planes (input activations, output gradients): default = the backward setting
        self.wg_a_planes = int(os.environ.get("NQ_WG_A_PLANES", str(self.bwd_a_planes)))
        self.wg_b_planes = int(os.environ.get("NQ_WG_B_PLANES", str(self.bwd_b_planes)))
        # head forward (measured at 2 x 640 x 1280 in a calibration step): tap-expanded tensor-core kernel 0.115 ms,
        # FFMA strip kernel 0.26 ms, generic tensor-core kernel with the head epilogue 0.29 ms
        head = os.environ.get("NQ_HEAD", "tapexp").lower()
        self.head_tc = head == "tc"
        # default: tensor cores with the taps as GEMM columns + nine shifted adds (nq_head_tc.cu, heads of <= 64 input
        # channels); NQ_HEAD=simt: the FFMA kernels; NQ_HEAD=tc: the generic tensor-core kernel with the head epilogue
        self.head_tapexp = head == "tapexp"
        # head weight gradient: tap-expanded kernel (default, 0.14 ms) or the generic tensor-core wgrad kernel (NQ_HEAD_WG=generic, 0.22 ms)
        self.head_wg_tapexp = os.environ.get("NQ_HEAD_WG", "tapexp").lower() == "tapexp"
        self.cluster = int(os.environ.get("NQ_CLUSTER", "2"))  # CTAs sharing a weight stream by TMA multicast
        self._plans: Dict[Tuple[int, int, int, bool], _Plan] = {}
        self._packed = None  # per-stage (wk, wt, bias_packed, deq_w scratch, deq_b scratch)
        self._weights_valid = False
        self._wt_valid = False
        self._packed_for = None  # (n, h0, w0) of the plan whose operand layouts the packed weights currently have
        self._retired = []  # outgrown packed-operand buffers: captured graphs may still pack into / read from them
        self.reg_sum = torch.zeros(1, device=self.device)
        self._grad = None
        self.sm = L.lib.nq_sm_count()
        self.launches = 0  # kernels launched through the C ABI (bench.py reports it)
        self._prof = None  # list of (name, flops, start_event, stop_event) while kernel_profile() runs
        self.dtype_name = "bf16x2 split operands, fp32 accumulate (tcgen05)" if self.use_tc else "f32"

    # ------------------------------------------------------------------ profiling
    def _run(self, name: str, d, fn, *args) -> int:
        """Launch one convolution-class kernel; under kernel_profile() bracket it with CUDA events on
        the launch stream.  d: the stage's ConvDesc (for the algorithmic FLOP count 2*M*N*K)."""
        if self._prof is None:
            return fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = fn(*args)
        e1.record()
        flops = 2.0 * d.n * d.h * d.w * d.cout * d.cin * d.ksize * d.ksize
        self._prof.append((name, flops, e0, e1))
        return st

    def kernel_profile(self, step_fn, reps: int = 3):
        """Per-kernel device time of the convolution launches of `step_fn` (CUDA events on the launch
        stream, median of `reps` runs).  Returns a list of dicts sorted by total time."""
        step_fn()
        torch.cuda.synchronize()
        self._prof = []
        for _ in range(reps):
            step_fn()
        torch.cuda.synchronize()
        rows, self._prof = self._prof, None
        agg = {}
        for name, flops, e0, e1 in rows:
            agg.setdefault(name, ([], flops))[0].append(e0.elapsed_time(e1))
        out = []
        for k, (ts, flops) in agg.items():
            ts.sort()
            ms = ts[len(ts) // 2]  # median: one slow repetition (a first-touch page fault, a clock dip) must not pick the "dominant" kernel
            out.append({"kernel": k, "ms": ms, "flops": flops, "tflops": flops / (ms * 1e-3) / 1e12})
        out.sort(key=lambda r: -r["ms"])
        return out

    # ------------------------------------------------------------------ bookkeeping
    def invalidate(self):
        """Call after any change to weights / quantiser state (Q9: cached packed weights stay legal
        only while nothing changed)."""
        self._weights_valid = False
        self._wt_valid = False

    def avg_bits(self) -> float:
        """quant_model.py:58-72."""
        num = sum(s.n_bits * (s.weight.numel() + s.bias.numel()) for s in self.stages)
        return num / sum(s.weight.numel() + s.bias.numel() for s in self.stages)

    def set_bitwidth(self, bits: Sequence[int]) -> float:
        for s, b in zip(self.stages, bits):
            s.set_bits(b)
        self.invalidate()
        return self.avg_bits()

    def init_scales(self):
        for s in self.stages:
            s.init_scales()
            self.launches += 2
        self.invalidate()

    def start_adaround(self):
        for s in self.stages:
            s.start_adaround()
            self.launches += 2
        self.mode, self.soft_w, self.soft_b = "ada", True, True
        self.invalidate()

    def plan(self, n: int, h0: int, w0: int, train: bool) -> _Plan:
        key = (n, h0, w0, train)
        if key not in self._plans:
            self._plans[key] = _Plan(self, n, h0, w0, train)
        return self._plans[key]

    def _alloc_packed(self, p: _Plan):
        if self._packed is not None:
            return
        self._packed = []
        self._tcw = []  # per non-head stage: (wpk_fwd bytes, wpk_dgrad bytes, scale_packed)
        last = len(self.stages) - 1
        for i, (s, d) in enumerate(zip(self.stages, p.desc)):
            tc = self.use_tc
            head_simt_fwd = tc and i == last and not self.head_tc
            wk = None if (tc and not head_simt_fwd) else torch.zeros(d.kdim, d.nout_p, device=self.device)
            wt = None if tc else torch.zeros(d.ksize * d.ksize * d.nout_p, d.cin_p, device=self.device)
            bp = torch.zeros(d.nout_p, device=self.device)
            deq_w = torch.empty_like(s.w_src)
            deq_b = torch.empty_like(s.bias)
            self._packed.append((wk, wt, bp, deq_w, deq_b))
            if tc:
                nb_f = p.tc_fwd[(i, 2)].wpk_bytes
                # the dgrad buffer is sized from a worst-case (2-plane) plan so that eval-only plans can share it
                pl = L.TcPlan()
                nb_d = 0
                if i > 0:
                    L.check(L.lib.nq_tc_plan_conv(C.byref(d), 1, 2, 2, C.byref(pl)), "nq_tc_plan_conv")
                    nb_d = pl.wpk_bytes
                self._tcw.append((torch.zeros(nb_f, dtype=torch.uint8, device=self.device),
                                  torch.zeros(max(nb_d, 16), dtype=torch.uint8, device=self.device),
                                  torch.ones(d.nout_p, device=self.device)))
            else:
                self._tcw.append(None)
        self._head_dgrad = self._tcw[last][1] if (self.use_tc and last > 0) else None

    def _fit_packed(self, p: _Plan):
        """The packed-operand buffers were sized by the first plan of this engine.  The layout the C library chooses
        depends on the number of pixel tiles, i.e. on the batch size: CTA pairs with the two weight planes side by side
        take 1.5x the bytes of the plain layout (nq_tc_plan.wpk_bytes), so a later plan of a LARGER batch can need more
        than a smaller first one reserved (decode one frame, then calibrate on two).  Grow what is too small; the old
        buffer stays alive because a captured graph of another plan packs into it and reads from it."""
        last = len(self.stages) - 1
        for i, tcw in enumerate(self._tcw):
            if tcw is None:
                continue
            wpk_f, wpk_d, scale_p = tcw
            need_f = max(p.tc_fwd[(i, 1)].wpk_bytes, p.tc_fwd[(i, 2)].wpk_bytes)
            need_d = p.tc_dgrad[i].wpk_bytes if (p.train and i > 0) else 0
            if need_f <= wpk_f.numel() and need_d <= wpk_d.numel():
                continue
            self._retired.append(tcw)
            if need_f > wpk_f.numel():
                wpk_f = torch.zeros(need_f, dtype=torch.uint8, device=self.device)
            if need_d > wpk_d.numel():
                wpk_d = torch.zeros(need_d, dtype=torch.uint8, device=self.device)
            self._tcw[i] = (wpk_f, wpk_d, scale_p)
            if i == last and last > 0:
                self._head_dgrad = wpk_d

    # ------------------------------------------------------------------ weights
    def prepare_weights(self, p: _Plan, need_wt: bool, reg_b: Optional[float] = None):
        """Fake-quantise every stage's weight and bias and pack them for the convolutions
        (quant_layer.py:68-78).  reg_b: also accumulate sum(1 - |2h-1|^b) over the WEIGHT alphas into
        self.reg_sum (calib_model.py:39-47; bias alphas are excluded there)."""
        self._alloc_packed(p)
        if self.use_tc:
            self._fit_packed(p)
        st = L.stream()
        if reg_b is not None:
            self.reg_sum.zero_()
        if not hasattr(self, "_fwd_bpl"):
            self._fwd_bpl = [2] * len(self.stages)
        if self.mode == "packed":
            # decoder rebuilt from a packed artefact: the integer codes ARE the state (no weights, no rounding); the
            # de-quantised copies serve the rotated stages and the FFMA head, the tensor-core stages pack the codes
            for s, (_, _, _, deq_w, deq_b) in zip(self.stages, self._packed):
                deq_b.copy_(s.bias)
                torch.mul(s.codes_w - s.zp_w, s.delta_w, out=deq_w)  # quantizer.py:299 (same two roundings as the kernel)
        elif self.mode != "off":
            # every weight and bias quantiser of the decoder in one multi-tensor launch
            tasks = []
            for i, (s, (_, _, _, deq_w, deq_b)) in enumerate(zip(self.stages, self._packed)):
                if self._stage_off(i):
                    continue
                mw, mb = self._round_modes(i)
                tasks.append(self._fq_task(s.w_src, s.alpha_w, s.delta_w, s.zp_w, s.n_bits, mw, s.codes_w, deq_w,
                                           reg_b is not None and mw == ROUND_SOFT))
                tasks.append(self._fq_task(s.bias, s.alpha_b, s.delta_b, s.zp_b, s.n_bits, mb, s.codes_b, deq_b, False))
            if tasks:
                arr = (L.FqTask * len(tasks))(*tasks)
                L.check(L.lib.nq_fakequant_fwd_multi(arr, len(tasks), L.ptr(self.reg_sum) if reg_b is not None else None,
                                                     float(reg_b or 0.0), st), "nq_fakequant_fwd_multi")
                self.launches += (len(tasks) + L.MULTI_MAX - 1) // L.MULTI_MAX
        packs = []
        for i, (s, d, (wk, wt, bp, deq_w, deq_b)) in enumerate(zip(self.stages, p.desc, self._packed)):
            if self._stage_off(i):
                w_for_conv, b_for_conv, cin_src = s.weight, s.bias, s.geom.cin
            else:
                if s.hadamard:  # quant_layer.py:71: rotate back, keep the first C_in channels
                    L.fwht_channel(deq_w, out=deq_w)
                    self.launches += 1
                w_for_conv, b_for_conv, cin_src = deq_w, deq_b, s.cin_src
            if self._tcw[i] is None:
                L.check(L.lib.nq_pack_weight(C.byref(d), L.ptr(w_for_conv), cin_src, L.ptr(b_for_conv), L.ptr(wk),
                                             L.ptr(wt) if need_wt else None, L.ptr(bp), st), "nq_pack_weight")
                self.launches += 1
                continue
            wpk_f, wpk_d, scale_p = self._tcw[i]
            if i == len(self.stages) - 1 and not self.head_tc:
                # head forward on the FFMA kernel: fp32 packed weights; its data gradient still runs on the tensor cores
                L.check(L.lib.nq_pack_weight(C.byref(d), L.ptr(w_for_conv), cin_src, L.ptr(b_for_conv), L.ptr(wk), None, L.ptr(bp), st),
                        "nq_pack_weight")
                self.launches += 1
                if need_wt and i > 0:
                    packs.append(L.TcPackTask(C.pointer(d), C.pointer(p.tc_dgrad[i]), L.ptr(w_for_conv), None, wpk_d.data_ptr(),
                                              None, None, None, None, cin_src, 0, 0, 0))
                continue
            # forward operand: integer weights (codes - zero_point), exact in ONE bf16 plane, whenever the
            # codes are integers and are what the conv multiplies (no rotation in between); the step size
            # is applied per output channel in the epilogue.  Otherwise the de-quantised fp32 weights,
            # split into two bf16 planes.
            integer = not self._stage_off(i) and not s.hadamard
            exact1 = integer and self._round_modes(i)[0] != ROUND_SOFT
            bpl = 1 if exact1 else 2
            self._fwd_bpl[i] = bpl
            # forward operand + per-column epilogue vectors: one task; data-gradient operand: another; all stages'
            # tasks go out in one multi-pack launch below
            if integer:
                ds = 1 if s.delta_w.numel() > 1 else 0
                packs.append(L.TcPackTask(C.pointer(d), C.pointer(p.tc_fwd[(i, bpl)]), L.ptr(s.codes_w), L.ptr(s.zp_w), wpk_f.data_ptr(),
                                          L.ptr(s.delta_w), L.ptr(b_for_conv), L.ptr(scale_p), L.ptr(bp), s.cin_src, ds, ds, 0))
            else:
                packs.append(L.TcPackTask(C.pointer(d), C.pointer(p.tc_fwd[(i, bpl)]), L.ptr(w_for_conv), None, wpk_f.data_ptr(),
                                          None, L.ptr(b_for_conv), L.ptr(scale_p), L.ptr(bp), cin_src, 0, 0, 0))
            if need_wt and i > 0:
                packs.append(L.TcPackTask(C.pointer(d), C.pointer(p.tc_dgrad[i]), L.ptr(w_for_conv), None, wpk_d.data_ptr(),
                                          None, None, None, None, cin_src, 0, 0, 0))
        if packs:
            arr = (L.TcPackTask * len(packs))(*packs)
            L.check(L.lib.nq_tc_pack_multi(arr, len(packs), st), "nq_tc_pack_multi")
            self.launches += (len(packs) + L.MULTI_MAX - 1) // L.MULTI_MAX
        self._weights_valid = True
        self._wt_valid = need_wt
        self._packed_for = (p.n, p.h0, p.w0)

    def _stage_off(self, i: int) -> bool:
        """Stage i runs on its full-precision weights: the whole decoder (mode 'off') or this stage alone ('off' in
        `stage_state`: quantize_model_till, data_utils.py:261-272, quantises a prefix of the decoder only)."""
        if self.mode == "off":
            return True
        ss = getattr(self, "stage_state", None)
        return bool(ss) and ss[i][0] == "off"

    def _round_modes(self, i: int):
        """(weight, bias) rounding mode of stage i.  `stage_state` (set by the module binding when a decoder mixes
        plain and AdaRound quantisers, e.g. after block-wise reconstruction of some blocks) overrides the global mode."""
        st = self.stage_state[i] if getattr(self, "stage_state", None) else (self.mode, self.soft_w, self.soft_b)
        if st[0] == "packed":
            return ROUND_HARD, ROUND_HARD
        if st[0] == "uaq":
            return ROUND_NEAREST, ROUND_NEAREST
        return (ROUND_SOFT if st[1] else ROUND_HARD), (ROUND_SOFT if st[2] else ROUND_HARD)

    @staticmethod
    def _fq_task(x, alpha, delta, zp, n_bits, mode, codes, deq, want_reg) -> "L.FqTask":
        cw = delta.numel() > 1
        rows, row_len = (delta.numel(), x.numel() // delta.numel()) if cw else (1, x.numel())
        return L.FqTask(L.ptr(x), L.ptr(alpha) if mode != ROUND_NEAREST else None, L.ptr(delta), L.ptr(zp), L.ptr(codes),
                        L.ptr(deq), rows, row_len, int(cw), n_bits, mode, int(bool(want_reg)))

    def _fq(self, x, alpha, delta, zp, n_bits, mode, codes, deq, reg_sum, reg_b):
        rows, row_len, d_stride = (delta.numel(), x.numel() // delta.numel(), 1) if delta.numel() > 1 else (1, x.numel(), 0)
        L.check(L.lib.nq_fakequant_fwd(L.ptr(x), L.ptr(alpha) if mode != ROUND_NEAREST else None, L.ptr(delta),
                                       L.ptr(zp), rows, row_len, d_stride, n_bits, mode, L.ptr(codes), L.ptr(deq),
                                       L.ptr(reg_sum), float(reg_b), L.stream()), "nq_fakequant_fwd")
        self.launches += 1

    # ------------------------------------------------------------------ forward
    def forward(self, embed: torch.Tensor, *, train: bool = False, target: Optional[torch.Tensor] = None,
                p_norm: float = 2.0, mean_pixels: Optional[float] = None, reuse_weights: bool = False,
                reg_b: Optional[float] = None, want_img: bool = True) -> torch.Tensor:
        """embed: (n, C0, h0, w0) NCHW as the reference's decode() takes it.  Returns the (n, 3, H, W)
        frame.  With `target` the loss sum is accumulated into plan.loss (read via last_loss())."""
        n, c0, h0, w0 = embed.shape
        if c0 != self.geoms[0].cin:
            raise L.NqError(f"embedding has {c0} channels, decoder stem expects {self.geoms[0].cin}")
        p = self.plan(n, h0, w0, train)
        self._last_plan = p
        st = L.stream()
        # packed operands are reusable only by a plan of the same geometry: the layout (stage size, CTA pairs, planes side
        # by side) is chosen per plan and depends on the number of pixel tiles, hence on the batch size
        if not (reuse_weights and self._weights_valid and (self._wt_valid or not train) and self._packed_for == (n, h0, w0)):
            self.prepare_weights(p, need_wt=train, reg_b=reg_b)
        embed = embed.detach().contiguous().float()
        if self.use_tc:
            L.check(L.lib.nq_nchw_to_split(L.ptr(embed), p.x[0].data_ptr(), n, c0, h0, w0, p.desc[0].cin_p, st), "nq_nchw_to_split")
        else:
            L.check(L.lib.nq_nchw_to_nhwc(L.ptr(embed), L.ptr(p.x[0]), n, c0, h0, w0, p.desc[0].cin_p, st), "nq_nchw_to_nhwc")
        self.launches += 1
        last = len(self.stages) - 1
        for i in range(last):
            wk, _, bp, _, _ = self._packed[i]
            z = p.z[i] if train else None
            if self._tcw[i] is None:
                L.check(self._run(f"conv_fwd[{i}]", p.desc[i], L.lib.nq_conv_fwd, C.byref(p.desc[i]), L.ptr(p.x[i]), L.ptr(wk),
                                  L.ptr(bp), L.ptr(z), L.ptr(p.x[i + 1]), st), "nq_conv_fwd")
            else:
                wpk_f, _, scale_p = self._tcw[i]
                L.check(self._run(f"conv_fwd[{i}]", p.desc[i], L.lib.nq_tc_conv_fwd, C.byref(p.desc[i]),
                                  C.byref(p.tc_fwd[(i, self._fwd_bpl[i])]), p.x[i].data_ptr(), wpk_f.data_ptr(), L.ptr(scale_p),
                                  L.ptr(bp), L.ptr(z), p.x[i + 1].data_ptr(), st), "nq_tc_conv_fwd")
            self.launches += 1
        wk, _, bp, _, _ = self._packed[last]
        tapexp = self.use_tc and not self.head_tc and self.head_tapexp and p.desc[last].cin_p <= 64
        target_u8 = None
        if target is not None:
            target = target.detach().contiguous()
            if tuple(target.shape) != (n, 3, p.H, p.W):
                raise L.NqError(f"target shape {tuple(target.shape)} != {(n, 3, p.H, p.W)}")
            if target.dtype == torch.uint8:
                # frames as the data set stores them (videosets/datasets.py:8-54): value / 255 is evaluated on the device,
                # inside the head kernel where it has a uint8 variant, else by the ingest kernel into a plan buffer
                if not target.is_cuda:
                    raise L.NqError("neuroquant_b200 kernels need CUDA tensors (there is no CPU fallback)")
                if tapexp:
                    target_u8 = target
                else:
                    if not hasattr(p, "tgt_f32"):
                        p.tgt_f32 = torch.empty(n, 3, p.H, p.W, device=self.device)
                    L.check(L.lib.nq_u8_to_f32(target.data_ptr(), L.ptr(p.tgt_f32), target.numel(), st), "nq_u8_to_f32")
                    self.launches += 1
                    target = p.tgt_f32
            else:
                target = target.float()
            p.loss.zero_()
            mp = float(mean_pixels if mean_pixels is not None else n * p.H * p.W)
        else:
            mp = 1.0
        want_dz = train and target is not None
        if target_u8 is not None:
            L.check(self._run("head_fwd_loss", p.desc[last], L.lib.nq_head_fwd_loss_tapexp_u8, C.byref(p.desc[last]),
                              p.x[last].data_ptr(), L.ptr(wk), L.ptr(bp), _HEAD[self.geoms[last].act], target_u8.data_ptr(),
                              float(p_norm), mp, L.ptr(p.img) if want_img else None, L.ptr(p.loss),
                              p.dz[last].data_ptr() if want_dz else None, st), "nq_head_fwd_loss_tapexp_u8")
        elif self.use_tc and not self.head_tc:
            head_fn = L.lib.nq_head_fwd_loss_tapexp if tapexp else L.lib.nq_head_fwd_loss_split
            L.check(self._run("head_fwd_loss", p.desc[last], head_fn, C.byref(p.desc[last]),
                              p.x[last].data_ptr(), L.ptr(wk), L.ptr(bp), _HEAD[self.geoms[last].act], L.ptr(target),
                              float(p_norm), mp, L.ptr(p.img) if (want_img or target is None) else None,
                              L.ptr(p.loss) if target is not None else None,
                              p.dz[last].data_ptr() if want_dz else None, st), "nq_head_fwd_loss_split")
        elif self.use_tc:
            wpk_f, _, scale_p = self._tcw[last]
            L.check(self._run("head_fwd_loss", p.desc[last], L.lib.nq_tc_head_fwd_loss, C.byref(p.desc[last]),
                              C.byref(p.tc_fwd[(last, self._fwd_bpl[last])]), p.x[last].data_ptr(), wpk_f.data_ptr(),
                              L.ptr(scale_p), L.ptr(bp), _HEAD[self.geoms[last].act], L.ptr(target), float(p_norm), mp,
                              L.ptr(p.img) if (want_img or target is None) else None,
                              L.ptr(p.loss) if target is not None else None,
                              p.dz[last].data_ptr() if want_dz else None, st), "nq_tc_head_fwd_loss")
        else:
            L.check(self._run("head_fwd_loss", p.desc[last], L.lib.nq_head_fwd_loss, C.byref(p.desc[last]), L.ptr(p.x[last]),
                              L.ptr(wk), L.ptr(bp), _HEAD[self.geoms[last].act], L.ptr(target), float(p_norm), mp,
                              L.ptr(p.img) if (want_img or target is None) else None,
                              L.ptr(p.loss) if target is not None else None,
                              L.ptr(p.dz[last]) if want_dz else None, st), "nq_head_fwd_loss")
        self.launches += 1
        self._mean_pixels = mp
        return p.img

    def last_loss(self) -> torch.Tensor:
        """lp_loss of the last forward-with-target: sum / mean_pixels (device scalar, no sync)."""
        return self._last_plan.loss / self._mean_pixels

    # ------------------------------------------------------------------ backward
    def _grad_buffers(self):
        if self._grad is None:
            sizes = []
            for s in self.stages:
                sizes += [s.w_src.numel(), s.bias.numel()]
            flat = torch.zeros(sum(sizes), device=self.device)
            views, o = [], 0
            for s in self.stages:
                nw, nb = s.w_src.numel(), s.bias.numel()
                views.append((flat[o:o + nw].view_as(s.w_src), flat[o + nw:o + nw + nb]))
                o += nw + nb
            self._grad = (flat, views)
        return self._grad

    def _wg_plan(self, d: L.ConvDesc):
        pl = L.TcWgradPlan()
        st = L.lib.nq_tc_plan_wgrad(C.byref(d), self.wg_a_planes, self.wg_b_planes, C.byref(pl))
        return pl if st == 0 else None

    def _wgrad_splits(self, d: L.ConvDesc) -> int:
        rows = d.kdim + 4
        n = d.nout_p
        bn = 64 if (n <= 64 or 0 < n % 128 <= 64) else 128
        tiles = ((rows + 127) // 128) * ((n + bn - 1) // bn)
        pix = d.n * d.h * d.w
        s = max(1, min((4 * self.sm + tiles - 1) // tiles, pix // 64))
        return int(s)

    def backward(self) -> torch.Tensor:
        """Back-propagate the loss of the last train-mode forward down to dL/d(dequantised weight) and
        dL/d(dequantised bias) in reference layout (rotated domain when --hadamard).  Returns the flat
        gradient buffer (all stages, W then b), ready for an all-reduce."""
        p = self._last_plan
        assert p.train and self._wt_valid
        st = L.stream()
        flat, views = self._grad_buffers()
        last = len(self.stages) - 1
        if not hasattr(p, "dwk"):
            p.dwk, p.ws = [], []
            p.head_desc16 = None
            for i, d in enumerate(p.desc):
                if self.use_tc:
                    # every stage's weight gradient on the tensor cores (the head's 3 output channels are
                    # padded to N = 16; its dz is stored with 8 channels)
                    pl = self._wg_plan(d)
                    if pl is None:
                        raise NotImplementedError(
                            f"stage {i}: no tensor-core weight-gradient plan for cin_p = {d.cin_p}, k = {d.ksize} "
                            "(nq_tc_plan_wgrad slices the input channels until the accumulators fit TMEM and found no slicing)")
                    p.dwk.append(torch.empty(d.kdim + 4, pl.N, device=self.device))
                    p.ws.append((torch.empty(pl.workspace_floats, device=self.device), pl))
                    if i == last:
                        p.head_desc16 = L.ConvDesc(d.n, d.h, d.w, d.cin, d.cin_p, d.ksize, d.cout, 1, 1, d.c_grp, pl.N, 0)
                elif i == last:
                    p.dwk.append(torch.empty(d.kdim + 4, d.nout_p, device=self.device))
                    blocks = L.lib.nq_head_wgrad_blocks(C.byref(d))
                    p.ws.append((torch.empty(blocks * (d.kdim + 4) * 4, device=self.device), 0))
                else:
                    p.dwk.append(torch.empty(d.kdim + 4, d.nout_p, device=self.device))
                    sp = self._wgrad_splits(d)
                    p.ws.append((torch.empty(sp * (d.kdim + 4) * d.nout_p, device=self.device) if sp > 1 else None, sp))
        finish = []
        for i in range(last, -1, -1):
            d = p.desc[i]
            _, wt, _, _, _ = self._packed[i]
            ws, sp = p.ws[i]
            if isinstance(sp, L.TcWgradPlan) and i == last and self.head_wg_tapexp and d.cin_p <= 64 and p.head_desc16 is not None \
                    and p.head_desc16.cg == 16:
                # head: tap-expanded weight gradient (nq_head_tc.cu); same partial layout, finished with the others
                if not hasattr(p, "head_wg_ws"):
                    n_sp = int(L.lib.nq_head_wgrad_tapexp_splits(C.byref(d)))
                    p.head_wg_ws = (torch.empty(n_sp * (9 * d.cin_p + 4) * 16, device=self.device), n_sp)
                hws, n_sp = p.head_wg_ws
                L.check(self._run(f"conv_wgrad[{i}]", d, L.lib.nq_head_wgrad_tapexp, C.byref(d), p.x[i].data_ptr(), p.dz[i].data_ptr(),
                                  L.ptr(hws), hws.numel(), st), "nq_head_wgrad_tapexp")
                self.launches += 1
                gw_, gb_ = views[i]
                finish.append(L.WgFinishTask(C.pointer(p.head_desc16), L.ptr(hws), L.ptr(gw_), L.ptr(gb_), n_sp, 16,
                                             self.stages[i].cin_src, 0))
            elif isinstance(sp, L.TcWgradPlan):
                # partial sums only (dwk = NULL): one multi-stage launch after the loop reduces and unpacks them all
                L.check(self._run(f"conv_wgrad[{i}]", d, L.lib.nq_tc_conv_wgrad, C.byref(d), C.byref(sp), p.x[i].data_ptr(),
                                  p.dz[i].data_ptr(), None, L.ptr(ws), ws.numel(), st), "nq_tc_conv_wgrad")
                self.launches += 1
                du = p.head_desc16 if (i == last and p.head_desc16 is not None) else d
                gw_, gb_ = views[i]
                finish.append(L.WgFinishTask(C.pointer(du), L.ptr(ws), L.ptr(gw_), L.ptr(gb_), sp.psplits, sp.N,
                                             self.stages[i].cin_src, 0))
            elif i == last:
                L.check(self._run("head_wgrad", d, L.lib.nq_head_wgrad, C.byref(d), L.ptr(p.x[i]), L.ptr(p.dz[i]), L.ptr(p.dwk[i]),
                                  L.ptr(ws), ws.numel(), st), "nq_head_wgrad")
                self.launches += 2
            else:
                L.check(self._run(f"conv_wgrad[{i}]", d, L.lib.nq_conv_wgrad, C.byref(d), L.ptr(p.x[i]), L.ptr(p.dz[i]),
                                  L.ptr(p.dwk[i]), L.ptr(ws), ws.numel() if ws is not None else 0, sp, st), "nq_conv_wgrad")
                self.launches += 2 if sp > 1 else 1
            if i > 0:
                g_prev = self.geoms[i - 1]
                if self.use_tc:
                    pl, wpk = (p.tc_head_dgrad, self._head_dgrad) if i == last else (p.tc_dgrad[i], self._tcw[i][1])
                    if pl.ksplit > 1 and (not hasattr(p, "dgrad_ws") or p.dgrad_ws.numel() < pl.workspace_floats):
                        p.dgrad_ws = torch.empty(int(pl.workspace_floats), device=self.device)  # shared by the split-K stages
                    dws = p.dgrad_ws if pl.ksplit > 1 else None
                    L.check(self._run(f"conv_dgrad[{i}]", d, L.lib.nq_tc_conv_dgrad, C.byref(d), C.byref(pl), p.dz[i].data_ptr(),
                                      wpk.data_ptr(), L.ptr(p.z[i - 1]), g_prev.rh, g_prev.rw, p.desc[i - 1].act,
                                      p.dz[i - 1].data_ptr(), L.ptr(dws), dws.numel() if dws is not None else 0, st),
                            "nq_tc_conv_dgrad")
                    self.launches += 1 if pl.ksplit > 1 else 0
                else:
                    L.check(self._run(f"conv_dgrad[{i}]", d, L.lib.nq_conv_dgrad, C.byref(d), L.ptr(p.dz[i]), L.ptr(wt),
                                      L.ptr(p.z[i - 1]), g_prev.rh, g_prev.rw, p.desc[i - 1].act, L.ptr(p.dz[i - 1]), st),
                            "nq_conv_dgrad")
                self.launches += 1
            s = self.stages[i]
            gw, gb = views[i]
            if not isinstance(sp, L.TcWgradPlan):
                L.check(L.lib.nq_unpack_wgrad(C.byref(d), L.ptr(p.dwk[i]), s.cin_src, L.ptr(gw), L.ptr(gb), st), "nq_unpack_wgrad")
                self.launches += 1
        if finish:
            arr = (L.WgFinishTask * len(finish))(*finish)
            L.check(L.lib.nq_tc_wgrad_finish_multi(arr, len(finish), st), "nq_tc_wgrad_finish_multi")
            self.launches += (len(finish) + L.MULTI_MAX - 1) // L.MULTI_MAX
        return flat

    def stage_input_grad(self, i: int) -> torch.Tensor:
        """dL/d(input of stage i) = dL/d(output of block i-1, after its activation and up-shuffle) of the last
        forward(train=True, target=...) + backward(), as (n, C, H, W) fp32.  backward() itself only keeps that gradient
        multiplied by the activation derivative (dz[i-1]); this re-runs stage i's data gradient with a linear
        predecessor.  Used for the output-gradient cache of the Fisher block losses (data_utils.py:209-258)."""
        p = self._last_plan
        assert p.train and hasattr(p, "dwk") and 0 < i < len(self.stages)
        st = L.stream()
        d, g_prev = p.desc[i], self.geoms[i - 1]
        last = len(self.stages) - 1
        out = torch.empty_like(p.dz[i - 1])
        if self.use_tc:
            pl, wpk = (p.tc_head_dgrad, self._head_dgrad) if i == last else (p.tc_dgrad[i], self._tcw[i][1])
            dws = p.dgrad_ws if pl.ksplit > 1 else None
            L.check(L.lib.nq_tc_conv_dgrad(C.byref(d), C.byref(pl), p.dz[i].data_ptr(), wpk.data_ptr(), None, g_prev.rh, g_prev.rw, 0,
                                           out.data_ptr(), L.ptr(dws), dws.numel() if dws is not None else 0, st), "nq_tc_conv_dgrad")
            out = out[0].float() + out[1].float()
        else:
            _, wt, _, _, _ = self._packed[i]
            L.check(L.lib.nq_conv_dgrad(C.byref(d), L.ptr(p.dz[i]), L.ptr(wt), None, g_prev.rh, g_prev.rw, 0, L.ptr(out), st),
                    "nq_conv_dgrad")
        n, h, w = out.shape[0], out.shape[1], out.shape[2]
        cg = out.shape[3] // (g_prev.rh * g_prev.rw)
        out = out.view(n, h, w, g_prev.rh, g_prev.rw, cg).permute(0, 5, 1, 3, 2, 4)
        return out.reshape(n, cg, h * g_prev.rh, w * g_prev.rw)[:, :g_prev.c_grp].contiguous()

    def stage_output_grad(self, i: int) -> torch.Tensor:
        """dL/d(convolution output of stage i), before its up-shuffle and activation, of the last forward(train=True,
        target=...) + backward(): the engine's own dz[i], returned as (n, C_out, h, w) fp32 in the reference's channel
        order (the engine's GEMM columns are sub-pixel major: column = s * cg + c, reference channel = c * rh*rw + s)."""
        p = self._last_plan
        assert p.train and hasattr(p, "dwk")
        g = self.geoms[i]
        dz = p.dz[i]
        dz = dz[0].float() + dz[1].float() if self.use_tc else dz
        n, h, w, ncol = dz.shape
        r2 = g.rh * g.rw
        cg = ncol // r2
        out = dz[..., :r2 * cg].view(n, h, w, r2, cg)[..., :g.c_grp].permute(0, 4, 3, 1, 2)
        return out.reshape(n, g.c_grp * r2, h, w).contiguous()

    def param_grads(self, grad_scale: float = 1.0, reg_w: float = 0.0, reg_b: float = 0.0, hyper: Optional[torch.Tensor] = None):
        """Chain the (possibly all-reduced) weight gradients through the rotation and the quantiser
        Jacobian.  Returns per stage (g_w, g_b): d_alpha (mode 'ada', soft) or d_delta (mode 'uaq').
        reg_w / reg_b add the rounding regulariser's gradient on the WEIGHT alphas only; with `hyper` (device
        array {reg_w, reg_b, ...}) they are read on the device instead (CUDA-graph replay)."""
        flat, views = self._grad_buffers()
        out = []
        if not hasattr(self, "_pg"):
            self._pg = {}
        for i, s in enumerate(self.stages):
            gw, gb = views[i]
            if s.hadamard:  # transpose of the (symmetric, orthonormal) rotation is the rotation
                L.fwht_channel(gw, out=gw)
                self.launches += 1
            if self.mode == "ada":
                key = ("a", i)
                if key not in self._pg:
                    self._pg[key] = (torch.empty_like(s.alpha_w), torch.empty_like(s.alpha_b))
                da_w, da_b = self._pg[key]
                if hyper is None:
                    L.fakequant_bwd(gw, s.w_src, s.alpha_w, s.delta_w, s.zp_w, s.n_bits, ROUND_SOFT, grad_scale, reg_w, reg_b, out=da_w)
                    L.fakequant_bwd(gb, s.bias, s.alpha_b, s.delta_b, s.zp_b, s.n_bits, ROUND_SOFT, grad_scale, 0.0, 0.0, out=da_b)
                else:
                    for g_, x_, a_, d_, z_, o_, use_reg in ((gw, s.w_src, s.alpha_w, s.delta_w, s.zp_w, da_w, 1),
                                                            (gb, s.bias, s.alpha_b, s.delta_b, s.zp_b, da_b, 0)):
                        rows, row_len, ds = (d_.numel(), x_.numel() // d_.numel(), 1) if d_.numel() > 1 else (1, x_.numel(), 0)
                        L.check(L.lib.nq_fakequant_bwd_soft_dev(L.ptr(g_), L.ptr(x_), L.ptr(a_), L.ptr(d_), L.ptr(z_), rows, row_len,
                                                                ds, s.n_bits, float(grad_scale), use_reg, L.ptr(hyper), L.ptr(o_),
                                                                L.stream()), "nq_fakequant_bwd_soft_dev")
                out.append((da_w, da_b))
            elif self.mode == "uaq":
                key = ("d", i)
                if key not in self._pg:
                    self._pg[key] = (torch.empty_like(s.delta_w), torch.empty_like(s.delta_b))
                dd_w, dd_b = self._pg[key]
                L.fakequant_bwd(gw, s.w_src, None, s.delta_w, s.zp_w, s.n_bits, ROUND_NEAREST, grad_scale, out=dd_w)
                L.fakequant_bwd(gb, s.bias, None, s.delta_b, s.zp_b, s.n_bits, ROUND_NEAREST, grad_scale, out=dd_b)
                out.append((dd_w, dd_b))
            else:
                raise L.NqError("param_grads needs quantisation on")
            self.launches += 2
        return out


def adaround_step_multi(eng: "DecoderEngine", opt: "AdamState", hyper: torch.Tensor, grad_scale: float = 1.0,
                        beta1=0.9, beta2=0.999, eps=1e-8) -> int:
    """d(loss + regulariser)/d alpha and Adam's update of alpha for every quantiser of the decoder in ONE launch
    (the graph-replayed form of param_grads() + AdamState.step_dev()).  `opt.params` must be the engine's
    [alpha_w, alpha_b] per stage, in order (CalibrationLoop.run_phase2)."""
    _, views = eng._grad_buffers()
    tasks = []
    k = 0
    for i, s in enumerate(eng.stages):
        gw, gb = views[i]
        if s.hadamard:  # transpose of the (symmetric, orthonormal) rotation is the rotation
            L.fwht_channel(gw, out=gw)
            eng.launches += 1
        for g_, x_, a_, d_, z_, use_reg in ((gw, s.w_src, s.alpha_w, s.delta_w, s.zp_w, 1), (gb, s.bias, s.alpha_b, s.delta_b, s.zp_b, 0)):
            if opt.params[k] is not a_:
                raise L.NqError("adaround_step_multi: optimiser parameters are not the engine's alpha tensors")
            cw = d_.numel() > 1
            rows, row_len = (d_.numel(), x_.numel() // d_.numel()) if cw else (1, x_.numel())
            tasks.append(L.AdaTask(L.ptr(g_), L.ptr(x_), L.ptr(a_), L.ptr(d_), L.ptr(z_), L.ptr(opt.m[k]), L.ptr(opt.v[k]),
                                   rows, row_len, int(cw), s.n_bits, use_reg, 0))
            k += 1
    arr = (L.AdaTask * len(tasks))(*tasks)
    L.check(L.lib.nq_adaround_step_multi(arr, len(tasks), float(grad_scale), beta1, beta2, eps, L.ptr(hyper), L.stream()),
            "nq_adaround_step_multi")
    n = (len(tasks) + L.MULTI_MAX - 1) // L.MULTI_MAX
    eng.launches += n
    return n


class AdamState:
    """torch.optim.Adam defaults (calib_model.py:134,195) over a list of tensors, one fused kernel each."""

    def __init__(self, params: Sequence[torch.Tensor], lr: float):
        self.params = list(params)
        self.lr = float(lr)
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    def step(self, grads: Sequence[torch.Tensor]) -> int:
        self.t += 1
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            L.adam_step(p, g, m, v, self.lr, self.t)
        return len(self.params)

    def hyper_of_next_step(self, beta1=0.9, beta2=0.999):
        """(lr / (1 - beta1^t), sqrt(1 - beta2^t)) of the step about to be taken; advances t."""
        self.t += 1
        return self.lr / (1.0 - beta1 ** self.t), (1.0 - beta2 ** self.t) ** 0.5

    def step_dev(self, grads: Sequence[torch.Tensor], hyper: torch.Tensor, beta1=0.9, beta2=0.999, eps=1e-8) -> int:
        """Adam step whose step size / bias correction come from the device array `hyper` (entries 2, 3)."""
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            L.check(L.lib.nq_adam_step_dev(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), beta1, beta2, eps, L.ptr(hyper),
                                           L.stream()), "nq_adam_step_dev")
        return len(self.params)
