"""Sensitivity criteria of bit_assign on the decoder engine (reference: methods/bit_assign.py:57-217).

omega        Omega = sum_batches v^T H_b v, H_b the Hessian of nn.MSELoss(decode(x_b), frame_b) w.r.t. the
             conv weights, v = W - Q(W).  The reference forms H v by a double backward pass; here
             v^T H v = d^2/d eps^2 L(w + eps v)|_0 is propagated FORWARD as a second-order jet
             (y, y', y'') through the decoder: 5 forward convolutions per stage on the tensor-core kernel
             (y*w, y'*w, y*v, y''*w, y'*v), the elementwise chain rule (nq_jet_act) and the MSE head
             (nq_jet_head).  No backward pass, no graph.  The per-layer terms the reference also logs,
             Omega_l = v_l^T (H v)_l (bit_assign.py:194-200), follow from the same kernel by polarisation:
             v_l^T H v = (Omega(v + v_l) - Omega(v - v_l)) / 4, two more jets per layer (omega_layers).
fisher_diag  sum_l sum (v_l^2 * g_l^2) with g the gradient accumulated over the batches: one engine
             forward/backward per batch and one fused multi-tensor reduction (nq_multi_dot).
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import List, Sequence

import torch

from . import _lib as L
from .engine import DecoderEngine, _ACT, _HEAD


class OmegaEvaluator:
    def __init__(self, engine: DecoderEngine):
        if engine.mode != "off":
            raise L.NqError("Omega is defined on the full-precision decoder (engine.mode == 'off')")
        self.eng = engine
        self.acc = torch.zeros(1, dtype=torch.float64, device=engine.device)
        self._v = None
        self._bufs = {}

    # ------------------------------------------------------------------ direction
    def set_direction(self, vecs: Sequence[torch.Tensor], n: int, h0: int, w0: int):
        """vecs: one (C_out, C_in, k, k) perturbation per stage (QuantModel.get_perturbation())."""
        eng = self.eng
        p = eng.plan(n, h0, w0, False)
        eng.prepare_weights(p, need_wt=False)
        st = L.stream()
        self._v, self._w, self._plans = [], [], []
        last = len(eng.stages) - 1
        for i, (s, d, v) in enumerate(zip(eng.stages, p.desc, vecs)):
            v = v.detach().contiguous().float()
            assert v.shape == s.weight.shape
            if eng.use_tc:
                # every stage, the head included (its 3 columns padded to 16), runs on the tensor-core forward kernel
                pl = p.tc_fwd[(i, 2)] if i < last else L.TcPlan()
                if i == last:
                    L.check(L.lib.nq_tc_plan_conv(C.byref(d), 0, eng.fwd_a_planes, 2, C.byref(pl)), "nq_tc_plan_conv")
                    pl.cluster = eng.cluster
                bufs = []
                for src in (s.weight, v):
                    buf = torch.zeros(pl.wpk_bytes, dtype=torch.uint8, device=eng.device)
                    L.check(L.lib.nq_tc_pack_weight(C.byref(d), C.byref(pl), L.ptr(src), s.geom.cin, None, 0, buf.data_ptr(), st),
                            "nq_tc_pack_weight")
                    bufs.append(buf)
                self._plans.append(pl)
                self._w.append(bufs[0])
                self._v.append(bufs[1])
            else:
                buf = torch.zeros(d.kdim, d.nout_p, device=eng.device)
                L.check(L.lib.nq_pack_weight(C.byref(d), L.ptr(v), s.geom.cin, None, L.ptr(buf), None, None, st), "nq_pack_weight")
                self._plans.append(None)
                self._w.append(eng._packed[i][0])
                self._v.append(buf)
        self._zero_bias = [torch.zeros(d.nout_p, device=eng.device) for d in p.desc]

    def _conv(self, p, i, x, use_v: bool, with_bias: bool, out):
        """out (fp32 pre-activation, shuffled grid) = conv(x; w or v) [+ bias]."""
        eng = self.eng
        d = copy.copy(p.desc[i])
        d.act = 0
        st = L.stream()
        bp = eng._packed[i][2]
        w = self._v[i] if use_v else self._w[i]
        if eng.use_tc:
            L.check(L.lib.nq_tc_conv_fwd(C.byref(d), C.byref(self._plans[i]), x.data_ptr(), w.data_ptr(), None,
                                         L.ptr(bp) if with_bias else None, L.ptr(out), None, st), "nq_tc_conv_fwd")
        else:
            L.check(L.lib.nq_conv_fwd(C.byref(d), L.ptr(x), L.ptr(w), L.ptr(bp if with_bias else self._zero_bias[i]), None,
                                      L.ptr(out), st), "nq_conv_fwd")
        eng.launches += 1

    # ------------------------------------------------------------------ one batch
    def add_batch(self, embed: torch.Tensor, target: torch.Tensor):
        eng = self.eng
        n, c0, h0, w0 = embed.shape
        p = eng.plan(n, h0, w0, False)
        key = (n, h0, w0)
        if key not in self._bufs:
            bufs = []
            for x, d in zip(p.x[1:], p.desc):
                shape = (n, d.h * d.rh, d.w * d.rw, d.cg)
                b = {k: torch.empty(shape, device=eng.device) for k in ("z", "zd1", "zd2", "zdd1", "zdd2")}
                for k in ("y", "yd", "ydd"):  # inputs of the next stage: split-bf16 on the tensor-core engine
                    b[k] = torch.empty_like(x)
                bufs.append(b)
            hd = p.desc[-1]
            bufs.append({k: torch.empty(n, hd.h, hd.w, 4, device=eng.device) for k in ("z", "zd1", "zd2", "zdd1", "zdd2")})
            self._bufs[key] = bufs
        bufs = self._bufs[key]
        st = L.stream()
        embed = embed.detach().contiguous().float()
        if eng.use_tc:
            L.check(L.lib.nq_nchw_to_split(L.ptr(embed), p.x[0].data_ptr(), n, c0, h0, w0, p.desc[0].cin_p, st), "nq_nchw_to_split")
        else:
            L.check(L.lib.nq_nchw_to_nhwc(L.ptr(embed), L.ptr(p.x[0]), n, c0, h0, w0, p.desc[0].cin_p, st), "nq_nchw_to_nhwc")
        x, xd, xdd = p.x[0], None, None
        last = len(eng.stages) - 1
        for i in range(last + 1):
            b = bufs[i]
            self._conv(p, i, x, False, True, b["z"])
            self._conv(p, i, x, True, False, b["zd2"])
            zd1 = zdd1 = zdd2 = None
            if xd is not None:
                self._conv(p, i, xd, False, False, b["zd1"])
                self._conv(p, i, xd, True, False, b["zdd2"])
                zd1, zdd2 = b["zd1"], b["zdd2"]
            if xdd is not None:
                self._conv(p, i, xdd, False, False, b["zdd1"])
                zdd1 = b["zdd1"]
            if i < last:
                L.check(L.lib.nq_jet_act(L.ptr(b["z"]), L.ptr(zd1), L.ptr(b["zd2"]), L.ptr(zdd1), L.ptr(zdd2), b["z"].numel(),
                                         _ACT[eng.geoms[i].act], b["y"].data_ptr(), b["yd"].data_ptr(), b["ydd"].data_ptr(),
                                         1 if eng.use_tc else 0, st), "nq_jet_act")
                x, xd, xdd = b["y"], b["yd"], b["ydd"]
            else:
                tgt = target.detach().contiguous().float()
                L.check(L.lib.nq_jet_head(L.ptr(b["z"]), L.ptr(zd1), L.ptr(b["zd2"]), L.ptr(zdd1), L.ptr(zdd2), L.ptr(tgt), n,
                                          p.H, p.W, _HEAD[eng.geoms[last].act], self.acc.data_ptr(), st), "nq_jet_head")
            eng.launches += 1

    def value(self) -> float:
        return float(self.acc)


def omega(engine: DecoderEngine, vecs: Sequence[torch.Tensor], batches) -> float:
    """Omega over `batches` of (embedding, frames); one direction, all batches."""
    ev = OmegaEvaluator(engine)
    shape = None
    for embed, img in batches:
        if tuple(embed.shape) != shape:  # a ragged last batch gets its own plan
            shape = tuple(embed.shape)
            ev.set_direction(vecs, embed.shape[0], embed.shape[2], embed.shape[3])
        ev.add_batch(embed, img)
    return ev.value()


def omega_layers(engine: DecoderEngine, vecs: Sequence[torch.Tensor], batches) -> List[float]:
    """The per-layer terms of Omega the reference logs (bit_assign.py:194-200): [v_l^T (H v)_l for every layer l], whose
    sum is Omega.  H is symmetric, so v_l^T H v = (Q(v + v_l) - Q(v - v_l)) / 4 with Q(u) = u^T H u the quantity the jet
    kernels evaluate; v +- v_l doubles / zeroes layer l's perturbation.  2 * n_layers evaluations."""
    out = []
    for l in range(len(vecs)):
        plus = [v * 2.0 if i == l else v for i, v in enumerate(vecs)]
        minus = [torch.zeros_like(v) if i == l else v for i, v in enumerate(vecs)]
        out.append(0.25 * (omega(engine, plus, batches) - omega(engine, minus, batches)))
    return out


def fisher_diag(engine: DecoderEngine, vecs: Sequence[torch.Tensor], batches, per_layer: bool = False):
    """bit_assign.py:122-168, :205-214: gradients of the MSE-mean loss accumulated over the batches,
    then sum v^2 g^2 (per_layer: the list of per-layer sums the reference logs, :208-214)."""
    if engine.mode != "off":
        raise L.NqError("fisher_diag is defined on the full-precision decoder")
    total = None
    for embed, target in batches:
        n, _, H, W = target.shape
        engine.forward(embed, train=True, target=target, p_norm=2.0, mean_pixels=float(3 * n * H * W), want_img=False)
        flat = engine.backward()
        total = flat.clone() if total is None else total + flat
    engine._grad[0].copy_(total)
    _, views = engine._grad_buffers()
    g = [gw.contiguous() for gw, _ in views]
    v = [x.detach().contiguous().float() for x in vecs]
    per = L.multi_dot(v, g, 1)
    return [float(x) for x in per] if per_layer else float(per.sum())
