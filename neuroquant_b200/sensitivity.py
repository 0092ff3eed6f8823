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


# ------------------------------------------------------------------------------------------------------------------
# bit_assign as a search: the Gram table of Omega and its exhaustive evaluation (BASELINE.json configs[3])
# ------------------------------------------------------------------------------------------------------------------
class OmegaTable:
    """Omega of EVERY per-layer bit-width configuration from one table.

    The reference scores a candidate by one Hessian-vector product over 10 mini-batches (bit_assign.py:57-118, ~15.6 s on
    its GPU) and therefore ships two hand-written candidates per architecture (:28-36).  Omega(v) = v^T H v is quadratic:
    with v_l(b) = W_l - Q_b(W_l), the perturbation of layer l alone at bit-width b,

        Omega(b_0 .. b_{L-1}) = sum_l G[(l,b_l),(l,b_l)] + 2 sum_{l<m} G[(l,b_l),(m,b_m)],
        G[(l,b),(m,b')] = v_l(b)^T H_lm v_m(b')

    so L*nb single-layer directions and C(L,2)*nb^2 two-layer directions, each ONE forward jet (OmegaEvaluator: no
    backward pass), give the whole table by polarisation, G[a,b] = (Q(a+b) - Q(a) - Q(b)) / 2 -- 49 + 1029 jets for 7
    layers x {2..8} bits -- after which all nb^L = 823 543 configurations are scored by table lookup on the device
    (nq_omega_search).  The jets are independent: under torch.distributed rank r evaluates directions r, r + world, ...
    and one all-reduce of the result vector assembles the table on every rank (candidate farming)."""

    def __init__(self, engine: DecoderEngine, perturbations, options: Sequence[int], n_params: Sequence[int]):
        """perturbations[l][k]: (C_out, C_in, k, k) perturbation of stage l at options[k] bits; n_params[l]: parameters
        of stage l counted in the average bit-width (weights + bias, quant_model.py:58-72)."""
        self.eng, self.pert, self.options = engine, perturbations, list(options)
        self.L, self.nb = len(perturbations), len(options)
        if self.L > 8 or self.nb > 16:
            raise L.NqError("OmegaTable: at most 8 layers x 16 options")
        self.n_params = [float(x) for x in n_params]
        self.gram = None
        self.jets = 0

    def directions(self):
        """The jets to evaluate: ('s', l, k) single-layer and ('p', l, k, m, k2) two-layer directions, in a fixed order."""
        d = [("s", l, k) for l in range(self.L) for k in range(self.nb)]
        d += [("p", l, k, m, k2) for l in range(self.L) for m in range(l + 1, self.L) for k in range(self.nb) for k2 in range(self.nb)]
        return d

    def _vecs(self, d):
        zero = [torch.zeros_like(self.pert[l][0]) for l in range(self.L)]
        zero[d[1]] = self.pert[d[1]][d[2]]
        if d[0] == "p":
            zero[d[3]] = self.pert[d[3]][d[4]]
        return zero

    def build(self, batches, rank: int = 0, world: int = 1, group=None):
        dirs = self.directions()
        q = torch.zeros(len(dirs), dtype=torch.float64, device=self.eng.device)
        for i in range(rank, len(dirs), world):
            q[i] = omega(self.eng, self._vecs(dirs[i]), batches)
            self.jets += 1
        if world > 1:
            torch.distributed.all_reduce(q, group=group)
        q = q.cpu()
        n = self.L * self.nb
        g = torch.zeros(n, n, dtype=torch.float64)
        single = {}
        for i, d in enumerate(dirs):
            if d[0] == "s":
                a = d[1] * self.nb + d[2]
                g[a, a] = q[i]
                single[a] = float(q[i])
        for i, d in enumerate(dirs):
            if d[0] == "p":
                a, b = d[1] * self.nb + d[2], d[3] * self.nb + d[4]
                g[a, b] = g[b, a] = 0.5 * (float(q[i]) - single[a] - single[b])
        self.gram = g
        return g

    def score(self, bits: Sequence[int]) -> float:
        """Omega of one configuration from the table (host arithmetic; the search kernel does the same sums)."""
        c = [self.options.index(b) for b in bits]
        s = 0.0
        for l in range(self.L):
            a = l * self.nb + c[l]
            s += float(self.gram[a, a])
            for m in range(l + 1, self.L):
                s += 2.0 * float(self.gram[a, m * self.nb + c[m]])
        return s

    def bits_weight(self) -> torch.Tensor:
        tot = sum(self.n_params)
        return torch.tensor([self.n_params[l] * b / tot for l in range(self.L) for b in self.options], dtype=torch.float64)

    def search(self, avg_bits_budget: float, want_scores: bool = False):
        """(best bits, its Omega, its average bit-width[, all scores]) over all nb^L configurations whose average
        bit-width does not exceed the budget."""
        dev = self.eng.device
        g = self.gram.to(dev).contiguous()
        bw = self.bits_weight().to(dev)
        n_cfg, ws_bytes = C.c_int64(0), C.c_int64(0)
        L.check(L.lib.nq_omega_search_workspace(self.L, self.nb, C.byref(n_cfg), C.byref(ws_bytes)), "nq_omega_search_workspace")
        ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
        scores = torch.empty(n_cfg.value, dtype=torch.float64, device=dev) if want_scores else None
        best = torch.zeros(1, dtype=torch.float64, device=dev)
        idx = torch.zeros(1, dtype=torch.int64, device=dev)
        L.check(L.lib.nq_omega_search(g.data_ptr(), bw.data_ptr(), self.L, self.nb, float(avg_bits_budget),
                                      scores.data_ptr() if want_scores else None, ws.data_ptr(), ws.numel(), best.data_ptr(),
                                      idx.data_ptr(), L.stream()), "nq_omega_search")
        self.eng.launches += 2
        i = int(idx)
        if i < 0:
            return None, float("inf"), float("nan"), scores
        bits, ab = [], 0.0
        for l in range(self.L):
            k = i % self.nb
            i //= self.nb
            bits.append(self.options[k])
            ab += float(bw[l * self.nb + k])
        return bits, float(best), ab, scores
