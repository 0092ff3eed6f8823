"""Binding between the reference-shaped module tree (HNeRV / NeRV, optionally wrapped by QuantModel) and
the decoder engine: the modules own the tensors under the reference's attribute names, the engine
executes.  Nothing is copied: QuantStage fields alias the Parameters' storage, so Adam updates done
by the kernels are visible through `module.weight_quantizer.alpha` and vice versa."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from .engine import DecoderEngine, QuantStage, StageGeom


def _act_name(m: nn.Module) -> str:
    if isinstance(m, nn.GELU):
        return "gelu"
    if isinstance(m, nn.Identity):
        return "none"
    raise NotImplementedError(f"decoder activation {type(m).__name__}: only 'gelu' has a fused epilogue "
                              "(every reference config uses dec_acts: gelu)")


def conv2d_nchw(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Stride-1 'same' convolution, NCHW in / NCHW out, on the exact-fp32 FFMA kernel (stand-alone
    QuantModule.forward; the fused decoder path does not come through here)."""
    n, cin, h, w = x.shape
    cout, cin_w, k, _ = weight.shape
    if cin != cin_w:
        raise L.NqError(f"input has {cin} channels, weight expects {cin_w}")
    cin_p, cg = (cin + 3) // 4 * 4, (cout + 3) // 4 * 4
    d = L.ConvDesc(n, h, w, cin, cin_p, k, cout, 1, 1, cout, cg, 0)
    dev = x.device
    st = L.stream()
    xin = torch.empty(n, h, w, cin_p, device=dev)
    L.check(L.lib.nq_nchw_to_nhwc(L.ptr(x.detach().contiguous().float()), L.ptr(xin), n, cin, h, w, cin_p, st), "nq_nchw_to_nhwc")
    wk = torch.zeros(d.kdim, d.nout_p, device=dev)
    bp = torch.zeros(d.nout_p, device=dev)
    b = bias if bias is not None else torch.zeros(cout, device=dev)
    L.check(L.lib.nq_pack_weight(C.byref(d), L.ptr(weight.float()), cin, L.ptr(b.float()), L.ptr(wk), None, L.ptr(bp), st),
            "nq_pack_weight")
    y = torch.empty(n, h, w, cg, device=dev)
    L.check(L.lib.nq_conv_fwd(C.byref(d), L.ptr(xin), L.ptr(wk), L.ptr(bp), None, L.ptr(y), st), "nq_conv_fwd")
    out = torch.empty(n, cout, h, w, device=dev)
    L.check(L.lib.nq_nhwc_to_nchw(L.ptr(y), L.ptr(out), n, cout, h, w, cg, st), "nq_nhwc_to_nchw")
    return out


class EmbedList(list):
    """`embed_list` of HNeRV.decode / NeRV.decode (HNeRV.py:50-61, NeRV.py:45-56): [input embedding, output of the stem,
    output of every block].  Only entry 0 is consumed anywhere in the reference (calibrate_network.py:104: the calibration
    inputs), so the feature maps -- which live in the engine's NHWC buffers -- are converted to NCHW tensors on first
    access beyond entry 0 (one more decode of the same embedding + one layout kernel per stage)."""

    def __init__(self, runner: "DecoderRunner", embed: torch.Tensor, unfold_stem: bool):
        super().__init__([embed])
        self._runner, self._unfold, self._full = runner, unfold_stem, False
        self._n = len(runner.layers)  # input + stem + blocks (the head's output is not listed)

    def _fill(self):
        if self._full:
            return
        self._full = True
        r = self._runner
        feats = r.features(list.__getitem__(self, 0))[: self._n - 1]
        if self._unfold and (r.geoms[0].rh, r.geoms[0].rw) != (1, 1):  # HNeRV lists the stem output before the fold
            g = r.geoms[0]
            n, c, hh, ww = feats[0].shape
            feats[0] = feats[0].view(n, c, hh // g.rh, g.rh, ww // g.rw, g.rw).permute(0, 1, 3, 5, 2, 4).reshape(
                n, c * g.rh * g.rw, hh // g.rh, ww // g.rw)
        list.extend(self, feats)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if not (isinstance(i, int) and (i == 0 or i == -self._n)):
            self._fill()
        return list.__getitem__(self, i)

    def __iter__(self):
        self._fill()
        return list.__iter__(self)

    def __reduce__(self):
        self._fill()
        return (list, (list(list.__iter__(self)),))


class DecoderRunner:
    """One per model instance (cached in the model's __dict__, never pickled)."""

    KEY = "_nq_runner"

    @classmethod
    def of(cls, model: nn.Module) -> "DecoderRunner":
        r = model.__dict__.get(cls.KEY)
        if r is None or not r.matches(model):
            r = cls(model)
            model.__dict__[cls.KEY] = r
        return r

    def __init__(self, model: nn.Module):
        from .quantization.quant_block import QuantNeRVBlock
        from .quantization.quant_layer import QuantModule
        from .models._layers import NeRVBlock

        self.model = model
        layers, geoms = [], []
        stem = model.decoder[0]
        conv0 = stem if not isinstance(stem, QuantModule) else stem
        w0 = conv0.weight
        geoms.append(StageGeom(w0.shape[1], w0.shape[0], w0.shape[2], int(model.fc_h), int(model.fc_w), "none"))
        layers.append(stem)
        for blk in list(model.decoder)[1:]:
            if isinstance(blk, QuantNeRVBlock):
                conv, shuffle, act = blk.conv, blk.pixelshuffle, blk.act
            elif isinstance(blk, NeRVBlock):
                if not isinstance(blk.norm, nn.Identity):
                    raise NotImplementedError("dec_norm other than 'none' has no fused epilogue")
                conv, shuffle, act = blk.conv[0], blk.conv[1], blk.act
            else:
                raise NotImplementedError(f"unexpected decoder block {type(blk).__name__}")
            r = shuffle.upscale_factor if isinstance(shuffle, nn.PixelShuffle) else 1
            w = conv.weight
            geoms.append(StageGeom(w.shape[1], w.shape[0], w.shape[2], r, r, _act_name(act)))
            layers.append(conv)
        wh = model.head_layer.weight
        if str(model.out_bias) not in ("tanh", "sigmoid"):
            raise NotImplementedError(f"out_bias={model.out_bias!r}: fused head supports tanh / sigmoid")
        geoms.append(StageGeom(wh.shape[1], wh.shape[0], wh.shape[2], 1, 1, str(model.out_bias)))
        layers.append(model.head_layer)
        self.layers, self.geoms = layers, geoms
        self._layer_ids = [id(l) for l in layers]
        self._quant = [isinstance(l, QuantModule) for l in layers]
        hadamard = [bool(getattr(l, "hadamard", False)) for l in layers]
        stages = []
        for l, g, q, had in zip(layers, geoms, self._quant, hadamard):
            if l.bias is None:
                raise NotImplementedError("decoder convolutions without bias")
            st = QuantStage(g, l.weight.detach(), l.bias.detach(), 8, had and q)
            if q and had:
                st.w_src = l.hadamard_weight  # the module's own rotated copy (quant_layer.py:49)
                st.codes_w = torch.empty_like(st.w_src)
            stages.append(st)
        self.engine = DecoderEngine(stages)
        self._key = None

    def matches(self, model) -> bool:
        cur = [model.decoder[0]] + [getattr(b, "conv", None) if not isinstance(getattr(b, "conv", None), nn.Sequential)
                                    else b.conv[0] for b in list(model.decoder)[1:]] + [model.head_layer]
        return [id(l) for l in cur] == self._layer_ids

    # ------------------------------------------------------------------ module state -> engine state
    def sync(self):
        """Module state -> engine state.  The packed-weight cache is keyed on (data_ptr, version counter) of every tensor
        the engine aliases; the aliases are `.detach()` views, which SHARE the owning Parameter's version counter (a
        `.data` alias would get a fresh one that never moves), so `load_state_dict` / in-place updates by torch invalidate
        the cache.  Kernels of this library that update tensors in place (Adam) call `engine.invalidate()` themselves."""
        from .quantization.quantizer import AdaRoundQuantizer

        eng = self.engine
        on = [q and l.use_weight_quant for l, q in zip(self.layers, self._quant)]
        key = [tuple(on)]
        if not any(on):
            eng.mode = "off"
            for l, st in zip(self.layers, eng.stages):
                src = l.org_weight if hasattr(l, "org_weight") else l.weight.detach()
                srb = l.org_bias if hasattr(l, "org_bias") else l.bias.detach()
                st.weight, st.bias = src, srb
                key += [src.data_ptr(), src._version, srb.data_ptr(), srb._version]
        else:
            ada = [o and isinstance(l.weight_quantizer, AdaRoundQuantizer) for l, o in zip(self.layers, on)]
            eng.mode = "ada" if all(ada) else "uaq"
            eng.stage_state = None
            if all(ada):
                eng.soft_w = bool(self.layers[0].weight_quantizer.soft_targets)
                eng.soft_b = bool(self.layers[0].bias_quantizer.soft_targets)
            if any(ada) or not all(on):
                # per-stage rounding state: layers calibrated block by block carry AdaRound quantisers next to plain ones,
                # the block-wise variant leaves both quantisers of a block hard (calib_block.py:180-183), and
                # quantize_model_till (data_utils.py:261-272) quantises a prefix of the decoder only
                eng.stage_state = [("off", False, False) if not o else
                                   ("ada", bool(l.weight_quantizer.soft_targets), bool(l.bias_quantizer.soft_targets)) if a
                                   else ("uaq", False, False) for l, a, o in zip(self.layers, ada, on)]
                if all(ada) and len(set(eng.stage_state)) == 1:
                    eng.stage_state = None
            for l, st, o in zip(self.layers, eng.stages, on):
                if not o:  # full-precision stage inside a partly quantised decoder
                    src = l.org_weight if hasattr(l, "org_weight") else l.weight.detach()
                    srb = l.org_bias if hasattr(l, "org_bias") else l.bias.detach()
                    st.weight, st.bias = src, srb
                    key += [src.data_ptr(), src._version, srb.data_ptr(), srb._version]
                    continue
                wq, bq = l.weight_quantizer, l.bias_quantizer
                st.weight, st.bias = l.weight.detach(), l.bias.detach()
                if not st.hadamard:
                    st.w_src = st.weight
                st.set_bits(wq.n_bits)
                if not wq.inited if hasattr(wq, "inited") else False:
                    # first quantised forward: UniformAffineQuantizer.init_quantization_scale (quantizer.py:112-115)
                    # 'max' | 'mse' | 'l1' | 'gaussian' (--init); per channel or per tensor (--channel_wise)
                    d, z = wq.init_quantization_scale(st.w_src, wq.channel_wise)
                    wq.delta, wq.zero_point, wq.inited = nn.Parameter(d), z, True
                if not bq.inited if hasattr(bq, "inited") else False:
                    d, z = bq.init_quantization_scale(st.bias, bq.channel_wise)
                    bq.delta, bq.zero_point, bq.inited = nn.Parameter(d), z, True
                st.delta_w, st.zp_w = wq.delta.detach(), wq.zero_point
                st.delta_b, st.zp_b = bq.delta.detach(), bq.zero_point
                is_ada = isinstance(wq, AdaRoundQuantizer)
                st.alpha_w = wq.alpha.detach() if is_ada else None
                st.alpha_b = bq.alpha.detach() if is_ada else None
                for tns in (st.w_src, st.bias, st.delta_w, st.zp_w, st.delta_b, st.zp_b, st.alpha_w, st.alpha_b):
                    key += [None] if tns is None else [tns.data_ptr(), tns._version]
                key += [wq.n_bits, eng.soft_w, eng.soft_b, is_ada, getattr(wq, "soft_targets", None), getattr(bq, "soft_targets", None)]
        if key != self._key:
            eng.invalidate()
            self._key = key

    def publish_codes(self):
        """quantizer.py:297: expose the integer codes of the last forward as `x_quant`."""
        if self.engine.mode == "off":
            return
        for l, st in zip(self.layers, self.engine.stages):
            l.weight_quantizer.x_quant = st.codes_w
            l.bias_quantizer.x_quant = st.codes_b

    def decode(self, embed: torch.Tensor) -> torch.Tensor:
        if not embed.is_cuda:
            raise L.NqError("decode needs CUDA tensors: neuroquant_b200 has no CPU path")
        self.sync()
        img = self.engine.forward(embed, reuse_weights=True)
        self.publish_codes()
        return img.clone()

    def features(self, embed: torch.Tensor) -> List[torch.Tensor]:
        """NCHW copies of every stage output of the last decode (the reference's embed_list)."""
        self.decode(embed)
        p = self.engine._last_plan
        out = []
        for x, d in zip(p.x[1:], p.desc):
            n, h, w, cp = x.shape[-4:]
            c = d.c_grp
            t = torch.empty(n, c, h, w, device=x.device)
            if self.engine.use_tc:  # split-bf16 planes -> fp32 NCHW
                L.check(L.lib.nq_split_to_nchw(x.data_ptr(), L.ptr(t), n, c, h, w, cp, L.stream()), "nq_split_to_nchw")
            else:
                L.check(L.lib.nq_nhwc_to_nchw(L.ptr(x), L.ptr(t), n, c, h, w, cp, L.stream()), "nq_nhwc_to_nchw")
            out.append(t)
        return out
