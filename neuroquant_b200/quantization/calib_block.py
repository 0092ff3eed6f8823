"""Block-wise calibration (reference: quantization/calib_block.py:91-183, data_utils.py:45-86,146-196).

One decoder block (conv -> PixelShuffle -> GELU, one engine stage) learns its rounding variables against the block's
own full-precision output.  Per iteration, on the GPU: mini-batch gather + QDrop mixing from the HBM-resident caches
(nq_qdrop_gather), fake-quant of the block's weight and bias (one launch), operand
pack (one launch), tcgen05 forward with the activation derivative saved, ONE fused kernel for lp_loss + activation
backward + un-shuffle (nq_block_loss_bwd), tcgen05 weight gradient + finish, and the fused quantiser Jacobian + Adam
(nq_adaround_step_multi).  No data gradient is needed: nothing upstream of the block learns.

Differences from the network-wise variant that the reference has and this keeps: only this block's quantisers become
AdaRound quantisers; the regulariser covers the block's weight only; BOTH its weight and bias quantisers end
hard-rounded (calib_block.py:180-183); the cached inputs / outputs drop the last len(cali) % 10 samples
(data_utils.py:67).  `opt_mode` 'fisher_diag' / 'fisher_full' (calib_block.py:66-72) weight the loss with the cached
output gradients of save_grad_data (data_utils.py:91-119): taken here with the engine's own backward, kept in HBM in the
target cache's layout, applied inside the fused loss kernel (nq_block_loss_bwd_fisher).  Hadamard blocks are not
supported (the reference's own block_reconstruction cannot run on a rotated layer: calib_block.py:125 initialises alpha
from the unrotated weight).
"""
from __future__ import annotations

import copy
import ctypes as C
import logging
import os

import torch
import torch.nn as nn

from .. import _lib as L
from ..calibration import LinearTempDecay
from ..engine import AdamState, ROUND_SOFT, StageGeom, _ACT, _ACT_SAVED_GRAD, _pad
from ..runner import DecoderRunner
from .data_utils import quantize_model_till  # noqa: F401  (data_utils.py:261-272)
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quantizer import AdaRoundQuantizer


class BlockStep:
    """Buffers, plans and the launch sequence of one block iteration for a fixed (batch, input grid)."""

    def __init__(self, stage, n: int, h: int, w: int, lr: float):
        g = stage.geom
        dev = stage.weight.device
        self.stage, self.n, self.h, self.w = stage, n, h, w
        cin_p = _pad(g.cin, 16)
        # nothing consumes the block's output as a K operand here, so its channel group only needs 16-byte stores (8)
        # as long as the GEMM's N = sub-pixels x group stays a multiple of 16 (HNeRV-3M block 5: 37 -> 40, N = 160)
        cg = _pad(g.c_grp, 8)
        if (g.rh * g.rw * cg) % 16:
            cg = _pad(g.c_grp, 16)
        act = _ACT[g.act]
        self.d = L.ConvDesc(n, h, w, g.cin, cin_p, g.k, g.cout, g.rh, g.rw, g.c_grp, cg, _ACT_SAVED_GRAD if act == 1 else act)
        self.H, self.W, self.cg, self.cin_p = h * g.rh, w * g.rw, cg, cin_p
        self.fwd = L.TcPlan()
        L.check(L.lib.nq_tc_plan_conv(C.byref(self.d), 0, 2, 2, C.byref(self.fwd)), "nq_tc_plan_conv")
        self.wg = L.TcWgradPlan()
        L.check(L.lib.nq_tc_plan_wgrad(C.byref(self.d), 2, 2, C.byref(self.wg)), "nq_tc_plan_wgrad")
        bf = dict(device=dev, dtype=torch.bfloat16)
        self.x = torch.zeros(2, n, h, w, cin_p, **bf)
        self.y = torch.zeros(2, n, self.H, self.W, cg, **bf)
        self.z = torch.zeros(n, self.H, self.W, cg, device=dev) if act != 0 else None
        self.tgt = torch.zeros(n, self.H, self.W, cg, device=dev)
        self.dz = torch.zeros(2, n, h, w, self.d.nout_p, **bf)
        self.wpk = torch.empty(int(self.fwd.wpk_bytes), device=dev, dtype=torch.uint8)
        self.scale = torch.empty(self.d.nout_p, device=dev)
        self.bias_p = torch.empty(self.d.nout_p, device=dev)
        self.deq_w, self.deq_b = torch.empty_like(stage.weight), torch.empty_like(stage.bias)
        self.ws = torch.empty(int(self.wg.workspace_floats), device=dev)
        self.gw, self.gb = torch.zeros_like(stage.weight), torch.zeros_like(stage.bias)
        self.loss = torch.zeros(1, device=dev)
        self.frame_dot = torch.zeros(n, device=dev)  # 'fisher_full': per-frame sum |d| F
        self.opt_mode = "mse"
        self.reg = torch.zeros(1, device=dev)
        self.hyper = torch.zeros(4, device=dev)
        # the host runs ahead of the device: pinned staging ring for the per-iteration scalars (as GraphedStep)
        self.hyper_host = [torch.zeros(4).pin_memory() for _ in range(8)]
        self.hyper_done = [None] * 8
        self.n_run = 0
        self._graphs = {}
        self.opt = AdamState([stage.alpha_w, stage.alpha_b], lr=lr)
        self.launches = 0

    def run(self, x_nchw: torch.Tensor, tgt_nchw: torch.Tensor, reg_w: float, reg_b: float, p: float, want_reg: bool = False):
        """One iteration from NCHW fp32 inputs (converted into the engine's layouts here)."""
        g, st = self.stage.geom, L.stream()
        L.check(L.lib.nq_nchw_to_split(L.ptr(x_nchw), self.x.data_ptr(), self.n, g.cin, self.h, self.w, self.cin_p, st), "nq_nchw_to_split")
        L.check(L.lib.nq_nchw_to_nhwc(L.ptr(tgt_nchw), L.ptr(self.tgt), self.n, g.c_grp, self.H, self.W, self.cg, st), "nq_nchw_to_nhwc")
        self.launches += 2
        self.run_cached(self.x, self.tgt, None, reg_w, reg_b, p, want_reg)

    def run_cached(self, x_split: torch.Tensor, tgt_cache: torch.Tensor, frame_idx, reg_w: float, reg_b: float, p: float,
                   want_reg: bool = False, opt_mode: str = "mse", fisher_cache: torch.Tensor = None, graph: bool = False):
        """One iteration from inputs already in the engine's layouts: x_split (2, n, h, w, cin_p) bf16 planes; tgt_cache
        (N, H, W, cg) fp32 NHWC with frame_idx (int32 device tensor of n entries) selecting the batch's frames, or a
        (n, H, W, cg) batch with frame_idx None.  graph=True: the launch sequence is captured once per set of buffers as
        a CUDA graph and replayed (the caller keeps x_split / frame_idx in the same buffers); the per-iteration scalars
        travel through the 16-byte device array either way."""
        self._set_hyper(reg_w, reg_b)
        if not graph or want_reg:
            self._body(x_split, tgt_cache, frame_idx, reg_b, p, want_reg, opt_mode, fisher_cache)
            return
        key = (x_split.data_ptr(), tgt_cache.data_ptr(), frame_idx.data_ptr() if frame_idx is not None else 0, float(p), opt_mode,
               fisher_cache.data_ptr() if fisher_cache is not None else 0)
        g = self._graphs.get(key)
        if g is None:
            self._body(x_split, tgt_cache, frame_idx, reg_b, p, False, opt_mode, fisher_cache)  # this iteration's real work
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            state = self.opt.params + self.opt.m + self.opt.v
            saved = [t.clone() for t in state]  # capture must not advance the state
            n0 = self.launches
            with torch.cuda.graph(g):
                self._body(x_split, tgt_cache, frame_idx, reg_b, p, False, opt_mode, fisher_cache)
            self.launches = n0
            for t, sv in zip(state, saved):
                t.copy_(sv)
            self._graphs[key] = g
        else:
            g.replay()
            self.launches += 8

    def _set_hyper(self, reg_w: float, reg_b: float):
        """(reg_w, reg_b, Adam step size, sqrt bias correction) of the step about to run -> the device array."""
        step_size, bc2 = self.opt.hyper_of_next_step()
        kk = self.n_run % len(self.hyper_host)
        self.n_run += 1
        if self.hyper_done[kk] is not None:
            self.hyper_done[kk].synchronize()
        hh = self.hyper_host[kk]
        hh[0], hh[1], hh[2], hh[3] = reg_w, reg_b, step_size, bc2
        self.hyper.copy_(hh, non_blocking=True)
        if self.hyper_done[kk] is None:
            self.hyper_done[kk] = torch.cuda.Event()
        self.hyper_done[kk].record()

    def _body(self, x_split, tgt_cache, frame_idx, reg_b, p, want_reg, opt_mode, fisher_cache):
        s, d, st = self.stage, self.d, L.stream()
        g = s.geom
        cw = lambda dl: int(dl.numel() > 1)  # noqa: E731
        rows = lambda x, dl: (dl.numel(), x.numel() // dl.numel()) if dl.numel() > 1 else (1, x.numel())  # noqa: E731
        # 1. soft fake-quant of weight and bias (quantizer.py:278-300)
        rw_, rl_ = rows(s.w_src, s.delta_w)
        rb_, bl_ = rows(s.bias, s.delta_b)
        fq = (L.FqTask * 2)(
            L.FqTask(L.ptr(s.w_src), L.ptr(s.alpha_w), L.ptr(s.delta_w), L.ptr(s.zp_w), L.ptr(s.codes_w), L.ptr(self.deq_w), rw_, rl_,
                     cw(s.delta_w), s.n_bits, ROUND_SOFT, int(want_reg)),
            L.FqTask(L.ptr(s.bias), L.ptr(s.alpha_b), L.ptr(s.delta_b), L.ptr(s.zp_b), L.ptr(s.codes_b), L.ptr(self.deq_b), rb_, bl_,
                     cw(s.delta_b), s.n_bits, ROUND_SOFT, 0))
        if want_reg:
            self.reg.zero_()
        L.check(L.lib.nq_fakequant_fwd_multi(fq, 2, L.ptr(self.reg) if want_reg else None, float(reg_b), st), "nq_fakequant_fwd_multi")
        # 2. operand pack + epilogue vectors
        pk = (L.TcPackTask * 1)(L.TcPackTask(C.pointer(d), C.pointer(self.fwd), L.ptr(self.deq_w), None, self.wpk.data_ptr(), None,
                                             L.ptr(self.deq_b), L.ptr(self.scale), L.ptr(self.bias_p), g.cin, 0, 0, 0))
        L.check(L.lib.nq_tc_pack_multi(pk, 1, st), "nq_tc_pack_multi")
        # 3. forward (activation derivative kept in z)
        L.check(L.lib.nq_tc_conv_fwd(C.byref(d), C.byref(self.fwd), x_split.data_ptr(), self.wpk.data_ptr(), L.ptr(self.scale),
                                     L.ptr(self.bias_p), L.ptr(self.z), self.y.data_ptr(), st), "nq_tc_conv_fwd")
        # 4. loss + backward through activation / up-shuffle: mean over n*H*W of sum_c |y - tgt|^p (quantizer.py:66-73)
        self.loss.zero_()
        self.opt_mode = opt_mode
        fi = frame_idx.data_ptr() if frame_idx is not None else None
        if opt_mode == "mse":
            L.check(L.lib.nq_block_loss_bwd(self.y.data_ptr(), L.ptr(tgt_cache), fi, L.ptr(self.z), self.n, self.h, self.w, g.rh, g.rw,
                                            self.cg, float(p), 1.0 / float(self.n * self.H * self.W), L.ptr(self.loss),
                                            self.dz.data_ptr(), st), "nq_block_loss_bwd")
        elif opt_mode in ("fisher_diag", "fisher_full"):
            # calib_block.py:66-72: sum_c d^2 F^2 averaged over n*H*W, or mean over n*C*H*W of (sum |d| F) |d| F / 100
            if fisher_cache is None or fisher_cache.shape[1:] != tgt_cache.shape[1:]:
                raise L.NqError("Fisher block loss needs the cached output gradients in the target cache's layout")
            diag = opt_mode == "fisher_diag"
            scale = 1.0 / float(self.n * self.H * self.W) if diag else 1.0 / (100.0 * self.n * g.c_grp * self.H * self.W)
            L.check(L.lib.nq_block_loss_bwd_fisher(self.y.data_ptr(), L.ptr(tgt_cache), L.ptr(fisher_cache), fi, L.ptr(self.z), self.n,
                                                   self.h, self.w, g.rh, g.rw, self.cg, 1 if diag else 2, scale, L.ptr(self.loss),
                                                   L.ptr(self.frame_dot), self.dz.data_ptr(), st), "nq_block_loss_bwd_fisher")
            self.launches += 0 if diag else 2
        else:
            raise ValueError('Not supported reconstruction loss function: {}'.format(opt_mode))  # calib_block.py:73-74
        # 5. weight / bias gradient
        L.check(L.lib.nq_tc_conv_wgrad(C.byref(d), C.byref(self.wg), x_split.data_ptr(), self.dz.data_ptr(), None, L.ptr(self.ws),
                                       self.ws.numel(), st), "nq_tc_conv_wgrad")
        fin = (L.WgFinishTask * 1)(L.WgFinishTask(C.pointer(d), L.ptr(self.ws), L.ptr(self.gw), L.ptr(self.gb), self.wg.psplits,
                                                  self.wg.N, g.cin, 0))
        L.check(L.lib.nq_tc_wgrad_finish_multi(fin, 1, st), "nq_tc_wgrad_finish_multi")
        # 6. quantiser Jacobian + Adam on (alpha_w, alpha_b); the regulariser acts on the weight only (calib_block.py:38-47)
        ad = (L.AdaTask * 2)(
            L.AdaTask(L.ptr(self.gw), L.ptr(s.w_src), L.ptr(s.alpha_w), L.ptr(s.delta_w), L.ptr(s.zp_w), L.ptr(self.opt.m[0]),
                      L.ptr(self.opt.v[0]), rw_, rl_, cw(s.delta_w), s.n_bits, 1, 0),
            L.AdaTask(L.ptr(self.gb), L.ptr(s.bias), L.ptr(s.alpha_b), L.ptr(s.delta_b), L.ptr(s.zp_b), L.ptr(self.opt.m[1]),
                      L.ptr(self.opt.v[1]), rb_, bl_, cw(s.delta_b), s.n_bits, 0, 0))
        L.check(L.lib.nq_adaround_step_multi(ad, 2, 1.0, 0.9, 0.999, 1e-8, L.ptr(self.hyper), st), "nq_adaround_step_multi")
        self.launches += 8

    def rec_loss(self) -> float:
        if self.opt_mode == "fisher_full":
            return float((self.frame_dot.double() ** 2).sum()) / (100.0 * self.n * self.stage.geom.c_grp * self.H * self.W)
        return float(self.loss) / float(self.n * self.H * self.W)


def cache_to_engine_layout(step: BlockStep, inps: torch.Tensor, syms, outs: torch.Tensor):
    """NCHW fp32 caches of N frames -> (2, N, h, w, cin_p) split-bf16 inputs and (N, H, W, cg) fp32 NHWC targets."""
    g, st = step.stage.geom, L.stream()
    N = inps.shape[0]
    dev = inps.device

    def split(t_):
        o = torch.zeros(2, N, step.h, step.w, step.cin_p, device=dev, dtype=torch.bfloat16)
        L.check(L.lib.nq_nchw_to_split(L.ptr(t_.contiguous()), o.data_ptr(), N, g.cin, step.h, step.w, step.cin_p, st), "nq_nchw_to_split")
        return o

    out_c = nhwc_cache(step, outs)
    return split(inps), (split(syms) if syms is not None else None), out_c


def nhwc_cache(step: BlockStep, t_: torch.Tensor) -> torch.Tensor:
    """(N, c_grp, H, W) fp32 -> (N, H, W, cg) fp32, pad channels zero: the layout of the target and Fisher caches."""
    g, N = step.stage.geom, t_.shape[0]
    o = torch.zeros(N, step.H, step.W, step.cg, device=t_.device)
    L.check(L.lib.nq_nchw_to_nhwc(L.ptr(t_.contiguous()), L.ptr(o), N, g.c_grp, step.H, step.W, step.cg, L.stream()), "nq_nchw_to_nhwc")
    return o


def assemble_batch(inp_s: torch.Tensor, sym_s, idx_dev: torch.Tensor, input_prob: float, out: torch.Tensor) -> int:
    """calib_block.py:160-164 in one launch: out[:, b] = inp_s[:, idx[b]], and with QDrop (input_prob < 1) each element is
    taken from inp_s where torch.rand_like(cur_inp) < input_prob, else from sym_s.  The uniform draw stays torch's (one
    call per iteration, as in the reference), in the batch's own (n, h, w, cin_p) layout."""
    n = out.shape[1]
    frame = out[0, 0].numel()
    rnd = None
    if input_prob < 1.0:
        rnd = torch.rand_like(out[0], dtype=torch.float32)
    L.check(L.lib.nq_qdrop_gather(inp_s.data_ptr(), sym_s.data_ptr() if rnd is not None else None, idx_dev.data_ptr(), L.ptr(rnd),
                                  float(input_prob), n, inp_s.shape[1], frame, out.data_ptr(), L.stream()), "nq_qdrop_gather")
    return 2 if rnd is not None else 1


def _features(runner: DecoderRunner, model, embed: torch.Tensor, weight_quant: bool):
    model.set_quant_state(weight_quant)
    return runner.features(embed)


def save_inp_oup_data(model, runner: DecoderRunner, k: int, cali_data: torch.Tensor, asym: bool, batch_size: int = 10,
                      layer=None):
    """data_utils.py:45-86 with input_prob=True: (block input the optimisation sees, full-precision block input,
    full-precision block output) over the calibration set, kept in HBM.  `layer` (a QuantModule): the hooked module is
    the convolution alone -- its output is taken before the up-shuffle and the activation (layer_reconstruction)."""
    inps, syms, outs = [], [], []
    for i in range(int(cali_data.size(0) / batch_size)):
        e = cali_data[i * batch_size:(i + 1) * batch_size].cuda()
        feats = _features(runner, model, e, False)
        syms.append(feats[k - 1].clone() if k > 0 else e.float().clone())
        if layer is not None:
            with torch.no_grad():
                outs.append(layer(syms[-1]))  # quant state is off: the module's own full-precision convolution
        else:
            outs.append(feats[k].clone())
        if asym and k > 0:  # input recomputed with the whole network quantised (data_utils.py:172-180)
            inps.append(_features(runner, model, e, True)[k - 1].clone())
        else:
            inps.append(syms[-1])
    model.set_quant_state(False)
    return (torch.cat(inps), torch.cat(syms)), torch.cat(outs)


def block_output_grads(model, runner: DecoderRunner, block, k: int, cali_data: torch.Tensor, layer: bool = False) -> torch.Tensor:
    """GetLayerGrad (data_utils.py:222-258) over the calibration set, one sample at a time: the gradient of
    mean((out_fp - out_q)^2) -- out_q with the decoder quantised up to and including the block -- w.r.t. the block's
    output in the QUANTISED pass (the call the reference's backward hook keeps; pinned in tests/golden/block_*_f*.npz).
    The engine runs the partly quantised decoder forward with out_fp as the target, back-propagates, and re-runs the
    data gradient of stage k+1 without the activation derivative."""
    eng = runner.engine
    out = []
    for i in range(cali_data.size(0)):
        e = cali_data[i:i + 1].cuda()
        model.set_quant_state(False)
        runner.sync()
        out_fp = eng.forward(e, reuse_weights=True).clone()
        quantize_model_till(model, block)
        runner.sync()
        _, _, hh, ww = out_fp.shape
        eng.forward(e, train=True, target=out_fp, p_norm=2.0, mean_pixels=float(out_fp[0].numel()), reuse_weights=True, want_img=False)
        eng.backward()
        out.append(eng.stage_output_grad(k) if layer else eng.stage_input_grad(k + 1))
    model.set_quant_state(False)
    block.set_quant_state(True)
    return torch.cat(out)


def save_grad_data(model, runner: DecoderRunner, block, k: int, cali_data: torch.Tensor, layer: bool = False) -> torch.Tensor:
    """data_utils.py:91-119 (batch_size=1): |g| + 1, kept in HBM."""
    return block_output_grads(model, runner, block, k, cali_data, layer).abs() + 1.0


def block_reconstruction(model, block: BaseQuantBlock, cali_data: torch.Tensor, batch_size: int = 8, iters: int = 20000,
                         weight: float = 0.01, opt_mode: str = "mse", asym: bool = False, b_range: tuple = (20, 2),
                         warmup: float = 0.0, input_prob: float = 1.0, p: float = 2.0, lr: float = 0.0015):
    """Block-wise calibration (calib_block.py:91-183); same arguments as the reference."""
    if not isinstance(block, BaseQuantBlock):
        raise ValueError("block_reconstruction expects a QuantNeRVBlock of the model's decoder")
    convs = [m for m in block.modules() if isinstance(m, QuantModule)]
    if len(convs) != 1:
        raise NotImplementedError("blocks with other than one quantised convolution")
    return _reconstruct(model, block, convs[0], False, cali_data, batch_size, iters, weight, opt_mode, asym, b_range, warmup,
                        input_prob, p, lr)


def _reconstruct(model, block, conv, layer_mode: bool, cali_data, batch_size, iters, weight, opt_mode, asym, b_range, warmup,
                 input_prob, p, lr):
    """Shared body of block_reconstruction and layer_reconstruction (calib_layer.py:89-179 is calib_block.py:91-183 on a
    lone QuantModule): layer_mode compares the convolution's own output and never applies the rounding regulariser."""
    if opt_mode not in ("mse", "fisher_diag", "fisher_full"):
        raise ValueError('Not supported reconstruction loss function: {}'.format(opt_mode))
    if conv.hadamard:
        raise NotImplementedError("block_reconstruction on a rotated layer (the reference cannot run it either: calib_block.py:125)")
    runner = DecoderRunner.of(model.model)
    k = [i for i, l in enumerate(runner.layers) if l is conv]
    if not k or (not layer_mode and (k[0] == 0 or k[0] == len(runner.layers) - 1)):
        raise ValueError("block is not one of this model's decoder blocks")
    k = k[0]
    # the caches depend on the predecessors only; take them before this block's quantisers are swapped
    model.eval()
    model.set_quant_state(True)
    runner.sync()  # initialises step sizes that no forward has initialised yet
    (cached_inps, cached_sym), cached_outs = save_inp_oup_data(model, runner, k, cali_data, asym, layer=conv if layer_mode else None)
    model.set_quant_state(False)
    block.set_quant_state(True)
    round_mode = "learned_hard_sigmoid"
    conv.weight_quantizer = AdaRoundQuantizer(uaq=conv.weight_quantizer, round_mode=round_mode, weight_tensor=conv.org_weight.data)
    conv.weight_quantizer.soft_targets = True
    conv.bias_quantizer = AdaRoundQuantizer(uaq=conv.bias_quantizer, round_mode=round_mode, weight_tensor=conv.bias.data)
    conv.bias_quantizer.soft_targets = True
    wq, bq = conv.weight_quantizer, conv.bias_quantizer
    wq.alpha, bq.alpha = nn.Parameter(wq.alpha.data.contiguous()), nn.Parameter(bq.alpha.data.contiguous())
    st = runner.engine.stages[k]
    st.weight, st.bias, st.w_src = conv.weight.data, conv.bias.data, conv.weight.data
    st.set_bits(wq.n_bits)
    st.delta_w, st.zp_w = wq.delta.data.contiguous(), wq.zero_point.contiguous()
    st.delta_b, st.zp_b = bq.delta.data.contiguous(), bq.zero_point.contiguous()
    st.alpha_w, st.alpha_b = wq.alpha.data, bq.alpha.data
    runner._key = None
    # calib_block.py:154-157: the output gradients are taken AFTER the block's quantisers were swapped (AdaRound, soft)
    cached_grads = save_grad_data(model, runner, block, k, cali_data, layer_mode) if opt_mode != "mse" else None
    n_cached = cached_inps.size(0)
    bsz = min(batch_size, n_cached)
    if layer_mode:
        # the layer alone: same quantiser state, but the compared output is the convolution's (no up-shuffle, no activation)
        st = copy.copy(st)
        st.geom = StageGeom(st.geom.cin, st.geom.cout, st.geom.k, 1, 1, "none")
    step = BlockStep(st, bsz, cached_inps.shape[2], cached_inps.shape[3], lr)
    decay = LinearTempDecay(iters, rel_start_decay=warmup + (1 - warmup) * 0.0, start_b=b_range[0], end_b=b_range[1])
    loss_start = iters * warmup
    model.train()
    # the caches go into the engine's own layouts ONCE (split-bf16 NHWC inputs, fp32 NHWC targets): an iteration then
    # only gathers its frames (contiguous per frame) and never transposes
    inp_s, sym_s, out_c = cache_to_engine_layout(step, cached_inps, cached_sym if input_prob < 1.0 else None, cached_outs)
    grad_c = nhwc_cache(step, cached_grads) if cached_grads is not None else None
    cur = torch.empty_like(inp_s[:, :bsz])
    idx = torch.zeros(bsz, dtype=torch.int32, device=cached_inps.device)
    use_graph = os.environ.get("NQ_GRAPH", "1") != "0"
    for i in range(iters):
        idx_h = torch.randperm(n_cached)[:batch_size]
        idx.copy_(idx_h)                       # static device buffer: the graphed step reads the same address every time
        # gather + QDrop (calib_block.py:160-164); one draw per element, shared by the hi and lo planes
        assemble_batch(inp_s, sym_s, idx, input_prob, cur)
        count = i + 1
        b = decay(count)
        # calib_layer.py:38-46: collect_round_loss walks the CHILDREN of the module it is given; a lone QuantModule has
        # none that is a QuantModule, so the layer-wise variant never sees the regulariser
        reg_on = not (count < loss_start) and not layer_mode
        want_log = count % 500 == 0
        step.run_cached(cur, out_c, idx, weight if reg_on else 0.0, float(b) if reg_on else 0.0, p,
                        want_reg=want_log and reg_on, opt_mode=opt_mode, fisher_cache=grad_c, graph=use_graph)
        if want_log:  # calib_block.py:85-87
            rec = step.rec_loss()
            rnd = float(step.reg) * weight if reg_on else 0.0
            logging.info('Total loss:\t{:.4f} (rec:{:.4f}, round:{:.4f})\tb={:.2f}\tcount={}'.format(
                rec + rnd, rec, rnd, b if reg_on else 0, count))
    torch.cuda.empty_cache()
    wq.soft_targets = False  # calib_block.py:180-183: both quantisers of the block go hard
    bq.soft_targets = False
    runner._key = None
    return step
