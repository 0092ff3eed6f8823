"""Network-wise calibration (reference: quantization/calib_model.py): same signature, same two Adam
phases and bookkeeping, executed by the decoder engine (neuroquant_b200/calibration.py) instead of
PyTorch autograd -- forward, loss, backward, rounding regulariser and Adam are libnq_sm100 kernels."""
import logging

import torch
import torch.nn as nn

from ..calibration import CalibrationLoop, LinearTempDecay
from ..parallel import shard_indices, world_info
from ..runner import DecoderRunner
from .quant_layer import QuantModule
from .quantizer import AdaRoundQuantizer, lp_loss


class LossFunction:
    """calib_model.py:16-89.  API mirror for callers that drive their own loop; model_reconstruction
    below uses the fused engine step, which computes the same two terms inside the kernels."""

    def __init__(self, model: nn.Module, round_loss: str = "relaxation", weight: float = 1.0, rec_loss: str = "mse",
                 max_count: int = 2000, b_range: tuple = (10, 2), decay_start: float = 0.0, warmup: float = 0.0,
                 p: float = 2.0):
        self.model = model
        self.round = round_loss
        self.weight = weight
        self.rec = rec_loss
        self.loss_start = max_count * warmup
        self.p = p
        self.temp_decay = LinearTempDecay(max_count, rel_start_decay=warmup + (1 - warmup) * decay_start,
                                          start_b=b_range[0], end_b=b_range[1])
        self.count = 0

    def collect_round_loss(self, module, b):
        for name, m in module.named_children():
            if "encoder" in name:
                continue
            elif isinstance(m, QuantModule):
                h = m.weight_quantizer.get_soft_targets()
                self.round_loss += self.weight * (1 - ((h - 0.5).abs() * 2).pow(b)).sum()
            else:
                self.collect_round_loss(m, b)

    def __call__(self, pred, tgt, grad=None):
        self.count += 1
        if self.rec != "mse":
            raise NotImplementedError("fisher_diag / fisher_full belong to the block-wise variant (SURVEY 8(f))")
        rec_loss = lp_loss(pred, tgt, p=self.p)
        b = self.temp_decay(self.count)
        if self.count < self.loss_start or self.round == "none":
            b = self.round_loss = 0
        elif self.round == "relaxation":
            self.round_loss = 0
            self.collect_round_loss(self.model, b)
        else:
            raise NotImplementedError
        total = self.round_loss + rec_loss
        if self.count % 500 == 0:
            logging.info("Total loss:\t{:.4f} (rec:{:.4f}, round:{:.4f})\tb={:.2f}\tcount={}".format(
                float(total), float(rec_loss), float(self.round_loss), b, self.count))
        return total


class _FrameSource:
    """Feeds the loop from the reference's `gt` DataLoader.  Frames are decoded by the loader ONCE (first
    epoch) and kept resident in HBM; later epochs only draw index batches from `gt.batch_sampler`, so
    the PNG decode + 20 MB host-to-device copy per iteration of calib_model.py:150 leaves the loop."""

    def __init__(self, gt, cali_data: torch.Tensor, rank: int, world: int):
        self.gt, self.cali, self.rank, self.world = gt, cali_data, rank, world
        self.frames = None
        self.have = None

    def batches(self):
        if self.have is not None and bool(self.have.all()) and hasattr(self.gt, "batch_sampler") and \
                self.gt.batch_sampler is not None:
            for idx in self.gt.batch_sampler:
                yield list(idx)
        else:
            for sample in self.gt:
                yield sample

    def fetch(self, item):
        dev = self.cali.device
        if isinstance(item, dict):
            idx = torch.as_tensor(item["idx"]).view(-1).to(dev)
            img = item["img"].to(dev, non_blocking=True).float()
            if self.frames is None:
                self.frames = torch.empty((self.cali.shape[0],) + tuple(img.shape[1:]), device=dev)
                self.have = torch.zeros(self.cali.shape[0], dtype=torch.bool, device=dev)
            self.frames[idx] = img
            self.have[idx] = True
        else:
            idx = torch.as_tensor(item, device=dev).view(-1)
        mine = shard_indices(idx, self.rank, self.world)
        return self.cali[mine], self.frames[mine]


def _install_adaround(runner: DecoderRunner, round_mode: str):
    """calib_model.py:169-184: swap every quantiser for an AdaRoundQuantizer.  The engine has already
    computed the fp16-rounded scales and alpha (start_adaround); the module objects adopt those tensors."""
    for l, st in zip(runner.layers, runner.engine.stages):
        for name, x_src, alpha, delta, zp in (("weight_quantizer", st.w_src, st.alpha_w, st.delta_w, st.zp_w),
                                              ("bias_quantizer", st.bias, st.alpha_b, st.delta_b, st.zp_b)):
            uaq = getattr(l, name)
            q = AdaRoundQuantizer.__new__(AdaRoundQuantizer)
            nn.Module.__init__(q)
            q.n_bits, q.sym, q.n_levels = uaq.n_bits, uaq.sym, uaq.n_levels
            q.round_mode, q.soft_targets, q.x_quant = round_mode, True, None
            q.gamma, q.zeta, q.beta = -0.1, 1.1, 2 / 3
            q.zero_point = zp
            q.alpha = nn.Parameter(alpha)
            q.delta = nn.Parameter(delta)
            setattr(l, name, q)


def model_reconstruction(model, cali_data: torch.Tensor, gt, arch: str = "hnerv", batch_size: int = 8,
                         iters: int = 20000, weight: float = 0.01, opt_mode: str = "mse", hadamard: bool = True,
                         b_range: tuple = (20, 2), warmup: float = 0.0, p: float = 2.0, lr: float = 0.0015):
    """Network-wise calibration (calib_model.py:92-240).  `model` is a QuantModel; `cali_data` the decoder
    inputs of every frame; `gt` the frame loader (dicts with 'img' and 'idx').  Under torch.distributed
    every rank calls this with the same arguments: mini-batches are sharded by frame and the weight
    gradients are all-reduced (neuroquant_b200/calibration.py)."""
    if arch not in ("hnerv", "nerv"):
        raise ValueError
    if opt_mode != "mse":
        raise NotImplementedError("opt_mode other than 'mse' (the command line hard-codes it, calibrate_network.py:264)")
    model.set_quant_state(True)
    round_mode = "learned_hard_sigmoid"
    runner = DecoderRunner.of(model.model)
    runner.sync()  # initialises the step sizes if no quantised forward has run yet
    eng = runner.engine
    rank, world, group = world_info()
    src = _FrameSource(gt, cali_data, rank, world)
    n_batches = len(gt)
    gb = getattr(gt, "batch_size", None) or batch_size
    loop = CalibrationLoop(eng, src.fetch, n_batches, iters, weight=weight, b_range=b_range, warmup=warmup, p=p, lr=lr,
                           group=group, global_batch=gb)
    model.train()
    loop.run_phase1(src.batches)
    torch.cuda.empty_cache()
    loop.run_phase2(src.batches, on_start=lambda: _install_adaround(runner, round_mode))
    torch.cuda.empty_cache()
    for l in runner.layers:  # calib_model.py:231-240: weight quantisers go hard, bias quantisers stay soft (Q3)
        l.weight_quantizer.soft_targets = False
    runner._key = None
