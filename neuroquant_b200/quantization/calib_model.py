"""Network-wise calibration (reference: quantization/calib_model.py): same signature, same two Adam
phases and bookkeeping, executed by the decoder engine (neuroquant_b200/calibration.py) instead of
PyTorch autograd -- forward, loss, backward, rounding regulariser and Adam are libnq_sm100 kernels."""
import logging

import torch
import torch.nn as nn

from ..calibration import CalibrationLoop, HostBatchPipe, LinearTempDecay
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


# facts about the last model_reconstruction() of this process (bench.py reports the host -> device bytes per step)
LAST_RUN = {}


class _Staged:
    """One mini-batch whose host -> device copy is in flight (HostBatchPipe slot)."""
    __slots__ = ("idx", "n_global", "resident", "pipe")

    def __init__(self, idx, n_global, resident, pipe):
        self.idx, self.n_global, self.resident, self.pipe = idx, n_global, resident, pipe


class _FrameSource:
    """Feeds the loop from the reference's `gt` loader (dicts with 'img' and 'idx', calib_model.py:147-151).

    Frames may arrive as fp32 in [0, 1] (the reference's `read_image / 255`) or as uint8 (VideoDataSet(as_uint8=True):
    a quarter of the bytes over PCIe and in HBM; value / 255 is evaluated inside the head kernel).  Batches coming from
    the loader go through a HostBatchPipe: the copy of batch k+1 runs on a side stream under the kernels of batch k.

    residency 'hbm'    frames are decoded by the loader ONCE (first epoch) and kept resident in HBM; later epochs only
                       draw index batches from `gt.batch_sampler`, so the PNG decode + host-to-device copy per
                       iteration of calib_model.py:150 leaves the loop.  Every rank copies the whole global batch.
              'stream' every iteration takes its frames from the loader (clips that do not fit in HBM); a rank copies
                       only its own shard of the global batch.
              'auto'   'hbm' when the clip takes less than a quarter of the free device memory."""

    def __init__(self, gt, cali_data: torch.Tensor, rank: int, world: int, residency: str = "auto", device=None):
        self.gt, self.rank, self.world = gt, rank, world
        self.dev = torch.device(device) if device is not None else (cali_data.device if cali_data.is_cuda else torch.device("cuda"))
        self.cali_host = None if cali_data.is_cuda else cali_data
        self.cali = cali_data if cali_data.is_cuda else None
        if residency not in ("auto", "hbm", "stream"):
            raise ValueError(f"residency {residency!r}: expected 'auto', 'hbm' or 'stream'")
        self.residency = residency
        self.frames = None
        self.have = None
        self.pipe = None  # the pipe staged into last
        self.pipes = {}
        self.h2d_bytes = 0  # bytes this rank copied host -> device (bench.py reports it per step)

    # -- host side ---------------------------------------------------------------------------------------------------
    def _decide(self, img: torch.Tensor):
        n_frames = (self.cali if self.cali is not None else self.cali_host).shape[0]
        if self.residency == "auto":
            free, _ = torch.cuda.mem_get_info(self.dev)
            self.residency = "hbm" if n_frames * img[0].numel() * img.element_size() < free // 4 else "stream"
        if self.residency == "hbm":
            if self.cali is None:
                self.cali = self.cali_host.to(self.dev)
            self.frames = torch.empty((n_frames,) + tuple(img.shape[1:]), device=self.dev, dtype=img.dtype)
            self.have = torch.zeros(n_frames, dtype=torch.bool)

    def _stage(self, sample) -> _Staged:
        img = sample["img"]
        idx = torch.as_tensor(sample["idx"]).view(-1).cpu()
        if img.dtype not in (torch.uint8, torch.float32):
            img = img.float()
        if self.frames is None and self.residency != "stream":
            self._decide(img)
        resident = self.residency == "hbm"
        mine = idx if resident else shard_indices(idx, self.rank, self.world)
        img_m = img if (resident or self.world == 1) else img[self.rank::self.world]
        tensors = [img_m]
        if self.cali is None:  # embeddings on the host as well: this batch's rows travel with the frames
            tensors.append(self.cali_host[mine].contiguous())
        # one pipe per batch shape: a loader with drop_last=False ends its epoch on a short batch, which is staged while
        # the batch before it is still pending in the pipe of the full shape (the staged batch remembers its pipe)
        specs = tuple((tuple(t.shape), t.dtype) for t in tensors)
        pipe = self.pipes.get(specs)
        if pipe is None:
            pipe = self.pipes[specs] = HostBatchPipe(specs, device=self.dev, depth=3)
        self.pipe = pipe
        pipe.put(*tensors)
        self.h2d_bytes += sum(t.numel() * t.element_size() for t in tensors if not t.is_cuda)
        return _Staged(idx, int(idx.numel()), resident, pipe)

    def batches(self):
        if self.have is not None and bool(self.have.all()):
            sampler = getattr(self.gt, "batch_sampler", None)
            if sampler is not None:
                for idx in sampler:
                    yield list(idx)
            else:  # a plain sequence of sample dicts: its index batches, without touching the frames again
                for sample in self.gt:
                    yield torch.as_tensor(sample["idx"]).view(-1).tolist()
            return
        it = iter(self.gt)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur = nxt
            try:
                nxt = self._stage(next(it))  # batch k+1 starts its copy before batch k is consumed
            except StopIteration:
                nxt = None
            yield cur

    # -- device side -------------------------------------------------------------------------------------------------
    def fetch(self, item):
        if isinstance(item, _Staged):
            got = item.pipe.get()
            img = got[0]
            if item.resident:
                idx_d = item.idx.to(self.dev)
                self.frames[idx_d] = img
                self.have[item.idx] = True
                mine = shard_indices(idx_d, self.rank, self.world)
                return self.cali[mine], self.frames[mine], item.n_global
            if self.cali is None:
                return got[1], img, item.n_global
            mine = shard_indices(item.idx, self.rank, self.world).to(self.dev)
            return self.cali[mine], img, item.n_global
        idx = torch.as_tensor(item, device=self.dev).view(-1)
        mine = shard_indices(idx, self.rank, self.world)
        return self.cali[mine], self.frames[mine], int(idx.numel())


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
                         b_range: tuple = (20, 2), warmup: float = 0.0, p: float = 2.0, lr: float = 0.0015, *,
                         frame_residency: str = "auto", on_iteration=None):
    """Network-wise calibration (calib_model.py:92-240).  `model` is a QuantModel; `cali_data` the decoder
    inputs of every frame; `gt` the frame loader (dicts with 'img' -- fp32 in [0, 1] or uint8 -- and 'idx').  Under
    torch.distributed every rank calls this with the same arguments: mini-batches are sharded by frame and the weight
    gradients are all-reduced (neuroquant_b200/calibration.py).  `gt` must then yield the SAME index batches on every rank
    (a shuffling loader needs an identically seeded generator, as methods/calibrate_network.py builds it; the reference
    itself is unseeded) and every global batch must split evenly over the ranks (parallel.shard_indices raises).

    Additive, keyword-only: `frame_residency` ('auto' | 'hbm' | 'stream', see _FrameSource) and
    `on_iteration(phase, count, loss)` -- called after every iteration with the reconstruction loss of that iteration as
    a DEVICE scalar (no synchronisation; this rank's share under data parallelism): progress reporting / loss read-back
    without waiting for the reference's one log line per 500 iterations."""
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
    src = _FrameSource(gt, cali_data, rank, world, residency=frame_residency, device=eng.device)
    LAST_RUN.clear()
    LAST_RUN["_src"] = src  # live during the run (bench.py samples its byte counter from the iteration callback)
    n_batches = len(gt)
    gb = getattr(gt, "batch_size", None) or batch_size
    loop = CalibrationLoop(eng, src.fetch, n_batches, iters, weight=weight, b_range=b_range, warmup=warmup, p=p, lr=lr,
                           group=group, global_batch=gb, on_iteration=on_iteration)
    model.train()
    loop.run_phase1(src.batches)
    torch.cuda.empty_cache()
    loop.run_phase2(src.batches, on_start=lambda: _install_adaround(runner, round_mode))
    torch.cuda.empty_cache()
    for l in runner.layers:  # calib_model.py:231-240: weight quantisers go hard, bias quantisers stay soft (Q3)
        l.weight_quantizer.soft_targets = False
    runner._key = None
    LAST_RUN.pop("_src", None)
    LAST_RUN.update(h2d_bytes=src.h2d_bytes, residency=src.residency, launches=eng.launches)
