"""Network-wise calibration loop on the decoder engine (reference: quantization/calib_model.py:92-240).

Two Adam phases over a fixed kernel sequence per iteration:
  phase 1  step sizes delta (UAQ, straight-through rounding), lr 1e-3        calib_model.py:120-165
  phase 2  rounding variables alpha (AdaRound soft targets), lr = `lr`,
           + b-annealed rounding regulariser after the warm-up               calib_model.py:169-226
then the weight quantisers switch to hard rounding (bias quantisers stay soft, :231-240).

Multi-GPU: frames of each global mini-batch are sharded over the ranks of `group`; every rank runs
the same iteration on its shard and the flat dW/db buffer is summed with ONE NCCL all-reduce before
the (replica-identical) quantiser Jacobian and Adam.  The loss is normalised by the GLOBAL pixel
count, so the sum of the per-rank gradients is the gradient of the reference's mean.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Iterable, List, Optional, Sequence

import torch

from .engine import AdamState, DecoderEngine, adaround_step_multi


class LinearTempDecay:
    """data_utils.py:24-41."""

    def __init__(self, t_max: int, rel_start_decay: float = 0.2, start_b: float = 10, end_b: float = 2):
        self.t_max = t_max
        self.start_decay = rel_start_decay * t_max
        self.start_b = start_b
        self.end_b = end_b

    def __call__(self, t):
        if t < self.start_decay:
            return self.start_b
        rel_t = (t - self.start_decay) / (self.t_max - self.start_decay)
        return self.end_b + (self.start_b - self.end_b) * max(0.0, (1 - rel_t))


class GraphedStep:
    """One calibration iteration (forward, loss, backward, quantiser Jacobian, Adam) captured ONCE as a CUDA graph
    and replayed: ~100 kernel launches become one graph launch.  Inputs are copied into static buffers; the four
    scalars that change per iteration go through a 16-byte device array (nq_*_dev kernels).  Both phases; the eager
    path stays for iterations that log.  Data parallel: the NCCL all-reduce of the flat dW buffer is captured INTO the
    graph between the backward kernels and the fused Jacobian + Adam launch (the communicator is created and warmed up
    by the eager first iteration; capture runs in thread-local error mode, so the NCCL watchdog thread's event queries
    do not invalidate it).  NQ_GRAPH_DP=0 falls back to the eager launch sequence around an eager all-reduce."""

    def __init__(self, eng: DecoderEngine, opt: AdamState, embed: torch.Tensor, frames: torch.Tensor, p_norm: float,
                 mean_pixels: float, group=None, world: int = 1, capture: bool = True, phase: str = "alpha"):
        self.eng, self.opt, self.p_norm, self.mean_pixels = eng, opt, p_norm, mean_pixels
        self.phase = phase  # 'alpha': AdaRound phase (one fused Jacobian + Adam launch); 'delta': step-size phase
        self.group, self.world = group, world
        self.capture = capture  # False: the same fused kernel sequence launched eagerly (data-parallel default)
        self.embed = torch.empty_like(embed)
        self.frames = torch.empty_like(frames)
        self.hyper = torch.zeros(4, device=embed.device)
        # the host runs ahead of the device: a ring of pinned staging buffers, each reused only after its copy is done
        self.hyper_host = [torch.zeros(4).pin_memory() for _ in range(8)]
        self.hyper_done = [None] * 8
        self.n_run = 0
        self.graph = None
        self.launches_per_replay = 0

    def _body(self):
        eng = self.eng
        eng.forward(self.embed, train=True, target=self.frames, p_norm=self.p_norm, mean_pixels=self.mean_pixels, want_img=False)
        flat = eng.backward()
        if self.world > 1:  # the NCCL all-reduce is captured into the graph with the kernels around it
            torch.distributed.all_reduce(flat, group=self.group)
        if self.phase == "alpha":
            adaround_step_multi(eng, self.opt, self.hyper)  # quantiser Jacobian + Adam for all 14 tensors: one launch
        else:  # step-size phase (5 % of a run): straight-through d_delta per tensor, Adam with the device-side step size
            grads = eng.param_grads(1.0)
            self.opt.step_dev([g for pair in grads for g in pair], self.hyper)
            eng.launches += len(self.opt.params)

    def run(self, embed, frames, reg_w: float, reg_b: float):
        self.embed.copy_(embed)
        self.frames.copy_(frames)
        step_size, bc2_sqrt = self.opt.hyper_of_next_step()
        k = self.n_run % len(self.hyper_host)
        self.n_run += 1
        if self.hyper_done[k] is not None:
            self.hyper_done[k].synchronize()
        hh = self.hyper_host[k]
        hh[0], hh[1], hh[2], hh[3] = reg_w, reg_b, step_size, bc2_sqrt
        self.hyper.copy_(hh, non_blocking=True)
        if self.hyper_done[k] is None:
            self.hyper_done[k] = torch.cuda.Event()
        self.hyper_done[k].record()
        if not self.capture:
            self._body()
        elif self.graph is None:
            l0 = self.eng.launches
            self._body()  # eager once: allocates every lazily-created buffer, and is this iteration's real work
            self.launches_per_replay = self.eng.launches - l0
            self._plan_of_body, self._mean_of_body = self.eng._last_plan, self.eng._mean_pixels
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            saved = [t.clone() for t in self.opt.params + self.opt.m + self.opt.v]  # capture must not advance the state
            # thread_local: other threads (NCCL's watchdog, bench.py's clock sampler) may call the CUDA API meanwhile
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._body()
            self.eng.launches -= self.launches_per_replay  # launches issued during capture do not execute
            for t, s in zip(self.opt.params + self.opt.m + self.opt.v, saved):
                t.copy_(s)
            self.graph = g
        else:
            self.graph.replay()
            self.eng.launches += self.launches_per_replay
            # what the Python side of forward() would have noted (last_loss() reads it): with a ragged last batch the loop
            # alternates between the graphs of two batch sizes
            self.eng._last_plan, self.eng._mean_pixels = self._plan_of_body, self._mean_of_body
            self.eng._packed_for = (self._plan_of_body.n, self._plan_of_body.h0, self._plan_of_body.w0)
        self.eng.invalidate()


class HostBatchPipe:
    """Host -> device input pipe for calibration batches that live in host memory (quantization/calib_model._FrameSource
    feeds the loop through it while frames are not resident in HBM).

    `put(*host_tensors)` starts the copy of the NEXT batch on a side stream -- straight from the source when it is pinned
    (a DataLoader with pin_memory=True), else through a pinned staging slot; `get()` makes the compute stream wait for the
    oldest pending batch and returns its device tensors.  With one batch in flight the PCIe copy of batch k+1 (4.9 MB for
    two 1280x640 uint8 frames, 19.7 MB as fp32) runs under the kernels of batch k instead of in front of them.  A slot is
    refilled only after the compute stream has consumed it (event recorded by the next `get()` or by `release()`).

    specs: one (shape, dtype) per tensor of a batch."""

    def __init__(self, specs, device="cuda", depth: int = 2):
        self.stream = torch.cuda.Stream(device=device)
        self.specs = [(tuple(sh), dt) for sh, dt in specs]
        self.slots = [tuple(torch.empty(sh, device=device, dtype=dt) for sh, dt in self.specs) for _ in range(depth)]
        self.pinned = [None] * depth  # staging for unpinned sources, allocated on first use
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [None] * depth  # event after which the slot may be overwritten
        self.head = self.tail = 0   # next slot to fill / next slot to hand out
        self.depth = depth
        self._last = None

    def pending(self) -> int:
        return self.head - self.tail

    def put(self, *host_tensors):
        if self.head - self.tail >= self.depth:
            raise RuntimeError("HostBatchPipe: all slots are pending; call get() first")
        if len(host_tensors) != len(self.specs):
            raise ValueError(f"HostBatchPipe: expected {len(self.specs)} tensors per batch, got {len(host_tensors)}")
        k = self.head % self.depth
        with torch.cuda.stream(self.stream):
            if self.free[k] is not None:
                self.stream.wait_event(self.free[k])
            for j, (dst, src) in enumerate(zip(self.slots[k], host_tensors)):
                if tuple(src.shape) != tuple(dst.shape) or src.dtype != dst.dtype:
                    raise ValueError(f"HostBatchPipe: batch tensor {j} is {tuple(src.shape)} {src.dtype}, slot is {tuple(dst.shape)} {dst.dtype}")
                if src.is_cuda:
                    dst.copy_(src, non_blocking=True)
                    continue
                if not src.is_pinned():
                    if self.pinned[k] is None:
                        self.pinned[k] = [torch.empty(sh, dtype=dt).pin_memory() for sh, dt in self.specs]
                    if self.free[k] is not None:
                        self.free[k].synchronize()  # the staging buffer's previous copy has left the host
                    self.pinned[k][j].copy_(src)
                    src = self.pinned[k][j]
                dst.copy_(src.contiguous(), non_blocking=True)
            self.ready[k].record(self.stream)
        self.head += 1

    def get(self):
        if self.tail >= self.head:
            raise RuntimeError("HostBatchPipe: nothing pending; call put() first")
        self.release()
        k = self.tail % self.depth
        torch.cuda.current_stream().wait_event(self.ready[k])
        self.tail += 1
        self._last = k
        return self.slots[k]

    def release(self):
        """Mark the batch handed out last as consumed by everything enqueued on the compute stream so far."""
        if self._last is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.free[self._last] = ev
            self._last = None


class CalibrationLoop:
    """One calibration run.  `fetch(idx) -> (embed, frames)` returns this rank's shard of the
    mini-batch `idx` as device tensors (NCHW)."""

    def __init__(self, engine: DecoderEngine, fetch: Callable, n_batches: int, iters: int, weight: float = 0.01,
                 b_range=(20, 2), warmup: float = 0.0, p: float = 2.0, lr: float = 0.0015,
                 group=None, global_batch: Optional[int] = None, log: Optional[list] = None,
                 on_iteration: Optional[Callable] = None):
        self.eng, self.fetch, self.n_batches, self.iters = engine, fetch, n_batches, iters
        self.weight, self.b_range, self.warmup, self.p, self.lr = weight, b_range, warmup, p, lr
        self.group = group
        self.world = 1
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
        self.global_batch = global_batch
        self.log = log
        self.on_iteration = on_iteration  # (phase, count, device loss scalar) after every iteration; no sync
        # CUDA-graph replay of the iteration, the NCCL all-reduce included when data parallel (NQ_GRAPH_DP=0: the same
        # fused sequence launched eagerly around an eager all-reduce); NQ_GRAPH=0: per-tensor eager path
        self.use_graph = os.environ.get("NQ_GRAPH", "1") != "0"
        self.capture = self.world == 1 or os.environ.get("NQ_GRAPH_DP", "1") != "0"
        self._graphed = {}
        self.ep1 = int(0.05 * iters / n_batches)  # calib_model.py:144
        self.ep2 = int(iters / n_batches) - self.ep1  # calib_model.py:205
        self.count = 0

    # -- one iteration: forward + loss + backward + (all-reduce) + Jacobian + Adam
    def iteration(self, idx, opt: AdamState, reg_w: float, reg_b: float, want_log: bool):
        eng = self.eng
        res = self.fetch(idx)
        embed, frames = res[0], res[1]
        n, _, H, W = frames.shape
        # lp_loss is a mean over the frames ACTUALLY in the (global) mini-batch (quantizer.py:71): a loader with
        # drop_last=False ends an epoch on a ragged batch, whose size `fetch` reports as a third value
        gb = res[2] if len(res) > 2 else (self.global_batch if self.global_batch is not None else n * self.world)
        if self.use_graph and eng.mode in ("ada", "uaq") and eng.stage_state is None and not want_log:
            key = (tuple(embed.shape), tuple(frames.shape), id(opt), gb)
            gs = self._graphed.get(key)
            if gs is None:
                gs = self._graphed[key] = GraphedStep(eng, opt, embed, frames, self.p, float(gb * H * W), self.group, self.world,
                                                      capture=self.capture, phase="alpha" if eng.mode == "ada" else "delta")
            gs.run(embed, frames, reg_w, reg_b)
            return
        eng.forward(embed, train=True, target=frames, p_norm=self.p, mean_pixels=float(gb * H * W),
                    reg_b=reg_b if (reg_w != 0.0 and want_log) else None, want_img=False)
        flat = eng.backward()
        if self.world > 1:
            torch.distributed.all_reduce(flat, group=self.group)
        grads = eng.param_grads(1.0, reg_w, reg_b)
        opt.step([g for pair in grads for g in pair])
        eng.launches += len(opt.params)
        eng.invalidate()

    def run_phase1(self, batches: Callable[[], Iterable]):
        eng = self.eng
        eng.mode = "uaq"
        params = []
        for s in eng.stages:
            params += [s.delta_w, s.delta_b]
        opt = AdamState(params, lr=0.001)  # calib_model.py:134
        count = 0
        for _ in range(self.ep1):
            for idx in batches():
                count += 1
                self.iteration(idx, opt, 0.0, 0.0, self.log is not None)
                if self.on_iteration is not None:
                    self.on_iteration("delta", count, self.eng.last_loss())
                if self.log is not None:
                    self.log.append(("delta", count, float(self._global_loss()), 0.0, 0.0))
        return count

    def run_phase2(self, batches: Callable[[], Iterable], on_start: Optional[Callable] = None):
        eng = self.eng
        eng.start_adaround()
        if on_start is not None:
            on_start()
        params = []
        for s in eng.stages:
            params += [s.alpha_w, s.alpha_b]
        opt = AdamState(params, lr=self.lr)
        decay = LinearTempDecay(self.iters, rel_start_decay=self.warmup, start_b=self.b_range[0], end_b=self.b_range[1])
        loss_start = self.iters * self.warmup
        count = 0
        for _ in range(self.ep2):
            for idx in batches():
                count += 1
                b = decay(count)
                reg_on = not (count < loss_start)
                want_log = self.log is not None or count % 500 == 0
                self.iteration(idx, opt, self.weight if reg_on else 0.0, float(b) if reg_on else 0.0, want_log)
                if self.on_iteration is not None:
                    self.on_iteration("alpha", count, self.eng.last_loss())
                if want_log:
                    rec = float(self._global_loss())
                    rnd = float(eng.reg_sum) * self.weight if reg_on else 0.0
                    if self.log is not None:
                        self.log.append(("alpha", count, rec, rnd, float(b) if reg_on else 0.0))
                    if count % 500 == 0:  # calib_model.py:86-88
                        logging.info('Total loss:\t{:.4f} (rec:{:.4f}, round:{:.4f})\tb={:.2f}\tcount={}'.format(
                            rec + rnd, rec, rnd, b if reg_on else 0, count))
        eng.soft_w = False  # calib_model.py:231-240 (bias quantisers stay soft)
        eng.invalidate()
        return count

    def _global_loss(self) -> torch.Tensor:
        loss = self.eng.last_loss().clone()
        if self.world > 1:
            torch.distributed.all_reduce(loss, group=self.group)
        return loss

    def run(self, batches: Callable[[], Iterable]):
        n1 = self.run_phase1(batches)
        n2 = self.run_phase2(batches)
        return n1, n2
