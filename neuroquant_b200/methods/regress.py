"""FP32 regression training of the decoder on the engine's convolution kernels (reference: methods/regress.py:151-322;
SURVEY 8(f) rank 4 -- the step that produces the `epoch300.pth` the calibration starts from).

Per iteration (regress.py:249-261): learning rate from adjust_lr, forward, loss_fn 'l2' (utils.py:115-116: per-frame
mean over C*H*W, batch mean), backward, Adam on every parameter.  Here the decoder -- stem, blocks, head: the
convolutions that hold all of NeRV's and nearly all of HNeRV's training FLOPs -- runs on the same tcgen05 forward / data
gradient / weight gradient kernels as the calibration loop, with the engine in full-precision mode, and its Adam steps are
CUDA kernels of the library.  HNeRV's ConvNeXt frame encoder stays stock PyTorch (as everywhere in this package): the
engine hands back dL/d(embedding) and autograd carries it through the encoder, whose parameters step in a torch Adam
at the same learning rate.  Losses: 'l2' and 'l1'; the SSIM mixtures are not provided.
"""
from __future__ import annotations

import logging
import os
import sys
import time
from datetime import datetime

import torch
import torch.nn.functional as F
from torch.utils.data import Subset

from ..engine import AdamState
from ..runner import DecoderRunner
from ..utils import RoundTensor, adjust_lr, data_split, get_config, psnr_fn_single, setup_logger, worker_init_fn


class DecoderTrainer:
    """One model's training state: the engine binding, a fused Adam over the decoder's (weight, bias) tensors and, for
    HNeRV, a torch Adam over the encoder.  `param_groups` makes it acceptable to utils.adjust_lr."""

    def __init__(self, model, arch: str, lr: float, loss: str = "l2"):
        if loss not in ("l2", "l1"):
            raise NotImplementedError(f"loss {loss!r}: 'l2' and 'l1' run on the fused head kernel (utils.py:115-118); the "
                                      "SSIM mixtures are not provided")
        self.model, self.arch = model, arch
        self.p_norm = 2.0 if loss == "l2" else 1.0
        self.runner = DecoderRunner.of(model)
        if any(self.runner._quant):
            raise ValueError("regression training takes the full-precision model, not a QuantModel")
        self.params = [t for l in self.runner.layers for t in (l.weight.data, l.bias.data)]
        self.opt = AdamState(self.params, lr)
        enc = [p for n, p in model.named_parameters() if n.startswith("encoder")] if arch == "hnerv" else []
        self.enc_opt = torch.optim.Adam(enc, weight_decay=0.) if enc else None
        self.param_groups = [{"lr": lr}] + (self.enc_opt.param_groups if self.enc_opt else [])
        self.launches = 0
        # decoder-only models (NeRV): the whole step -- weight pack, forward + loss, backward, Adam -- is one CUDA graph,
        # replayed with the batch in static buffers and the learning rate in a 16-byte device array (NQ_GRAPH=0: eager)
        self.use_graph = self.enc_opt is None and os.environ.get("NQ_GRAPH", "1") != "0"
        self._graphs = {}
        self._hyper_host = [torch.zeros(4).pin_memory() for _ in range(8)] if self.use_graph else []
        self._hyper_done = [None] * 8
        self._n_run = 0

    def _step_graphed(self, embed: torch.Tensor, frames: torch.Tensor):
        eng = self.runner.engine
        key = (tuple(embed.shape), tuple(frames.shape))
        st = self._graphs.get(key)
        if st is None:
            st = self._graphs[key] = {"embed": torch.empty_like(embed, dtype=torch.float32), "frames": torch.empty_like(frames),
                                      "hyper": torch.zeros(4, device=frames.device), "graph": None, "launches": 0}
        st["embed"].copy_(embed)
        st["frames"].copy_(frames)
        self.opt.lr = float(self.param_groups[0]["lr"])
        step_size, bc2 = self.opt.hyper_of_next_step()
        k = self._n_run % len(self._hyper_host)
        self._n_run += 1
        if self._hyper_done[k] is not None:
            self._hyper_done[k].synchronize()
        hh = self._hyper_host[k]
        hh[2], hh[3] = step_size, bc2
        st["hyper"].copy_(hh, non_blocking=True)
        if self._hyper_done[k] is None:
            self._hyper_done[k] = torch.cuda.Event()
        self._hyper_done[k].record()
        n, _, hh_, ww_ = frames.shape

        def body():
            eng.forward(st["embed"], train=True, target=st["frames"], p_norm=self.p_norm, mean_pixels=float(n * 3 * hh_ * ww_))
            eng.backward()
            _, views = eng._grad_buffers()
            return self.opt.step_dev([g for pair in views for g in pair], st["hyper"])

        if st["graph"] is None:
            l0 = eng.launches
            body()                              # eager once: this iteration's real work, and every lazy allocation
            st["launches"] = eng.launches - l0
            torch.cuda.synchronize()
            loss, img = eng.last_loss().clone(), eng._last_plan.img.clone()
            g = torch.cuda.CUDAGraph()
            state = self.opt.params + self.opt.m + self.opt.v
            saved = [t.clone() for t in state]  # capture must not advance the state
            eng.invalidate()
            with torch.cuda.graph(g):
                body()
            eng.launches = l0 + st["launches"]
            for t, sv in zip(state, saved):
                t.copy_(sv)
            st["graph"] = g
        else:
            st["graph"].replay()
            eng.launches += st["launches"]
            loss, img = eng.last_loss().clone(), eng._last_plan.img
        self.launches += len(self.opt.params)  # the Adam kernels (the engine counts its own)
        eng.invalidate()
        return loss, img

    def step(self, inputs: torch.Tensor, frames: torch.Tensor):
        """inputs: the frames themselves (hnerv) or their normalised indices (nerv).  Returns (loss, img_out), both on
        the device; the loss is the value BEFORE the update, as regress.py:260 logs it."""
        eng = self.runner.engine
        if self.enc_opt is not None:
            embed = self.model.encode(inputs)
        else:
            with torch.no_grad():
                embed = self.model.encode(inputs)
        self.runner.sync()
        if self.use_graph:
            return self._step_graphed(embed, frames)
        n, _, hh, ww = frames.shape
        img = eng.forward(embed.detach(), train=True, target=frames, p_norm=self.p_norm, mean_pixels=float(n * 3 * hh * ww),
                          reuse_weights=True)
        loss = eng.last_loss().clone()
        flat = eng.backward()
        _, views = eng._grad_buffers()
        if self.enc_opt is not None:
            # dL/d(embedding): the stem's data gradient, a k x k (1 x 1 in every shipped config) transposed convolution of
            # a (n, C, h0, w0) map of a few hundred values -- left to torch, like the encoder it feeds
            st0 = eng.stages[0]
            d_embed = F.conv_transpose2d(eng.stage_output_grad(0), st0.weight, padding=st0.geom.k // 2)
            self.enc_opt.zero_grad()
            embed.backward(d_embed)
            self.enc_opt.step()
        self.opt.lr = float(self.param_groups[0]["lr"])
        self.launches += self.opt.step([g for pair in views for g in pair])
        eng.invalidate()  # the kernels updated the weights in place
        return loss, img


def train(args, cfg):
    """regress.py:151-322 without TensorBoard: data set, model, per-epoch training + evaluation, checkpoints
    (`model_latest.pth`, `epoch<N>.pth`: plain state_dicts, loadable by the reference)."""
    from ..videosets import VideoDataSet
    from .common import build_model, evaluate, init_distributed
    rank, world, _ = init_distributed()
    if world > 1:
        raise NotImplementedError("regression training is single-GPU, as in the reference")
    if cfg["loss"] not in ("l2", "l1"):
        raise NotImplementedError(f"loss {cfg['loss']!r}: only 'l2' / 'l1' run on the fused head kernel")
    device = "cuda"
    full_dataset = VideoDataSet(cfg, args)
    full_loader = torch.utils.data.DataLoader(full_dataset, batch_size=cfg["batch_size"], shuffle=False, num_workers=cfg["workers"],
                                              pin_memory=True, drop_last=False, worker_init_fn=worker_init_fn)
    args.final_size = full_dataset.final_size
    args.full_data_length = len(full_dataset)
    split = [int(x) for x in args.data_split.split("_")]
    train_idx, args.val_ind_list = data_split(list(range(args.full_data_length)), split, False, 0)
    gen = torch.Generator()
    gen.manual_seed(args.seed)
    train_loader = torch.utils.data.DataLoader(Subset(full_dataset, train_idx), batch_size=cfg["batch_size"], shuffle=True,
                                               num_workers=cfg["workers"], pin_memory=True, drop_last=True,
                                               worker_init_fn=worker_init_fn, generator=gen)
    model = build_model(args, cfg).to(device)
    args.outf = os.path.join(args.outf, f"Encoder_{round(args.encoder_param, 2)}M_Decoder_{round(args.decoder_param, 2)}M_"
                                        f"Total_{round(args.total_param, 2)}M")
    os.makedirs(args.outf, exist_ok=True)
    setup_logger(args.outf + "/" + time.strftime("%Y%m%d_%H%M%S") + ".log")
    logging.info("[PID] %s" % os.getpid())
    logging.info(str(model))
    if args.weight != "None":
        logging.info("=> loading checkpoint '{}'".format(args.weight))
        model.load_state_dict(torch.load(args.weight, map_location="cpu"), strict=False)
        model.to(device)
    if args.eval_only:
        results, _, _ = evaluate(model, full_loader, args, cfg, args.dump_vis)
        logging.info(f"best_pred_seen_psnr: {RoundTensor(results[0].max(), 2)} | ")
        return
    args.lr = cfg["learning_rate"]
    trainer = DecoderTrainer(model, args.arch, args.lr, cfg["loss"])
    start = datetime.now()
    # The loader decodes every PNG once (first epoch); the frames then stay resident in HBM and later epochs only draw the
    # loader's shuffled index batches -- at several hundred iterations/s four decoding workers cannot keep up
    # (same scheme as quantization/calib_model._FrameSource).
    n_full = args.full_data_length
    train_pos = torch.as_tensor(train_idx, device=device)
    frames, have = None, torch.zeros(n_full, dtype=torch.bool, device=device)
    for epoch in range(cfg["epoch"]):
        model.train()
        epoch_start, psnrs = datetime.now(), []
        resident = frames is not None and bool(have[train_pos].all())
        for i, sample in enumerate(train_loader.batch_sampler if resident else train_loader):
            cur_epoch = (epoch + float(i) / len(train_loader)) / cfg["epoch"]
            lr = adjust_lr(trainer, cur_epoch, args)
            if resident:
                idx = train_pos[torch.as_tensor(sample, device=device)]     # positions in the Subset -> frame numbers
                img, norm_idx = frames[idx], idx.double() / n_full                # float(idx) / len(video), as the data set
            else:
                img = sample["img"].to(device, non_blocking=True).float()
                idx, norm_idx = sample["idx"].to(device).view(-1), sample["norm_idx"].to(device)
                if frames is None:
                    frames = torch.empty((n_full,) + tuple(img.shape[1:]), device=device)
                frames[idx], have[idx] = img, True
            inputs = img if args.arch == "hnerv" else norm_idx
            _, img_out = trainer.step(inputs, img)
            psnrs.append(psnr_fn_single(img_out, img))
            if i % args.print_freq == 0 or i == len(train_loader) - 1:
                logging.info("[{}], Epoch[{}/{}], Step [{}/{}], lr:{:.2e} pred_PSNR: {}".format(
                    datetime.now().strftime("%Y/%m/%d %H:%M:%S"), epoch + 1, cfg["epoch"], i + 1, len(train_loader), lr,
                    RoundTensor(torch.cat(psnrs).mean(), 2)))
        logging.info("Time/epoch: \tCurrent:{:.2f} \tAverage:{:.2f}".format(
            (datetime.now() - epoch_start).total_seconds(), (datetime.now() - start).total_seconds() / (epoch + 1)))
        if (epoch + 1) % cfg["eval_freq"] == 0 or (cfg["epoch"] - epoch) in [1, 3, 5]:
            results, hw, _ = evaluate(model, full_loader, args, cfg, args.dump_vis if epoch == cfg["epoch"] - 1 else False)
            logging.info(f"Eval at epoch {epoch + 1} for {hw}: pred_seen_psnr: {RoundTensor(results[0], 2)} | ")
        torch.save(model.state_dict(), "{}/model_latest.pth".format(args.outf))
        if (epoch + 1) % cfg["epoch"] == 0:
            torch.save(model.state_dict(), f"{args.outf}/epoch{epoch + 1}.pth")
    logging.info(f"Training complete in: {str(datetime.now() - start)}")


def parse_args(argv):
    import argparse
    parser = argparse.ArgumentParser()  # regress.py:33-55
    parser.add_argument("--seed", default=903, type=int)
    parser.add_argument("--outf", default="unify")
    parser.add_argument("--config", type=str)
    parser.add_argument("--arch", type=str)
    parser.add_argument("--data_path", type=str)
    parser.add_argument("--vid", type=str)
    parser.add_argument("--data_split", type=str, default="1_1_1")
    parser.add_argument("-p", "--print-freq", default=50, type=int)
    parser.add_argument("--lr_type", type=str, default="cosine_0.1_1_0.1")
    parser.add_argument("--weight", default="None", type=str)
    parser.add_argument("--eval_only", action="store_true", default=False)
    parser.add_argument("--dump_vis", action="store_true", default=False)
    parser.add_argument("--eval_fps", action="store_true", default=False)
    return parser.parse_args(argv)


def main(argv):
    args = parse_args(argv)
    cfg = get_config(args.config)
    torch.manual_seed(args.seed)
    args.outf = os.path.join("results", args.outf)
    args.exp_id = f"{args.vid}_e{cfg['epoch']}_b{cfg['batch_size']}_lr{cfg['learning_rate']}_{cfg['loss']}"
    args.outf = os.path.join(args.outf, args.exp_id)
    train(args, cfg)


if __name__ == "__main__":
    main(sys.argv[1:])
