"""Shared pieces of the two command lines: model construction, evaluation, distributed setup."""
import logging
import os
from datetime import datetime

import numpy as np
import torch

from ..models import HNeRV, NeRV
from ..parallel import frame_range_of_rank, world_info
from ..utils import RoundTensor, psnr_fn_batch


def init_distributed():
    """torchrun sets RANK / LOCAL_RANK / WORLD_SIZE: one process per GPU over NCCL.  Additive to the
    reference, which is single-GPU."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("neuroquant_b200 needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world_info()


def build_model(args, cfg):
    """calibrate_network.py:168-186 / bit_assign.py:286-300: model + the parameter counts used in the
    output directory name."""
    if args.arch == "hnerv":
        model = HNeRV(cfg)
        enc = sum(p.data.nelement() for p in model.encoder.parameters()) / 1e6
        dec = sum(p.data.nelement() for p in model.decoder.parameters()) / 1e6
        embed = float(cfg["enc_channel"][-1]) / np.prod(cfg["enc_strides"]) ** 2 * args.final_size * args.full_data_length
        total = dec + embed / 1e6
    elif args.arch == "nerv":
        model = NeRV(cfg)
        enc, embed = 0.0, 0.0
        dec = sum(p.data.nelement() for p in model.decoder.parameters()) / 1e6
        total = dec
    else:
        raise ValueError("model arch wrong!")
    args.encoder_param, args.decoder_param, args.total_param = enc, dec, total
    return model


@torch.no_grad()
def evaluate(model, full_dataloader, args, cfg, dump_vis=False):
    """calibrate_network.py:82-145: decode every frame, PSNR per frame (device kernel), FPS from the
    per-call decode time.  Frames whose index is in `args.val_ind_list` (--data_split) count as unseen, the rest as
    seen (:110-114); returns [seen PSNR, seen MS-SSIM, unseen PSNR, unseen MS-SSIM] like the reference, with the
    MS-SSIM entries zero (eval cosmetic, SURVEY row 14 -- not computed).  `dump_vis` writes ground truth | prediction
    side by side to {outf}/visualize_calib_network (:91-95, :116-123).  Under torch.distributed every rank decodes a
    contiguous range of frames (decode sharding) and the metrics are gathered."""
    rank, world, group = world_info()
    model.eval()
    device = next(model.parameters()).device
    embeds, psnrs, idxs, dec_times = [], [], [], []
    n_batches = len(full_dataloader)
    lo, hi = frame_range_of_rank(n_batches, rank, world)
    unseen = set(int(v) for v in getattr(args, "val_ind_list", []) or [])
    visual_dir = None
    if dump_vis:
        visual_dir = f"{args.outf}/visualize_calib_network"
        logging.info(f"Saving predictions to {visual_dir}...")
        os.makedirs(visual_dir, exist_ok=True)
    for i, sample in enumerate(full_dataloader):
        img, norm_idx, img_idx = sample["img"].to(device), sample["norm_idx"].to(device), sample["idx"].to(device)
        embed = model.encode(img) if args.arch == "hnerv" else model.encode(norm_idx)
        embeds.append(embed)
        if not (lo <= i < hi):
            continue
        img_out, _, dec_time = model.decode(embed)
        dec_times.append(dec_time)
        pred_psnr = psnr_fn_batch([img_out], img)
        psnrs.append(pred_psnr[0])
        idxs.append(img_idx.cpu().view(-1))
        if visual_dir is not None:
            from torchvision.utils import save_image
            for b in range(img.shape[0]):
                full_ind = i * cfg["batch_size"] + b
                tag = ",".join(str(round(x[b].item(), 2)) for x in pred_psnr)
                save_image(torch.cat([img[b], img_out[b]], dim=2), f"{visual_dir}/pred_{full_ind:04d}_{tag}.png")
        if (i - lo) % args.print_freq == 0 or i == hi - 1:
            fps = cfg["batch_size"] / (sum(dec_times) / len(dec_times))
            seen_now = [p for p, ix in zip(torch.cat(psnrs).tolist(), torch.cat(idxs).tolist()) if ix not in unseen]
            logging.info("[{}], Eval at Step [{}/{}], FPS {}, PSNR {}".format(
                datetime.now().strftime("%Y/%m/%d %H:%M:%S"), i + 1, n_batches, round(fps, 1),
                RoundTensor(torch.tensor(seen_now).mean().view(1) if seen_now else torch.zeros(1), 2)))
    psnr = torch.cat(psnrs) if psnrs else torch.zeros(0)
    idx = torch.cat(idxs) if idxs else torch.zeros(0, dtype=torch.long)
    if world > 1:
        bucket = [None] * world
        torch.distributed.all_gather_object(bucket, (psnr, idx, sum(dec_times), len(dec_times)), group=group)
        psnr = torch.cat([b[0] for b in bucket])
        idx = torch.cat([b[1] for b in bucket])
        args.fps = cfg["batch_size"] * sum(b[3] for b in bucket) / max(1e-12, sum(b[2] for b in bucket)) * world
    else:
        args.fps = cfg["batch_size"] / (sum(dec_times) / max(1, len(dec_times)))
    is_unseen = torch.tensor([int(v) in unseen for v in idx.tolist()], dtype=torch.bool)
    seen_psnr = psnr[~is_unseen].mean().view(1) if bool((~is_unseen).any()) else torch.zeros(1)
    unseen_psnr = psnr[is_unseen].mean().view(1) if bool(is_unseen.any()) else torch.zeros(1)
    args.psnr_per_frame = (idx, psnr)
    model.train()
    h, w = img.shape[-2:]
    return [seen_psnr, torch.zeros(1), unseen_psnr, torch.zeros(1)], (h, w), embeds
