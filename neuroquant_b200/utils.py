"""Helpers with the reference's names (utils.py): YAML config, logger, data split, PSNR, the LR schedule of the FP
regression (methods/regress.py).  MS-SSIM and the SSIM-mixing training losses are not provided."""
import logging
import math
import random
import sys

import numpy as np
import torch
import yaml

from . import _lib as L


def setup_logger(log_dir):
    fmt = logging.Formatter("%(asctime)s [%(levelname)-5.5s]  %(message)s")
    root = logging.getLogger()
    root.setLevel(logging.INFO)
    fh = logging.FileHandler(log_dir, encoding="utf-8")
    fh.setFormatter(fmt)
    root.addHandler(fh)
    sh = logging.StreamHandler(sys.stdout)
    sh.setFormatter(fmt)
    root.addHandler(sh)
    logging.info("Logging file is %s" % log_dir)


def get_config(config_path):
    with open(config_path, "r") as stream:
        return yaml.load(stream, Loader=yaml.FullLoader)


def data_split(img_list, split_num_list, shuffle_data, rand_num=0):
    """utils.py:42-53."""
    valid_train_length, total_train_length, total_data_length = split_num_list
    train, val = [], []
    if shuffle_data:
        random.Random(rand_num).shuffle(img_list)
    for cur_i, frame_id in enumerate(img_list):
        if (cur_i % total_data_length) < valid_train_length:
            train.append(frame_id)
        elif (cur_i % total_data_length) >= total_train_length:
            val.append(frame_id)
    return train, val


def worker_init_fn(worker_id):
    seed = torch.initial_seed() % 2 ** 32
    np.random.seed(seed)
    random.seed(seed)


def RoundTensor(x, num=2, group_str=False):
    if group_str:
        return "/".join(",".join(str(round(e, num)) for e in x[i].tolist()) for i in range(x.size(0)))
    return ",".join(str(round(e, num)) for e in x.flatten().tolist())


def psnr_fn_single(output, gt):
    """utils.py:148-151: per-frame -10 log10(mse + 1e-9); one fused reduction kernel on the device."""
    return L.psnr(output.detach().contiguous().float(), gt.detach().contiguous().float()).cpu()


def psnr_fn_batch(output_list, gt):
    return torch.stack([psnr_fn_single(o, gt) for o in output_list], 0).cpu()


def msssim_fn_single(output, gt):
    raise NotImplementedError("MS-SSIM is an evaluation cosmetic outside the calibration path (SURVEY row 14)")


def adjust_lr(optimizer, cur_epoch, args, eta_min=0.05):
    """utils.py:79-99: 'cosine_<up_ratio>_<up_pow>_<min_lr>' / 'hybrid_...' multiplier applied to every param group of
    `optimizer` (anything with a `param_groups` list of dicts, e.g. methods.regress.DecoderTrainer)."""
    if "hybrid" in args.lr_type:
        up_ratio, up_pow, down_pow, min_lr, final_lr = [float(x) for x in args.lr_type.split("_")[1:]]
        if cur_epoch < up_ratio:
            lr_mult = min_lr + (1. - min_lr) * (cur_epoch / up_ratio) ** up_pow
        else:
            lr_mult = 1 - (1 - final_lr) * ((cur_epoch - up_ratio) / (1. - up_ratio)) ** down_pow
    elif "cosine" in args.lr_type:
        up_ratio, up_pow, min_lr = [float(x) for x in args.lr_type.split("_")[1:]]
        if cur_epoch < up_ratio:
            lr_mult = min_lr + (1. - min_lr) * (cur_epoch / up_ratio) ** up_pow
        else:
            lr_mult = max(0.5 * (math.cos(math.pi * (cur_epoch - up_ratio) / (1 - up_ratio)) + 1.0), eta_min)
    else:
        raise NotImplementedError
    for param_group in optimizer.param_groups:
        param_group["lr"] = args.lr * lr_mult
    return args.lr * lr_mult
