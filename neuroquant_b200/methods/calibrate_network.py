"""Network-wise calibration command line (reference: methods/calibrate_network.py): same flags, same
output-directory naming, same checkpoint file name and whole-object pickle layout.

    python -m neuroquant_b200.methods.calibrate_network --config configs/HNeRV/Bunny_1280x640_3M.yaml \\
        --arch hnerv --data_path bunny --vid Bunny --batch_size 2 --precision 6 5 4 5 5 6 6 --channel_wise \\
        --iters_w 21000 --weight 0.01 --b_start 20 --b_end 2 --warmup 0.2 --lr 0.003 --ckpt epoch300.pth
    torchrun --nproc-per-node 8 -m neuroquant_b200.methods.calibrate_network ...   # frame-sharded data parallel

Flags the reference parses but ignores are kept and ignored the same way: --opt_mode (kwargs
hard-code 'mse'), --input_prob (logged / file name only).  --seed: the reference never calls seed_all, so its shuffle is
unseeded; here it seeds the mini-batch shuffle, because data-parallel ranks must draw the same batches.
"""
import argparse
import logging
import os
import sys
import time
from datetime import datetime

import torch
from torch.utils.data import Subset

from ..compat import install_reference_aliases
from ..quantization import QuantModel, model_reconstruction
from ..utils import RoundTensor, data_split, get_config, setup_logger, worker_init_fn
from ..videosets import VideoDataSet
from .common import build_model, evaluate, init_distributed


def parse_args(argv):
    p = argparse.ArgumentParser(description="running parameters", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--seed", default=903, type=int, help="random seed for results reproduction")
    p.add_argument("--outf", default="unify", help="folder to output images and model checkpoints")
    p.add_argument("--config", type=str, help="config file path")
    p.add_argument("--arch", type=str, help="the architecture of NeRV")
    p.add_argument("-p", "--print-freq", default=50, type=int)
    p.add_argument("--data_path", type=str, help="data path for vid")
    p.add_argument("--vid", type=str, help="video id")
    p.add_argument("--data_split", type=str, default="1_1_1")
    p.add_argument("--batch_size", default=12, type=int, help="mini-batch size for data loader (global, across ranks)")
    p.add_argument("--precision", type=int, nargs="+", default=[8, 8, 8, 8, 8, 8, 8], help="layer-wise precision")
    p.add_argument("--channel_wise", action="store_true", help="apply channel_wise quantization for weights")
    p.add_argument("--hadamard", action="store_true", help="apply hadamard transform for weights")
    p.add_argument("--iters_w", default=20000, type=int, help="number of iteration for adaround")
    p.add_argument("--weight", default=0.01, type=float, help="weight of rounding cost vs the reconstruction loss.")
    p.add_argument("--b_start", default=20, type=int, help="temperature at the beginning of calibration")
    p.add_argument("--b_end", default=2, type=int, help="temperature at the end of calibration")
    p.add_argument("--warmup", default=0.2, type=float, help="in the warmup period no regularization is applied")
    p.add_argument("--input_prob", default=1.0, type=float)
    p.add_argument("--lr", default=0.0015, type=float)
    p.add_argument("--norm_p", default=2.0, type=float, help="the norm of L-p")
    p.add_argument("--init", default="max", type=str, choices=["max", "mse", "gaussian", "l1", "l2"])
    p.add_argument("--opt_mode", default="mse", type=str, choices=["mse", "fisher_diag", "fisher_full", "lp_norm"])
    p.add_argument("--ckpt", default="None", type=str, help="model for test")
    p.add_argument("--dump_vis", action="store_true", default=False, help="dump the prediction images")
    return p.parse_args(argv)


def _log_results(results):
    s = "Evaluation ... \n {} \n".format(datetime.now().strftime("%Y_%m_%d_%H_%M_%S"))
    s += f"best_pred_seen_psnr: {RoundTensor(results[0].max(), 2)} | "
    logging.info(s)


def calibrate(args, cfg):
    rank, world, _ = init_distributed()
    device = "cuda"
    full_dataset = VideoDataSet(cfg, args)
    full_loader = torch.utils.data.DataLoader(full_dataset, batch_size=cfg["batch_size"], shuffle=False,
                                              num_workers=cfg["workers"], pin_memory=True, drop_last=False,
                                              worker_init_fn=worker_init_fn)
    args.final_size = full_dataset.final_size
    args.full_data_length = len(full_dataset)
    split = [int(x) for x in args.data_split.split("_")]
    train_idx, args.val_ind_list = data_split(list(range(args.full_data_length)), split, False, 0)
    gen = torch.Generator()
    gen.manual_seed(args.seed)  # every rank must draw the same shuffled batches; the reference is unseeded (Q7)
    # the calibration loader hands out uint8 frames (value / 255 happens inside the head-loss kernel): a quarter of the
    # PCIe bytes in the first epoch and of the resident clip in HBM afterwards
    train_dataset = VideoDataSet(cfg, args, as_uint8=True)
    train_loader = torch.utils.data.DataLoader(Subset(train_dataset, train_idx), batch_size=args.batch_size, shuffle=True,
                                               num_workers=cfg["workers"], pin_memory=True, drop_last=True,
                                               worker_init_fn=worker_init_fn, persistent_workers=cfg["workers"] > 0,
                                               generator=gen)
    model = build_model(args, cfg).to(device)
    args.outf = os.path.join(args.outf, f"Encoder_{round(args.encoder_param, 2)}M_Decoder_{round(args.decoder_param, 2)}M_"
                                        f"Total_{round(args.total_param, 2)}M")
    args.outf = os.path.join(args.outf, "network-wise_calib/hadamard-{}_{}-init_batch{}_CW_weight{}_brange{}-{}_warmup{}_lr{}".format(
        args.hadamard, args.init, args.batch_size, args.weight, args.b_start, args.b_end, args.warmup, args.lr))
    if rank == 0:
        os.makedirs(args.outf, exist_ok=True)
        setup_logger(args.outf + "/" + time.strftime("%Y%m%d_%H%M%S") + ".log")
    logging.info("[PID] %s" % os.getpid())
    logging.info("================== Model Architecture=================")
    logging.info(str(model))
    assert args.ckpt != "None"
    logging.info("=> loading checkpoint '{}'".format(args.ckpt))
    model.load_state_dict(torch.load(args.ckpt, map_location="cpu"), strict=False)
    model.to(device)

    logging.info("=======================Full-precision model========================")
    results, _, embedding_list = evaluate(model, full_loader, args, cfg, args.dump_vis)
    _log_results(results)

    wq_params = {"n_bits": 8, "channel_wise": args.channel_wise, "scale_method": args.init}
    qnn = QuantModel(model=model, hadamard=args.hadamard, weight_quant_params=wq_params).to(device)
    args.qbits = qnn.set_bitwidth(args.precision)
    qnn.eval()
    logging.info("quantized model architecture: {}".format(qnn))
    cali_data = torch.cat(embedding_list, dim=0)
    logging.info("input embedding shape: {}".format(cali_data.shape))

    qnn.set_quant_state(True)
    t0 = time.time()
    _ = qnn(cali_data[:args.batch_size].to(device))
    logging.info("Init time: {}".format(time.time() - t0))

    logging.info("=======================Close quantization model========================")
    qnn.set_quant_state(False)
    _log_results(evaluate(qnn, full_loader, args, cfg, args.dump_vis)[0])
    logging.info("=======================Weight quantization model w/o opt========================")
    qnn.set_quant_state(True)
    _log_results(evaluate(qnn, full_loader, args, cfg, args.dump_vis)[0])

    kwargs = dict(cali_data=cali_data, gt=train_loader, arch=args.arch, batch_size=args.batch_size, iters=args.iters_w,
                  weight=args.weight, opt_mode="mse", hadamard=args.hadamard, b_range=(args.b_start, args.b_end),
                  warmup=args.warmup, p=args.norm_p, lr=args.lr)
    logging.info("======================= Hyper Parameters =======================")
    for k in ("init", "channel_wise", "seed", "iters_w", "batch_size", "weight", "input_prob", "qbits"):
        logging.info("{}: {}".format(k, getattr(args, k)))
    logging.info(f"begin training in {device} x {world}")
    start = datetime.now()
    qnn.set_quant_state(weight_quant=True)
    model_reconstruction(qnn, **kwargs)
    logging.info(f"Training complete in: {str(datetime.now() - start)}")
    qnn.set_quant_state(weight_quant=True)
    logging.info("=======================Weight quantization model w/ opt========================")
    _log_results(evaluate(qnn, full_loader, args, cfg, args.dump_vis)[0])
    if rank == 0:
        logging.info("save quantized model in {}".format(args.outf))
        install_reference_aliases()  # pickle GLOBALs under the reference's module paths
        torch.save(qnn, "{}/{}_W{}_prob{}_{}-init_{}.pth".format(args.outf, args.arch, args.qbits, args.input_prob, args.init,
                                                                  "CW" if args.channel_wise else "LW"))
    return qnn


def main(argv):
    args = parse_args(argv)
    cfg = get_config(args.config)
    args.outf = os.path.join("results", args.outf)
    args.exp_id = f"{args.vid}_e{cfg['epoch']}_b{cfg['batch_size']}_lr{cfg['learning_rate']}_{cfg['loss']}"
    args.outf = os.path.join(args.outf, args.exp_id)
    torch.set_printoptions(precision=2)
    calibrate(args, cfg)


if __name__ == "__main__":
    main(sys.argv[1:])
