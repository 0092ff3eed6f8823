"""Mixed-precision bit assignment by sensitivity score (reference: methods/bit_assign.py): same flags,
same two-candidate tables and log lines; the candidates are farmed out one per GPU under torchrun and
the scores gathered.  Additive: `--search --options 2 .. 8 --avg_bits B` scores EVERY per-layer
configuration within the budget (7^7 = 823 543 for 7 layers) from a Gram table of Omega measured with forward jets
farmed over the GPUs (sensitivity.OmegaTable, nq_omega_search).

    python -m neuroquant_b200.methods.bit_assign --config ... --arch hnerv --data_path bunny --vid Bunny \\
        --batch_size 2 --channel_wise --init max --mode omega --ckpt epoch300.pth
"""
import argparse
import copy
import logging
import os
import random
import sys
import time

import numpy as np
import torch

from ..parallel import candidates_of_rank, gather_scores
from ..quantization import QuantModel
from ..runner import DecoderRunner
from ..sensitivity import OmegaTable, fisher_diag, omega, omega_layers
from ..utils import data_split, get_config, setup_logger, worker_init_fn
from ..videosets import VideoDataSet
from .common import build_model, evaluate, init_distributed

# toy examples of the reference (bit_assign.py:28-36)
hnerv_candidate = {"candidate1": [2, 3, 4, 6, 4, 4, 2], "candidate2": [6, 5, 4, 5, 5, 6, 6]}
nerv_candidate = {"candidate1": [5, 6, 3, 4, 5, 4, 3], "candidate2": [6, 5, 5, 6, 7, 6, 7]}


def sensitivity_criterion(mode, arch, net, qnn, dataloader, use_cuda=True, max_batches=10, per_layer=True):
    """bit_assign.py:171-217.  `net` is the full-precision model, `qnn` the QuantModel whose perturbation
    W - Q(W) is scored; the first 10 batches of `dataloader` are used (:115-117).  per_layer: also log the reference's
    "[i-th layer]" terms (:194-200, :208-214); for 'omega' they cost 2 extra jets per layer (sensitivity.omega_layers) --
    pass False when only the score matters (candidate search)."""
    vec = qnn.get_perturbation()
    runner = DecoderRunner.of(net)
    for l in runner.layers:
        if hasattr(l, "set_quant_state"):
            l.set_quant_state(False)
    runner.sync()
    device = next(net.parameters()).device
    batches = []
    with torch.no_grad():
        for i, sample in enumerate(dataloader):
            img = sample["img"].to(device)
            embed = net.encode(img) if arch == "hnerv" else net.encode(sample["norm_idx"].to(device))
            batches.append((embed, img))
            if len(batches) >= max_batches:
                break
    if mode == "omega":
        if per_layer:
            for count, cur in enumerate(omega_layers(runner.engine, vec, batches)):
                logging.info(f"[{count:d}-th layer] {cur:.3e}")
        return torch.tensor(omega(runner.engine, vec, batches))
    if mode == "fisher_diag":
        per = fisher_diag(runner.engine, vec, batches, per_layer=True)
        if per_layer:
            for count, cur in enumerate(per):
                logging.info(f"[{count:d}-th layer] {cur:.3e}")
        return torch.tensor(sum(per))
    raise ValueError("Not implemented sensitivity criteria: {}".format(mode))


def first_batches(arch, net, dataloader, max_batches=10):
    """The mini-batches the reference's Hessian-vector product runs over (bit_assign.py:82-117: the first 10 of the
    loader), as (embedding, frames) pairs on the device."""
    device = next(net.parameters()).device
    batches = []
    with torch.no_grad():
        for sample in dataloader:
            img = sample["img"].to(device)
            embed = net.encode(img) if arch == "hnerv" else net.encode(sample["norm_idx"].to(device))
            batches.append((embed, img))
            if len(batches) >= max_batches:
                break
    return batches


def search_bit_assignment(arch, model, dataloader, cali_data, options, avg_bits_budget, hadamard=False, channel_wise=True,
                          init="max", batch_size=2, max_batches=10):
    """BASELINE.json configs[3]: the Omega-optimal per-layer bit-widths among ALL len(options)^L configurations whose
    average bit-width stays within the budget.  The perturbation of every layer at every bit-width comes from the
    reference's own construction -- QuantModel + set_bitwidth + first quantised forward + get_perturbation
    (bit_assign.py:346-359, quant_layer.py:86-89) -- once per option; sensitivity.OmegaTable measures the Gram table
    with forward jets (farmed over the ranks of torch.distributed) and nq_omega_search scores every configuration.
    Returns (bits, omega, average bits, OmegaTable)."""
    from ..parallel import world_info
    rank, world, group = world_info()
    device = next(model.parameters()).device
    pert, n_params = None, None
    for b in options:
        qnn = QuantModel(model=copy.deepcopy(model), hadamard=hadamard,
                         weight_quant_params={"n_bits": 8, "channel_wise": channel_wise, "scale_method": init}).to(device)
        qnn.eval()
        mods = qnn.quant_modules()
        qnn.set_bitwidth([b] * len(mods))
        qnn.set_quant_state(True)
        _ = qnn(cali_data[:batch_size].to(device))
        vec = [v.detach().clone() for v in qnn.get_perturbation()]
        if pert is None:
            pert = [[] for _ in vec]
            n_params = [m.weight.numel() + m.bias.numel() for m in mods]
        for l, v in enumerate(vec):
            pert[l].append(v)
        del qnn
    net = copy.deepcopy(model)
    runner = DecoderRunner.of(net)
    runner.sync()
    table = OmegaTable(runner.engine, pert, options, n_params)
    table.build(first_batches(arch, net, dataloader, max_batches), rank, world, group)
    bits, score, avg_bits, _ = table.search(avg_bits_budget)
    return bits, score, avg_bits, table


def assign(args, cfg):
    rank, world, _ = init_distributed()
    device = "cuda"
    full_dataset = VideoDataSet(cfg, args)
    gen = torch.Generator()
    gen.manual_seed(args.seed)
    loader = torch.utils.data.DataLoader(full_dataset, batch_size=args.batch_size, shuffle=True, num_workers=cfg["workers"],
                                         pin_memory=True, drop_last=False, worker_init_fn=worker_init_fn, generator=gen)
    args.final_size = full_dataset.final_size
    args.full_data_length = len(full_dataset)
    split = [int(x) for x in args.data_split.split("_")]
    _, args.val_ind_list = data_split(list(range(args.full_data_length)), split, False, 0)
    model = build_model(args, cfg).to(device)
    args.outf = os.path.join(args.outf, f"Encoder_{round(args.encoder_param, 2)}M_Decoder_{round(args.decoder_param, 2)}M_"
                                        f"Total_{round(args.total_param, 2)}M")
    args.outf = os.path.join(args.outf, "sensitivity-{}_{}-init_batch{}_CW".format(args.mode, args.init, args.batch_size))
    if rank == 0:
        os.makedirs(args.outf, exist_ok=True)
        setup_logger(args.outf + "/" + time.strftime("%Y%m%d_%H%M%S") + ".log")
    assert args.ckpt != "None"
    model.load_state_dict(torch.load(args.ckpt, map_location="cpu"), strict=False)
    model.to(device)
    logging.info("=======================Full-precision model========================")
    _, _, embedding_list = evaluate(model, loader, args, cfg)
    cali_data = torch.cat(embedding_list, dim=0)
    candidate_dict = hnerv_candidate if args.arch == "hnerv" else nerv_candidate
    names = list(candidate_dict)
    if getattr(args, "search", False):
        # additive: instead of the two toy candidates, every configuration over --options within --avg_bits
        if args.mode != "omega":
            raise NotImplementedError("--search scores configurations by Omega (a quadratic form); fisher_diag is not one")
        t0 = time.time()
        bits, score, avg_bits, table = search_bit_assignment(args.arch, model, loader, cali_data, args.options, args.avg_bits,
                                                            hadamard=args.hadamard, channel_wise=args.channel_wise, init=args.init,
                                                            batch_size=args.batch_size)
        n_cfg = len(args.options) ** table.L
        logging.info("=" * 60)
        logging.info(f"Searched {n_cfg} configurations over bits {args.options} with average bit-width <= {args.avg_bits} "
                     f"({len(table.directions())} forward jets on {world} GPU(s), {time.time() - t0:.1f} s)")
        for name, cand in candidate_dict.items():
            if all(b in args.options for b in cand):
                logging.info(f"[{name}: {cand}] The omega sensitivity score =\t{table.score(cand):.3e}")
        logging.info(f"Best Configuration: {bits}")
        logging.info(f"Average Quantization Bit-Width:\t{avg_bits:.4f}")
        logging.info(f"Minimum Score: {score:.4e}")
        logging.info("=" * 60)
        return "search", bits, score
    local = []
    for ci in candidates_of_rank(len(names), rank, world):
        bits = candidate_dict[names[ci]]
        wq_params = {"n_bits": 8, "channel_wise": args.channel_wise, "scale_method": args.init}
        qnn = QuantModel(model=copy.deepcopy(model), hadamard=args.hadamard, weight_quant_params=wq_params).to(device)
        qnn.eval()
        avg_bits = float(qnn.set_bitwidth(bits))
        qnn.set_quant_state(True)
        _ = qnn(cali_data[:args.batch_size].to(device))
        logging.info(f"[{names[ci]}: {bits}] Average Quantization Bit-Width:\t{avg_bits:.4f}")
        score = sensitivity_criterion(args.mode, args.arch, copy.deepcopy(model), qnn, loader, use_cuda=True).item()
        logging.info(f"[{names[ci]}: {bits}] The {args.mode} sensitivity score =\t{score:.3e}")
        local.append((ci, score))
    scores = gather_scores(local, len(names))
    best = int(np.argmin(scores))
    logging.info("=" * 60)
    logging.info(f"Best Candidate: {names[best]}")
    logging.info(f"Bit Configuration: {candidate_dict[names[best]]}")
    logging.info(f"Minimum Score: {scores[best]:.4e}")
    logging.info("=" * 60)
    return names[best], candidate_dict[names[best]], scores[best]


def parse_args(argv):
    p = argparse.ArgumentParser(description="running parameters", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--seed", default=903, type=int)
    p.add_argument("--outf", default="unify")
    p.add_argument("--config", type=str)
    p.add_argument("--arch", type=str)
    p.add_argument("-p", "--print-freq", default=50, type=int)
    p.add_argument("--data_path", type=str)
    p.add_argument("--vid", type=str)
    p.add_argument("--data_split", type=str, default="1_1_1")
    p.add_argument("--batch_size", default=12, type=int)
    p.add_argument("--hadamard", action="store_true")
    p.add_argument("--channel_wise", action="store_true")
    p.add_argument("--init", default="max", type=str, choices=["max", "mse", "gaussian", "l1", "l2"])
    p.add_argument("--mode", default="omega", type=str, choices=["omega", "fisher_diag"])
    p.add_argument("--ckpt", default="None", type=str)
    # additive (not in the reference): exhaustive search instead of the two hand-written candidates
    p.add_argument("--search", action="store_true", help="score every configuration over --options within --avg_bits by Omega")
    p.add_argument("--options", type=int, nargs="+", default=[2, 3, 4, 5, 6, 7, 8], help="--search: bit-widths a layer may take")
    p.add_argument("--avg_bits", type=float, default=5.0, help="--search: budget on the average bit-width")
    return p.parse_args(argv)


def seed_all(seed=903):
    random.seed(seed)
    np.random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def main(argv):
    seed_all()
    args = parse_args(argv)
    cfg = get_config(args.config)
    args.outf = os.path.join("results", args.outf)
    args.exp_id = f"{args.vid}_e{cfg['epoch']}_b{cfg['batch_size']}_lr{cfg['learning_rate']}_{cfg['loss']}"
    args.outf = os.path.join(args.outf, args.exp_id)
    assign(args, cfg)


if __name__ == "__main__":
    main(sys.argv[1:])
