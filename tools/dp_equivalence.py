"""frame-sharded data parallelism against ONE GPU taking the global batch in the same frame order.
Launch under torchrun with N >= 2 GPUs (rendezvous on 127.0.0.1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_equivalence.py

Every rank runs K AdaRound iterations of HNeRV-Bunny-3M with --hadamard on its shard (batch 2 per GPU, one NCCL
all-reduce of the flat dW / db buffer per iteration); rank 0 then repeats them alone with batch 2N and compares the loss
trajectory and the rounding variables.  Prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import neuroquant_b200 as nq  # noqa: E402
from neuroquant_b200.parallel import shard_indices  # noqa: E402
from neuroquant_b200.workloads import WORKLOADS, embed_shape, random_decoder  # noqa: E402

BITS = [6, 5, 4, 5, 5, 6, 6]


def build(hadamard):
    arch, cfg = WORKLOADS["hnerv-bunny-3m"]
    geoms, params = random_decoder(cfg, arch, 903)
    stages = [nq.QuantStage(g, w.cuda(), b.cuda(), nb, hadamard) for g, (w, b), nb in zip(geoms, params, BITS)]
    eng = nq.DecoderEngine(stages)
    eng.init_scales()
    return eng, cfg, arch


def run(eng, embeds, frames, order, rank, world, group, iters, want_log=True):
    def fetch(idx):
        idx = shard_indices(torch.as_tensor(idx), rank, world).cuda()
        return embeds[idx], frames[idx]
    log = [] if want_log else None     # a log forces the eager launch sequence; without it the iteration replays as a CUDA graph
    loop = nq.CalibrationLoop(eng, fetch, len(order), iters=iters, weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003,
                              group=group, global_batch=len(order[0]), log=log)
    loop.world = world
    loop.run_phase2(lambda: order)
    if not want_log:
        assert loop._graphed and all(g.graph is not None or not g.capture for g in loop._graphed.values())
    return log


if __name__ == "__main__":
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hadamard = "--no-hadamard" not in sys.argv
    K = 12
    eng, cfg, arch = build(hadamard)
    c, h0, w0 = embed_shape(cfg, arch)
    gen = torch.Generator().manual_seed(5)
    F = 4 * world
    embeds = torch.randn(F, c, h0, w0, generator=gen).cuda()
    frames = torch.rand(F, 3, cfg["crop_h"], cfg["crop_w"], generator=gen).cuda()
    perm = torch.randperm(F, generator=gen).tolist()
    order = [perm[i:i + 2 * world] for i in range(0, F, 2 * world)]          # global batches of 2 per GPU
    iters = K // len(order) * len(order)
    log_dp = run(eng, embeds, frames, order, rank, world, None, iters)
    alpha_dp = [s.alpha_w.clone() for s in eng.stages]
    dist.barrier()
    del eng
    # the same iterations replayed as ONE CUDA graph per step with the NCCL all-reduce captured inside it: the kernels and
    # the reduction are the eager run's, so the rounding variables must come out bit-identical
    eng_g, _, _ = build(hadamard)
    run(eng_g, embeds, frames, order, rank, world, None, iters, want_log=False)
    graph_equal = all(torch.equal(a, s.alpha_w) for a, s in zip(alpha_dp, eng_g.stages))
    graph_far = max(float(((a - s.alpha_w).abs() > 1e-3).float().mean()) for a, s in zip(alpha_dp, eng_g.stages))
    dist.barrier()
    del eng_g
    dist.destroy_process_group()     # the single-GPU run below must see no process group (CalibrationLoop would all-reduce)
    if rank == 0:
        eng1, _, _ = build(hadamard)
        log_1 = run(eng1, embeds, frames, order, 0, 1, None, iters)
        rec_dp = torch.tensor([r[2] for r in log_dp]); rec_1 = torch.tensor([r[2] for r in log_1])
        far = [float(((a - s.alpha_w).abs() > 1e-3).float().mean()) for a, s in zip(alpha_dp, eng1.stages)]
        print(json.dumps({"check": "dp_equivalence", "n_gpus": world, "hadamard": hadamard, "iterations": len(log_dp),
                          "loss_rel_diff_max": float(((rec_dp - rec_1).abs() / rec_1).max()),
                          "alpha_far_fraction_max": max(far), "graph_with_nccl_bit_identical_to_eager": graph_equal,
                          "graph_alpha_far_fraction_max": graph_far, "loss_first": float(rec_1[0]), "loss_last": float(rec_1[-1])}))
