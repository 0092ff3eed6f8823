"""CPU (gloo, world_size 2) tests of the multi-GPU host logic in neuroquant_b200/parallel.py: frame
sharding + global-mean normalisation + one all-reduce reproduces the single-process gradient; candidate
farming and decode sharding cover every unit exactly once.  The per-rank compute is the CPU oracle
(test infrastructure) because this container has no GPU; the partitioning code under test is the
product's."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nq_oracle as O
from tests.helpers import case_stages, t


def _oracle_flat_grad(stages, g, cali, frames, idx, global_pixels):
    """dL/dW of every stage for the frames `idx`, loss = sum_c |.|^2 summed over pixels / global_pixels."""
    ws = [s.weight.clone().requires_grad_(True) for s in stages]
    out = O.decode(stages, cali[idx], ws, [s.bias for s in stages])
    loss = (out - frames[idx]).pow(2).sum() / global_pixels
    return torch.cat([x.flatten() for x in torch.autograd.grad(loss, ws)])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neuroquant_b200.parallel import (candidates_of_rank, frame_range_of_rank, gather_scores, shard_indices,
                                          world_info)
    torch.set_num_threads(2)
    g, arch, cfg, stages = case_stages("tiny_hnerv")
    cali, frames = t(g["cali"]), t(g["frames"])
    r, w, _ = world_info()
    assert (r, w) == (rank, world)
    batch = torch.tensor([5, 2, 7, 0])  # one global mini-batch of 4 frames
    mine = shard_indices(batch, r, w)
    H, W = frames.shape[-2:]
    flat = _oracle_flat_grad(stages, g, cali, frames, mine, batch.numel() * H * W)
    dist.all_reduce(flat)  # the single collective of a calibration step
    scores = gather_scores([(i, float(i * i + 1)) for i in candidates_of_rank(5, r, w)], 5)
    lo, hi = frame_range_of_rank(7, r, w)
    q.put((rank, mine.tolist(), flat, scores, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharded_gradient_equals_single_process():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [5, 7] and res[1][1] == [2, 0]  # positions rank, rank + world
    g, arch, cfg, stages = case_stages("tiny_hnerv")
    cali, frames = t(g["cali"]), t(g["frames"])
    H, W = frames.shape[-2:]
    want = _oracle_flat_grad(stages, g, cali, frames, torch.tensor([5, 2, 7, 0]), 4 * H * W)
    for r in res:
        assert torch.allclose(r[2], want, rtol=1e-4, atol=1e-9)  # both ranks hold the full-batch gradient
        assert r[3] == [1.0, 2.0, 5.0, 10.0, 17.0]
    assert res[0][4] == (0, 4) and res[1][4] == (4, 7)


def test_partition_helpers_cover_everything_once():
    from neuroquant_b200.parallel import candidates_of_rank, frame_range_of_rank, shard_indices
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in candidates_of_rank(11, r, world))
        assert seen == list(range(11))
        ranges = [frame_range_of_rank(132, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 132 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        idx = torch.arange(2 * world)
        parts = [shard_indices(idx, r, world).tolist() for r in range(world)]
        assert sorted(x for p in parts for x in p) == idx.tolist() and all(len(p) == 2 for p in parts)
    with pytest.raises(ValueError):
        shard_indices(torch.arange(3), 0, 2)
