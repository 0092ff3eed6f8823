"""Named decoder configurations for the CPU legs of bench.py and the tests (TEST INFRASTRUCTURE, like the rest of
oracle/): the reference's YAML files restated as dicts, and seeded random-init decoders of them.  Deliberately
independent of `neuroquant_b200` so that `bench.py --impl reference` maps no product library;
tests/test_host_logic.py checks that these agree with neuroquant_b200/workloads.py."""
from __future__ import annotations

import torch

from . import nq_oracle as O

# /root/reference/configs/HNeRV/Bunny_1280x640_3M.yaml
HNERV_BUNNY_3M = dict(crop_h=640, crop_w=1280, diff_enc=False, stage_block=1, enc_strides=[5, 4, 4, 2, 2],
                      enc_channel=[64, 64, 64, 64, 16], channel_reduce=1.2, channel_lbound=12, dec_in_channel=92,
                      dec_kernels=[1, 3, 5, 5, 5], dec_strides=[5, 4, 4, 2, 2], dec_norm="none", dec_acts="gelu",
                      out_bias="tanh", batch_size=1)
# /root/reference/configs/NeRV/Bunny_1280x640_3M.yaml
NERV_BUNNY_3M = dict(crop_h=640, crop_w=1280, diff_enc=False, base=1.25, level=80, channel_reduce=2, channel_lbound=24,
                     dec_in_channel=145, dec_kernels=[3, 3, 3, 3, 3], dec_strides=[5, 4, 4, 2, 2], dec_norm="none",
                     dec_acts="gelu", out_bias="tanh", batch_size=1)
# BASELINE.json configs[4]: not in the reference; synthesised per SURVEY 8(d) (12.02 M decoder parameters)
HNERV_1080P_12M = dict(crop_h=1080, crop_w=1920, diff_enc=False, stage_block=1, enc_strides=[5, 3, 2, 2, 2],
                       enc_channel=[64, 64, 64, 64, 16], channel_reduce=1.2, channel_lbound=12, dec_in_channel=278,
                       dec_kernels=[1, 3, 5, 5, 5], dec_strides=[5, 3, 2, 2, 2], dec_norm="none", dec_acts="gelu",
                       out_bias="tanh", batch_size=1)

WORKLOADS = {"hnerv-bunny-3m": ("hnerv", HNERV_BUNNY_3M), "nerv-bunny-3m": ("nerv", NERV_BUNNY_3M),
             "hnerv-1080p-12m": ("hnerv", HNERV_1080P_12M)}
# the reference's own YAML of each workload (relative to the reference root), where it has one
REFERENCE_YAML = {"hnerv-bunny-3m": "configs/HNeRV/Bunny_1280x640_3M.yaml", "nerv-bunny-3m": "configs/NeRV/Bunny_1280x640_3M.yaml"}


def embed_shape(cfg: dict, arch: str):
    """(C, h, w) of one decoder input (HNeRV.py:19, NeRV.py:26)."""
    import numpy as np
    if arch == "hnerv":
        s = int(np.prod(cfg["enc_strides"]))
        return cfg["enc_channel"][-1], cfg["crop_h"] // s, cfg["crop_w"] // s
    return int(cfg["level"] * 2), 1, 1


def random_stages(cfg: dict, arch: str, seed: int = 903):
    """nn.Conv2d default-initialised weights of every decoder stage, drawn in stage order under `seed` (the same draw
    as neuroquant_b200.workloads.random_decoder)."""
    geo = O.decoder_geometry(cfg, arch)
    torch.manual_seed(seed)
    stages = []
    for ci, co, k, rh, rw, act in geo:
        conv = torch.nn.Conv2d(ci, co, k, 1, k // 2)
        stages.append(O.Stage(conv.weight.detach().clone(), conv.bias.detach().clone(), rh, rw, act))
    return stages


def conv_flops(cfg: dict, arch: str, n: int = 1) -> float:
    """2*M*N*K summed over the decoder's stages, forward only."""
    _, h, w = embed_shape(cfg, arch)
    total = 0.0
    for ci, co, k, rh, rw, _ in O.decoder_geometry(cfg, arch):
        total += 2.0 * n * h * w * co * ci * k * k
        h, w = h * rh, w * rw
    return total
