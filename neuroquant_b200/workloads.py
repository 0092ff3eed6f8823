"""Named decoder configurations (the reference's YAML files restated as dicts) and synthetic,
random-init instances of them for benchmarks and smoke tests (no checkpoints ship with the
reference: readme.md:130)."""
from __future__ import annotations

from typing import List, Tuple

import torch

from .engine import StageGeom, geometry_from_cfg

# configs/HNeRV/Bunny_1280x640_3M.yaml
HNERV_BUNNY_3M = dict(crop_h=640, crop_w=1280, diff_enc=False, stage_block=1, enc_strides=[5, 4, 4, 2, 2],
                      enc_channel=[64, 64, 64, 64, 16], channel_reduce=1.2, channel_lbound=12, dec_in_channel=92,
                      dec_kernels=[1, 3, 5, 5, 5], dec_strides=[5, 4, 4, 2, 2], dec_norm="none", dec_acts="gelu",
                      out_bias="tanh", batch_size=1)
# configs/NeRV/Bunny_1280x640_3M.yaml
NERV_BUNNY_3M = dict(crop_h=640, crop_w=1280, diff_enc=False, base=1.25, level=80, channel_reduce=2, channel_lbound=24,
                     dec_in_channel=145, dec_kernels=[3, 3, 3, 3, 3], dec_strides=[5, 4, 4, 2, 2], dec_norm="none",
                     dec_acts="gelu", out_bias="tanh", batch_size=1)
# BASELINE.json configs[4]: not in the reference; synthesised per SURVEY 8(d) (12.02 M decoder parameters)
HNERV_1080P_12M = dict(crop_h=1080, crop_w=1920, diff_enc=False, stage_block=1, enc_strides=[5, 3, 2, 2, 2],
                       enc_channel=[64, 64, 64, 64, 16], channel_reduce=1.2, channel_lbound=12, dec_in_channel=278,
                       dec_kernels=[1, 3, 5, 5, 5], dec_strides=[5, 3, 2, 2, 2], dec_norm="none", dec_acts="gelu",
                       out_bias="tanh", batch_size=1)

WORKLOADS = {
    "hnerv-bunny-3m": ("hnerv", HNERV_BUNNY_3M),
    "nerv-bunny-3m": ("nerv", NERV_BUNNY_3M),
    "hnerv-1080p-12m": ("hnerv", HNERV_1080P_12M),
}


def embed_shape(cfg: dict, arch: str) -> Tuple[int, int, int]:
    """(C, h, w) of one decoder input (HNeRV.py:19, NeRV.py:26)."""
    import numpy as np

    if arch == "hnerv":
        s = int(np.prod(cfg["enc_strides"]))
        return cfg["enc_channel"][-1], cfg["crop_h"] // s, cfg["crop_w"] // s
    return int(cfg["level"] * 2), 1, 1


def random_decoder(cfg: dict, arch: str, seed: int = 903) -> Tuple[List[StageGeom], List[Tuple[torch.Tensor, torch.Tensor]]]:
    """nn.Conv2d default-initialised weights of every decoder stage (CPU tensors)."""
    geoms = geometry_from_cfg(cfg, arch)
    torch.manual_seed(seed)
    params = []
    for g in geoms:
        conv = torch.nn.Conv2d(g.cin, g.cout, g.k, 1, g.k // 2)
        params.append((conv.weight.detach().clone(), conv.bias.detach().clone()))
    return geoms, params


def conv_flops(geoms: List[StageGeom], h0: int, w0: int, n: int = 1) -> float:
    """2*M*N*K summed over stages (forward); dgrad and wgrad cost the same each."""
    h, w, total = h0, w0, 0.0
    for g in geoms:
        total += 2.0 * n * h * w * g.cout * g.cin * g.k * g.k
        h, w = h * g.rh, w * g.rw
    return total
