"""Shared helpers for the parity tests: load golden fixtures into oracle structures."""
import os

import numpy as np
import torch

from oracle import nq_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# the tiny configs tests/golden/make_golden.py ran the reference on
TINY_HNERV = dict(crop_h=32, crop_w=64, diff_enc=False, stage_block=1,
                  enc_strides=[2, 2, 2, 2, 1], enc_channel=[8, 8, 8, 8, 4],
                  channel_reduce=1.2, channel_lbound=6, dec_in_channel=20,
                  dec_kernels=[1, 3, 5, 5, 3], dec_strides=[2, 2, 2, 2, 1],
                  dec_norm="none", dec_acts="gelu", out_bias="tanh")
TINY_NERV = dict(crop_h=32, crop_w=64, diff_enc=False, base=1.25, level=10,
                 channel_reduce=2, channel_lbound=6, dec_in_channel=18,
                 dec_kernels=[3, 3, 3, 3, 3], dec_strides=[2, 2, 2, 2, 1],
                 dec_norm="none", dec_acts="gelu", out_bias="tanh")
CASES = {
    "tiny_hnerv": ("hnerv", TINY_HNERV),
    "tiny_hnerv_had": ("hnerv", TINY_HNERV),
    "tiny_nerv": ("nerv", TINY_NERV),
    "tiny_nerv_had": ("nerv", TINY_NERV),
}


# block-wise reconstruction fixtures (tests/golden/make_block_golden.py)
# per-tensor scales (the command line without --channel_wise); fixture from `make_golden.py layerwise`
LW_CASES = {"tiny_hnerv_lw": ("hnerv", TINY_HNERV)}


def cw(g) -> bool:
    """channel_wise flag of a model fixture (older fixtures predate the key: per-channel)."""
    return bool(g["channel_wise"]) if "channel_wise" in g.files else True


BLOCK_CASES = {
    "block_tiny_hnerv": ("hnerv", TINY_HNERV),
    "block_tiny_hnerv_qdrop": ("hnerv", TINY_HNERV),
    "block_tiny_nerv": ("nerv", TINY_NERV),
    "block_tiny_hnerv_fdiag": ("hnerv", TINY_HNERV),   # opt_mode 'fisher_diag'
    "block_tiny_hnerv_ffull": ("hnerv", TINY_HNERV),   # opt_mode 'fisher_full', asym, QDrop 0.5
}
# layer_reconstruction (calib_layer.py:89-179), the reference's own source with its missing `opt_params = []` inserted in memory
LAYER_CASES = {
    "layer_tiny_hnerv_conv": ("hnerv", TINY_HNERV),    # a block's convolution, asym, QDrop 0.5
    "layer_tiny_hnerv_head": ("hnerv", TINY_HNERV),    # head layer, fisher_diag
    "layer_tiny_nerv_stem": ("nerv", TINY_NERV),       # stem (1x1), input = the embedding
}


def block_case(tag):
    arch, cfg = (BLOCK_CASES[tag] if tag in BLOCK_CASES else LAYER_CASES[tag])
    g = load(tag)
    sd = {k[3:]: t(g[k]) for k in g.files if k.startswith("sd/")}
    return g, arch, cfg, O.stages_from_state_dict(sd, cfg, arch)


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def t(a):
    return torch.from_numpy(np.asarray(a))


def case_stages(tag):
    arch, cfg = CASES[tag] if tag in CASES else LW_CASES[tag]
    g = load(tag)
    sd = {k[3:]: t(g[k]) for k in g.files if k.startswith("sd/")}
    return g, arch, cfg, O.stages_from_state_dict(sd, cfg, arch)
