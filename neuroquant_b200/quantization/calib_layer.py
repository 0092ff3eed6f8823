"""Layer-wise calibration (reference: quantization/calib_layer.py:89-179).

As shipped the reference's layer_reconstruction stops at calib_layer.py:130 (`opt_params += ...` before any assignment).
With that one statement supplied it is block_reconstruction on a lone QuantModule, and that is what this provides, on the
same GPU step (quantization/calib_block.py): the compared output is the convolution's own -- before the up-shuffle and the
activation -- the cached input of the stem is the embedding itself, and the rounding regulariser is never applied
(LossFunction.collect_round_loss, calib_layer.py:38-46, walks the children of the module it is given; a QuantModule's
children are its quantisers).  Pinned against the reference's own source with the missing statement inserted in memory
(tests/golden/make_block_golden.py: layer_tiny_*.npz).
"""
from __future__ import annotations

import torch

from .calib_block import _reconstruct
from .quant_layer import QuantModule


def layer_reconstruction(model, layer: QuantModule, cali_data: torch.Tensor, batch_size: int = 8, iters: int = 20000,
                         weight: float = 0.01, opt_mode: str = "mse", asym: bool = False, b_range: tuple = (20, 2),
                         warmup: float = 0.0, input_prob: float = 1.0, p: float = 2.0, lr: float = 0.0015):
    """Same arguments as the reference.  `layer`: any QuantModule of the model's decoder (stem, a block's convolution,
    head)."""
    if not isinstance(layer, QuantModule):
        raise ValueError("layer_reconstruction expects a QuantModule of the model's decoder")
    return _reconstruct(model, layer, layer, True, cali_data, batch_size, iters, weight, opt_mode, asym, b_range, warmup,
                        input_prob, p, lr)
