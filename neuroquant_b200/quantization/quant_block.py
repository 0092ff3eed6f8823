"""Quantised decoder blocks (reference: quantization/quant_block.py).

A block here is a thin handle: inside a model the convolution, the PixelShuffle and the activation of a block run as ONE
fused engine stage (runner.DecoderRunner); these classes carry the reference's attribute names -- pickled QuantModels
hold them -- and the switch that turns the block's quantisers on and off.
"""
import torch.nn as nn

from ..models._layers import NeRVBlock
from .quant_layer import QuantModule

_FLAGS = ("use_weight_quant", "trained", "ignore_reconstruction")   # quant_block.py:11-13


class BaseQuantBlock(nn.Module):
    """Anything that owns QuantModules and switches them together."""

    def __init__(self):
        super().__init__()
        for flag in _FLAGS:
            setattr(self, flag, False)

    def quant_layers(self):
        return [m for m in self.modules() if isinstance(m, QuantModule)]

    def set_quant_state(self, weight_quant: bool = False):
        self.use_weight_quant = weight_quant
        for layer in self.quant_layers():
            layer.set_quant_state(weight_quant)


class QuantNeRVBlock(BaseQuantBlock):
    """conv (QuantModule) -> PixelShuffle -> activation.  The reference drops the block's norm layer without a word
    (quant_block.py:27-29); that is only right for `dec_norm: none`, so anything else is refused."""

    def __init__(self, basic_block: NeRVBlock, hadamard: bool = True, weight_quant_params: dict = {}):
        super().__init__()
        if not isinstance(basic_block.norm, nn.Identity):
            raise ValueError("QuantNeRVBlock supports dec_norm: none only (the reference silently drops the norm)")
        conv, shuffle = basic_block.conv[0], basic_block.conv[1]
        self.conv = QuantModule(conv, hadamard, weight_quant_params)
        self.pixelshuffle, self.act = shuffle, basic_block.act

    def forward(self, x):
        """Stand-alone use; inside a model the block is one fused engine stage."""
        return self.act(self.pixelshuffle(self.conv(x)))


# which wrapper replaces which full-precision block (quant_model.py looks types up here)
specials = {NeRVBlock: QuantNeRVBlock}
