"""QuantNeRVBlock (reference: quantization/quant_block.py)."""
import torch.nn as nn

from ..models._layers import NeRVBlock
from .quant_layer import QuantModule


class BaseQuantBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.use_weight_quant = False
        self.trained = False
        self.ignore_reconstruction = False

    def set_quant_state(self, weight_quant: bool = False):
        self.use_weight_quant = weight_quant
        for m in self.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant)


class QuantNeRVBlock(BaseQuantBlock):
    """conv (QuantModule) -> PixelShuffle -> activation; the reference drops the norm layer here
    (quant_block.py:27-29), which is only valid for `dec_norm: none` -- asserted."""

    def __init__(self, basic_block: NeRVBlock, hadamard: bool = True, weight_quant_params: dict = {}):
        super().__init__()
        if not isinstance(basic_block.norm, nn.Identity):
            raise ValueError("QuantNeRVBlock supports dec_norm: none only (the reference silently drops the norm)")
        self.conv = QuantModule(basic_block.conv[0], hadamard, weight_quant_params)
        self.pixelshuffle = basic_block.conv[1]
        self.act = basic_block.act

    def forward(self, x):
        """Stand-alone use; inside a model the block is one fused engine stage."""
        return self.act(self.pixelshuffle(self.conv(x)))


specials = {
    NeRVBlock: QuantNeRVBlock,
}
