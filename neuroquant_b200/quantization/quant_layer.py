"""QuantModule (reference: quantization/quant_layer.py): one nn.Conv2d with fake-quantised weight and bias
and the optional Walsh-Hadamard rotation of the weight's input-channel axis."""
import math
from typing import Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from .quantizer import StraightThrough, UniformAffineQuantizer  # noqa: F401


def _next_power_of_two(n: int):
    return 1 if n == 0 else 2 ** math.ceil(math.log2(n))


def hadamard_along_channel_weight(x: torch.Tensor, normalize: bool = True):
    """Orthonormal WHT over C_in of a (C_out, C_in, KH, KW) tensor (quant_layer.py:16-22): one
    warp-shuffle kernel (nq_fwht), self-inverse."""
    return L.fwht_channel(x.detach().contiguous().float())


def _same_conv_or_raise(conv: nn.Conv2d):
    """The engine's kernels cover what the decoders use: odd square kernels, stride 1, 'same' padding, no groups."""
    k = conv.kernel_size
    ok = (conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1 and k[0] == k[1] and k[0] % 2 == 1
          and conv.padding == (k[0] // 2, k[0] // 2))
    if not ok:
        raise ValueError("QuantModule kernels cover the decoders' stride-1 'same' convolutions only: {}".format(conv))


class QuantModule(nn.Module):
    """Attribute names are the reference's (quant_layer.py:29-65): pickled QuantModels carry them and the runner reads
    them -- weight / org_weight / bias / org_bias, hadamard + C + hadamard_weight, use_weight_quant, the two quantisers,
    fwd_kwargs / fwd_func."""

    def __init__(self, org_module: Union[nn.Conv2d,], hadamard: bool = True, weight_quant_params: dict = {}):
        super().__init__()
        if not isinstance(org_module, nn.Conv2d):
            raise ValueError("Not supported modules: {}".format(org_module))
        _same_conv_or_raise(org_module)
        self.fwd_kwargs = {k: getattr(org_module, k) for k in ("stride", "padding", "dilation", "groups")}
        self.fwd_func = F.conv2d  # kept for layout compatibility; the forward below runs libnq_sm100 kernels
        # the learnable tensors stay the module's own Parameters; full-precision copies serve the un-quantised state
        self.weight, self.org_weight = org_module.weight, org_module.weight.data.clone()
        has_bias = org_module.bias is not None
        self.bias = org_module.bias if has_bias else None
        self.org_bias = org_module.bias.data.clone() if has_bias else None
        self.hadamard = hadamard
        if hadamard:
            # rotated copy of the weight: input channels zero-padded to a power of two, then the orthonormal WHT
            self.C = self.weight.shape[1]
            grow = _next_power_of_two(self.C) - self.C
            self.hadamard_weight = hadamard_along_channel_weight(F.pad(self.org_weight, (0, 0, 0, 0, 0, grow)))
        self.use_weight_quant = False
        self.weight_quantizer, self.bias_quantizer = (UniformAffineQuantizer(**weight_quant_params) for _ in range(2))
        self.extra_repr = org_module.extra_repr

    def quantized_weight_bias(self):
        """(weight, bias) the convolution sees (quant_layer.py:68-78)."""
        if self.use_weight_quant:
            if self.hadamard:
                weight = hadamard_along_channel_weight(self.weight_quantizer(self.hadamard_weight))[:, :self.C, :, :]
            else:
                weight = self.weight_quantizer(self.weight)
            bias = self.bias_quantizer(self.bias)
        else:
            weight, bias = self.org_weight, self.org_bias
        return weight, bias

    def forward(self, input: torch.Tensor):
        """Stand-alone use of one layer (NCHW in, NCHW out, no autograd graph).  Inside a QuantModel the
        whole decoder runs fused on the engine instead (models/HNeRV.py decode)."""
        from ..runner import conv2d_nchw
        weight, bias = self.quantized_weight_bias()
        return conv2d_nchw(input, weight.detach().contiguous(), None if bias is None else bias.detach().contiguous())

    def set_quant_state(self, weight_quant: bool = False):
        self.use_weight_quant = weight_quant

    def get_weight_perturbation(self):
        """quant_layer.py:86-89: org_weight - UAQ(weight), never rotated."""
        return self.org_weight - self.weight_quantizer(self.weight).detach()
