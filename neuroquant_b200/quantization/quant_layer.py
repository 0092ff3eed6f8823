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


class QuantModule(nn.Module):
    def __init__(self, org_module: Union[nn.Conv2d,], hadamard: bool = True, weight_quant_params: dict = {}):
        super().__init__()
        if isinstance(org_module, nn.Conv2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding, dilation=org_module.dilation,
                                   groups=org_module.groups)
            self.fwd_func = F.conv2d  # kept for layout compatibility; the forward below runs libnq_sm100 kernels
        else:
            raise ValueError("Not supported modules: {}".format(org_module))
        k = org_module.kernel_size
        if org_module.stride != (1, 1) or org_module.dilation != (1, 1) or org_module.groups != 1 or k[0] != k[1] or \
                k[0] % 2 == 0 or org_module.padding != (k[0] // 2, k[0] // 2):
            raise ValueError("QuantModule kernels cover the decoders' stride-1 'same' convolutions only: {}".format(org_module))
        self.weight = org_module.weight
        self.org_weight = org_module.weight.data.clone()
        self.hadamard = hadamard
        if self.hadamard:
            C_out, C_in, KH, KW = self.weight.shape
            self.C = C_in
            pad_channels = _next_power_of_two(self.C) - self.C
            x_padded = F.pad(org_module.weight.data.clone(), (0, 0, 0, 0, 0, pad_channels))
            self.hadamard_weight = hadamard_along_channel_weight(x_padded)
        if org_module.bias is not None:
            self.bias = org_module.bias
            self.org_bias = org_module.bias.data.clone()
        else:
            self.bias = None
            self.org_bias = None
        self.use_weight_quant = False
        self.weight_quantizer = UniformAffineQuantizer(**weight_quant_params)
        self.bias_quantizer = UniformAffineQuantizer(**weight_quant_params)
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
