"""QuantModel (reference: quantization/quant_model.py): module surgery, global quantisation switch,
per-layer bit-widths, accessors for codes and perturbations."""
from typing import Union

import torch.nn as nn

from ..models import HNeRV, NeRV
from .quant_block import BaseQuantBlock, specials
from .quant_layer import QuantModule, StraightThrough


class QuantModel(nn.Module):
    def __init__(self, model: Union[NeRV, HNeRV], hadamard: bool = True, weight_quant_params: dict = {}):
        super().__init__()
        self.model = model
        self.hadamard = hadamard
        self.quant_module_refactor(self.model, weight_quant_params)

    def _wrapped(self, child: nn.Module, params: dict):
        """The quantised stand-in of one child, or None when the child is kept and searched instead."""
        wrapper = specials.get(type(child))
        if wrapper is not None:
            return wrapper(child, self.hadamard, params)
        if isinstance(child, nn.Conv2d):
            return QuantModule(child, self.hadamard, params)
        return None

    def quant_module_refactor(self, module: nn.Module, weight_quant_params: dict = {}):
        """Module surgery of quant_model.py:19-41: NeRVBlock -> QuantNeRVBlock, bare Conv2d -> QuantModule, depth first;
        children whose name contains 'encoder' (never quantised) and StraightThrough placeholders are left alone."""
        for name, child in list(module.named_children()):
            if "encoder" in name or isinstance(child, StraightThrough):
                continue
            repl = self._wrapped(child, weight_quant_params)
            if repl is None:
                self.quant_module_refactor(child, weight_quant_params)
            else:
                setattr(module, name, repl)

    def quant_modules(self):
        return [m for m in self.model.modules() if isinstance(m, QuantModule)]

    def set_quant_state(self, weight_quant: bool = False):
        for m in self.model.modules():
            if isinstance(m, (QuantModule, BaseQuantBlock)):
                m.set_quant_state(weight_quant)

    def encode(self, input):
        return self.model.encode(input)

    def decode(self, input):
        return self.model.decode(input)

    def forward(self, input):
        return self.model.decode(input)

    def set_bitwidth(self, bit, init=False):
        """Give layer i (weight and bias quantiser alike) bit[i] bits and mark the scales (un)initialised; returns the
        average bits per decoder parameter (quant_model.py:58-72).  Integer sums: exact, so the ratio is the reference's."""
        total_bits = total_params = 0
        for i, m in enumerate(self.quant_modules()):
            for q in (m.weight_quantizer, m.bias_quantizer):
                q.bitwidth_refactor(bit[i])
                q.inited = init
            n_w, n_b = m.weight.numel(), m.bias.numel()
            total_bits += m.weight_quantizer.n_bits * n_w + m.bias_quantizer.n_bits * n_b
            total_params += n_w + n_b
        return float(total_bits) / float(total_params)

    def get_quantized_param(self):
        """Codes cached by the last forward, W and b interleaved per layer (quant_model.py:74-80)."""
        out = []
        for m in self.quant_modules():
            out += [m.weight_quantizer.x_quant, m.bias_quantizer.x_quant]
        return out

    def get_perturbation(self):
        return [m.get_weight_perturbation() for m in self.quant_modules()]
