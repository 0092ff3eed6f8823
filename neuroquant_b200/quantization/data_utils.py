"""The reference's `quantization/data_utils.py` surface: `LinearTempDecay` (:24-41) and the per-layer / per-block input,
output and output-gradient caches of the block- and layer-wise calibration (:45-272) -- `save_inp_oup_data`,
`save_grad_data`, `GetLayerInpOut`, `GetLayerGrad`, `quantize_model_till` -- with the reference's names, argument order
and return shapes.

The reference obtains these tensors with forward / backward hooks on PyTorch modules.  Here the decoder runs as fused
kernels (neuroquant_b200/engine.py), where no module hook would fire: the same tensors are read off the engine's own
stage buffers (`DecoderRunner.features`, `DecoderEngine.stage_input_grad / stage_output_grad`); the implementation is
shared with quantization/calib_block.py.  The hook classes themselves (`DataSaverHook`, `GradSaverHook`) are kept as
plain callables for code that registers them on stand-alone `QuantModule`s.
"""
from __future__ import annotations

from typing import Union

import torch

from ..calibration import LinearTempDecay  # noqa: F401
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule


def _locate(model, layer):
    """(runner, stage index, layer_mode) of a QuantModule / QuantNeRVBlock of `model`'s decoder."""
    from ..runner import DecoderRunner
    runner = DecoderRunner.of(model.model)
    if isinstance(layer, BaseQuantBlock):
        convs = [m for m in layer.modules() if isinstance(m, QuantModule)]
        if len(convs) != 1:
            raise NotImplementedError("blocks with other than one quantised convolution")
        conv, layer_mode = convs[0], False
    elif isinstance(layer, QuantModule):
        conv, layer_mode = layer, True
    else:
        raise ValueError("expected a QuantModule or a QuantNeRVBlock of the model's decoder")
    k = [i for i, l in enumerate(runner.layers) if l is conv]
    if not k:
        raise ValueError("layer is not part of this model's decoder")
    return runner, k[0], conv, layer_mode


class StopForwardException(Exception):
    """data_utils.py:122-126."""


class DataSaverHook:
    """data_utils.py:129-147: forward hook storing a module's input / output."""

    def __init__(self, store_input=False, store_output=False, stop_forward=False):
        self.store_input, self.store_output, self.stop_forward = store_input, store_output, stop_forward
        self.input_store = self.output_store = None

    def __call__(self, module, input_batch, output_batch):
        if self.store_input:
            self.input_store = input_batch
        if self.store_output:
            self.output_store = output_batch
        if self.stop_forward:
            raise StopForwardException


class GradSaverHook:
    """data_utils.py:209-219: backward hook storing a module's output gradient."""

    def __init__(self, store_grad=True):
        self.store_grad, self.stop_backward, self.grad_out = store_grad, False, None

    def __call__(self, module, grad_input, grad_output):
        if self.store_grad:
            self.grad_out = grad_output[0]
        if self.stop_backward:
            raise StopForwardException


class GetLayerInpOut:
    """data_utils.py:149-205: `__call__(model_input)` returns (input the optimisation sees, full-precision output[,
    full-precision input]) of `layer` for one batch of decoder inputs.  With `asym` the input comes from a pass with the
    whole network quantised.  Leaves the model with only `layer` quantised, in train mode, like the reference."""

    def __init__(self, model, layer: Union[QuantModule, BaseQuantBlock], device: torch.device, asym: bool = False,
                 input_prob: bool = False):
        self.model, self.layer, self.device, self.asym, self.input_prob = model, layer, device, asym, input_prob

    def __call__(self, model_input):
        from . import calib_block as cb
        runner, k, conv, layer_mode = _locate(self.model, self.layer)
        self.model.eval()
        x = model_input.to(self.device)
        (inp, sym), out = cb.save_inp_oup_data(self.model, runner, k, x, self.asym, batch_size=x.size(0),
                                               layer=conv if layer_mode else None)
        self.model.set_quant_state(False)
        self.layer.set_quant_state(True)
        self.model.train()
        return (inp, out, sym) if self.input_prob else (inp, out)


class GetLayerGrad:
    """data_utils.py:222-258: raw gradient of mean((out_fp - out_q)^2), the decoder quantised up to and including
    `layer`, with respect to `layer`'s output."""

    def __init__(self, model, layer: Union[QuantModule, BaseQuantBlock], device: torch.device):
        self.model, self.layer, self.device = model, layer, device

    def __call__(self, model_input):
        from . import calib_block as cb
        runner, k, conv, layer_mode = _locate(self.model, self.layer)
        self.model.eval()
        g = cb.block_output_grads(self.model, runner, self.layer, k, model_input.to(self.device), layer_mode)
        self.model.train()
        return g


def save_inp_oup_data(model, layer: Union[QuantModule, BaseQuantBlock], cali_data: torch.Tensor, asym: bool = False,
                      batch_size: int = 8, keep_gpu: bool = True, input_prob: bool = False):
    """data_utils.py:45-88.  Returns ((inputs,), outputs), or ((inputs, full-precision inputs), outputs) with
    `input_prob`; the last cali_data.size(0) % batch_size samples are dropped as in the reference (:67)."""
    from . import calib_block as cb
    runner, k, conv, layer_mode = _locate(model, layer)
    (inps, syms), outs = cb.save_inp_oup_data(model, runner, k, cali_data, asym, batch_size, layer=conv if layer_mode else None)
    if not keep_gpu:
        inps, syms, outs = inps.cpu(), syms.cpu(), outs.cpu()
    return ((inps, syms), outs) if input_prob else ((inps,), outs)


def save_grad_data(model, layer: Union[QuantModule, BaseQuantBlock], cali_data: torch.Tensor, batch_size: int = 8,
                   keep_gpu: bool = True):
    """data_utils.py:91-119: |g| + 1 over the calibration set (the per-sample loss is a mean over the sample, so the
    batch size does not change any sample's gradient; samples are processed one at a time)."""
    from . import calib_block as cb
    runner, k, conv, layer_mode = _locate(model, layer)
    n = int(cali_data.size(0) / batch_size) * batch_size
    g = cb.save_grad_data(model, runner, layer, k, cali_data[:n], layer_mode)
    return g if keep_gpu else g.cpu()


def quantize_model_till(model, layer: Union[QuantModule, BaseQuantBlock]):
    """data_utils.py:261-272: quantise every layer / block up to and including `layer`, in module order."""
    model.set_quant_state(False)
    for _, module in model.named_modules():
        if isinstance(module, (QuantModule, BaseQuantBlock)):
            module.set_quant_state(True)
        if module is layer:
            break
