"""Checkpoint-layout compatibility with the reference.

The reference saves the calibrated model as a whole-object pickle (`torch.save(qnn, ...)`,
calibrate_network.py:305-308) whose GLOBALs are `quantization.quant_model.QuantModel`,
`quantization.quant_layer.QuantModule`, `quantization.quant_block.QuantNeRVBlock`,
`quantization.quantizer.AdaRoundQuantizer`, `models.HNeRV.HNeRV`, `models._layers.*`, ...
install_reference_aliases() registers this package's modules under those top-level names and stamps the
classes' __module__ accordingly, so that checkpoints written here carry the reference's paths and
checkpoints written by the reference unpickle into these classes (the attribute names already match).
"""
import sys


def install_reference_aliases():
    from . import models, quantization, utils, videosets
    from .models import _layers
    from .quantization import calib_block, calib_layer, calib_model, data_utils, quant_block, quant_layer, quant_model, quantizer

    table = {
        "models": models, "models.HNeRV": sys.modules[models.__name__ + ".HNeRV"],
        "models.NeRV": sys.modules[models.__name__ + ".NeRV"], "models._layers": _layers,
        "quantization": quantization, "quantization.quantizer": quantizer, "quantization.quant_layer": quant_layer,
        "quantization.quant_block": quant_block, "quantization.quant_model": quant_model,
        "quantization.calib_model": calib_model, "quantization.data_utils": data_utils,
        "quantization.calib_block": calib_block, "quantization.calib_layer": calib_layer,
        "utils": utils, "videosets": videosets,
    }
    for name, mod in table.items():
        if name in sys.modules and sys.modules[name] is not mod:
            if not getattr(sys.modules[name], "__file__", "").startswith(__file__.rsplit("/", 1)[0]):
                raise ImportError(f"module {name!r} is already imported from elsewhere; cannot alias the reference layout")
        sys.modules[name] = mod
    for name, mod in table.items():
        if "." not in name:
            continue
        for obj in vars(mod).values():
            if isinstance(obj, type) and obj.__module__ == mod.__name__:
                obj.__module__ = name
