"""neuroquant_b200: B200 (sm_100a) implementation of NeuroQuant's post-training-quantisation hot path.

Importing the package loads libnq_sm100.so (C ABI in include/neuroquant_b200.h); there is no CPU or
PyTorch fallback -- a missing library raises here.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is not built)
from .engine import DecoderEngine, QuantStage, StageGeom, geometry_from_cfg  # noqa: F401
from .calibration import CalibrationLoop, LinearTempDecay  # noqa: F401

__all__ = ["DecoderEngine", "QuantStage", "StageGeom", "geometry_from_cfg", "CalibrationLoop", "LinearTempDecay"]
