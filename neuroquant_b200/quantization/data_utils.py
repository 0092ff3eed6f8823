"""Schedules of the calibration loop (reference: quantization/data_utils.py:24-41).  The layer/block
input-output caching hooks of that file serve only the block-/layer-wise variants (calib_block.py,
calib_layer.py), which no command line invokes: SURVEY 8(f), not built yet."""
from ..calibration import LinearTempDecay  # noqa: F401
