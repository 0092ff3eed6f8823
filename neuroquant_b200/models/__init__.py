from .HNeRV import HNeRV  # noqa: F401
from .NeRV import NeRV  # noqa: F401
# PNeRV1 / PNeRV2 (reference models/PNeRV.py) are not supported by QuantModel (quant_model.py:13) nor by
# the calibrate / bit_assign command lines: out of scope.
