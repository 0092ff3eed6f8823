from .quant_block import BaseQuantBlock  # noqa: F401
from .quant_layer import QuantModule  # noqa: F401
from .quant_model import QuantModel  # noqa: F401
from .calib_model import model_reconstruction  # noqa: F401
from .calib_block import block_reconstruction  # noqa: F401
from .calib_layer import layer_reconstruction  # noqa: F401  (the reference's, repaired: see calib_layer.py)
