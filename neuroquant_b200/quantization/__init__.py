from .quant_block import BaseQuantBlock  # noqa: F401
from .quant_layer import QuantModule  # noqa: F401
from .quant_model import QuantModel  # noqa: F401
from .calib_model import model_reconstruction  # noqa: F401
from .calib_block import block_reconstruction  # noqa: F401
# layer_reconstruction (reference calib_layer.py) is imported by the reference package but called by no command line and
# crashes there (calib_layer.py:130); it is not provided.
