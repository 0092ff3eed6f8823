from .quant_block import BaseQuantBlock  # noqa: F401
from .quant_layer import QuantModule  # noqa: F401
from .quant_model import QuantModel  # noqa: F401
from .calib_model import model_reconstruction  # noqa: F401
# layer_reconstruction / block_reconstruction (reference calib_layer.py, calib_block.py) are imported by the
# reference package but called by no command line; they are the first "next" row of SURVEY 8(f).
