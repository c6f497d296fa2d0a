"""Alias of qutlass_b200.utils (same names as the reference's qutlass/utils.py)."""
from qutlass_b200.utils import (ceil_div, get_padded_shape_mx, get_padded_shape_nv, pad_to_block,  # noqa: F401
                                to_blocked)
