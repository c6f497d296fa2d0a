"""Drop-in alias: ``import qutlass`` resolves to the B200-native implementation in ``qutlass_b200`` so the
reference's own tests/ and benchmarks/ (which do ``from qutlass import ...``) run unmodified against it."""
from qutlass_b200 import *  # noqa: F401,F403
from qutlass_b200 import (fusedQuantizeMx, fusedQuantizeNv, matmul_mxf4_bf16_tn, matmul_nvf4_bf16_tn,  # noqa: F401
                          matmul_ada_mxf4_bf16_tn, matmul_mxf8_bf16_tn, matmul_mxf8_bf16_nn, backward_t_bf16,
                          backward_qt_bf16, backward_bf16_square_double_mxfp8, mxfp4_transpose_mxfp8)
from . import utils  # noqa: F401
