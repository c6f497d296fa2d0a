"""CPU checks around the compiled reference (oracle/_ref/qutlass_ref_C.so, oracle/build_ref.py): the library is the
unmodified reference, so the op schemas it registers are THE drop-in contract -- ours must be string-identical.  The two
register the same torch namespace (`_qutlass_C`), hence one child process each.  Skipped when the library has not been
built (it can only be built where /root/reference exists)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle as O
from oracle import ref_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_DUMP = r"""
import json, sys, torch
{setup}
names = [n for n in {names!r}]
out = {{}}
for n in names:
    try:
        out[n] = str(getattr(torch.ops._qutlass_C, n).default._schema)
    except Exception as e:
        out[n] = None
print("SCHEMAS " + json.dumps(out))
"""

OPS = ["matmul_mxf4_bf16_tn", "matmul_nvf4_bf16_tn", "matmul_ada_mxf4_bf16_tn", "matmul_mxf8_bf16_tn", "matmul_mxf8_bf16_nn",
       "fusedQuantizeMxQuest", "fusedQuantizeMxAbsMax", "fusedQuantizeNvQuest", "fusedQuantizeNvAbsMax",
       "fusedQuantizeMxQuestWithMask", "backward_t_bf16", "backward_qt_bf16", "backward_bf16_square_double_mxfp8",
       "mxfp4_transpose_mxfp8"]


def _schemas(setup: str):
    r = subprocess.run([sys.executable, "-c", _DUMP.format(setup=setup, names=OPS)], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("SCHEMAS ")]
    assert lines, r.stderr[-1000:]
    return json.loads(lines[-1][len("SCHEMAS "):])


def test_child_swizzle_matches_the_oracle():
    """oracle/ref_gpu.py feeds the reference GEMM with scales it swizzles itself (the reference's to_blocked lives in its
    Python package, which does not travel): that helper must be the pinned layout."""
    rng = np.random.default_rng(3)
    for r, c in [(128, 4), (256, 128), (384, 12), (14336 // 8, 128)]:
        a = rng.integers(0, 256, size=(r, c), dtype=np.uint8)
        got = ref_gpu._to_blocked(torch, torch.from_numpy(a).view(torch.float8_e8m0fnu)).view(torch.uint8).numpy()
        np.testing.assert_array_equal(got, O.to_blocked(a))
    assert ref_gpu._padded(200, 5) == tuple(O.padded_sf_shape(200, 5))


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref/qutlass_ref_C.so not built (python oracle/build_ref.py)")
def test_op_schemas_are_identical_to_the_compiled_reference():
    ref = _schemas(f"torch.ops.load_library({ref_gpu.LIB!r})")
    ours = _schemas("import qutlass_b200")
    assert all(v is not None for v in ref.values()), ref           # the reference build registers all 14 ops
    assert ours == ref
