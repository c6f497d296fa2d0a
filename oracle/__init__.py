"""CPU oracle for the microscaled-FP4 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the
product package ``qutlass_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and
there only as the checker / the reported CPU baseline, never as the thing
shipped.

Parity status: PINNED.  ``oracle/fp4_oracle.py`` is checked against golden
vectors generated in the build container by running the reference's own test
helpers (``/root/reference/tests/mxfp4_test.py`` and ``nvfp4_test.py``:
``_rtne_fp4``, ``_dq_fp4``, ``_forward_quantize_ref``) and ``qutlass/utils.py``
``to_blocked`` -- see ``tests/golden/make_golden.py`` (the generator, committed)
and ``tests/test_oracle_golden.py`` (the check).

``oracle/_ref/`` (git-ignored, travels to the GPU box) holds the reference ITSELF compiled for sm_100a from the sources
where they lie (``oracle/build_ref.py``); ``oracle/ref_gpu.py`` runs it in a child process as the GPU-side checker and
as the kernel to beat (``tests/test_gpu_reference_lib.py``, ``bench.py``'s ``reference_gpu`` leg).
"""
from .fp4_oracle import *  # noqa: F401,F403
from . import bwd_oracle  # noqa: F401  (backward re-quantisers: oracle.bwd_oracle.*)
