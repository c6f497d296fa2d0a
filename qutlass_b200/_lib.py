"""ctypes loader for libb200q.so -- the C-ABI declared in include/b200q.h.

There is NO CPU fallback and no alternative backend: if the library is missing the import of any
compute entry point fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200q.so")
# B200Q_LIB=prof selects the profiling build (python -m qutlass_b200.build --profiling): probe tools only
if os.environ.get("B200Q_LIB"):
    LIB_PATH = os.path.join(_HERE, "lib", "libb200q_%s.so" % os.environ["B200Q_LIB"])

# name -> (restype, argtypes); must match include/b200q.h exactly (tests/test_cabi.py checks the header)
_vp, _i64, _i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
SIGNATURES = {
    "b200q_abi_version": (_i32, []),
    "b200q_last_error": (ctypes.c_char_p, []),
    "b200q_reload_env": (None, []),
    "b200q_profiling_build": (_i32, []),
    "b200q_sf_write_generation": (ctypes.c_uint, [_vp]),
    "b200q_debug_tmap_cache_stats": (_i32, [ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]),
    "b200q_quantize_mx": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "b200q_quantize_nv": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "b200q_swizzle_sf": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "b200q_gemm_fp4": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "b200q_gemm_fp4_cfg": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200q_gemm_fp4_plan": (_i32, [_i32, _i32, _i32, _i32, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "b200q_gemm_fp4_launches": (_i32, [_i32, _i32, _i32, _i32]),
    "b200q_linear_fp4_workspace_bytes": (_i64, [_i32]),
    "b200q_linear_fp4": (_i32, [_vp] * 11 + [_i32] * 6 + [_vp]),
    "b200q_linear_fp4_launches": (_i32, [_i32] * 6),
    "b200q_linear_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "b200q_linear_host_slabs": (_i32, [_i32, ctypes.POINTER(ctypes.c_int), _i32]),
    "b200q_backward_t_bf16": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "b200q_backward_qt_bf16": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "b200q_backward_bf16_square_double_mxfp8": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200q_mxfp4_transpose_mxfp8": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "b200q_linear_fp4_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
}

_lib = None


class B200QError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m qutlass_b200.build` "
                "(there is no CPU or library fallback for this path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def reload_env() -> None:
    """Re-read the library's environment switches (they are cached at first use; include/b200q.h: b200q_reload_env)."""
    load().b200q_reload_env()


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200q_last_error().decode(errors="replace")
        raise B200QError(msg or f"libb200q call failed with code {rc}")
