"""qutlass_b200 -- B200-native (sm_100a) microscaled-FP4 hot path behind the QuTLASS Python surface.

Drop-in for the hot-path subset of the reference's ``qutlass`` package
(/root/reference/qutlass/__init__.py:34-203):

    fusedQuantizeMx, fusedQuantizeNv, matmul_mxf4_bf16_tn, matmul_nvf4_bf16_tn, utils.to_blocked

plus, from the "next" rows of the scope table, the MXFP8 GEMMs (matmul_mxf8_bf16_tn / _nn) and the transposing
re-quantisers of the QAT backward pass (backward_t_bf16, backward_qt_bf16, backward_bf16_square_double_mxfp8,
mxfp4_transpose_mxfp8; reference qutlass/__init__.py:206-309)

and the same op names/schemas under ``torch.ops._qutlass_C`` (bindings.cpp:498-507).  Host side =
this module (argument checks, output allocation, stream selection) calling hand-written sm_100a
CUDA in ``lib/libb200q.so`` through the thin C-ABI of ``include/b200q.h``.  No CUTLASS, FlashInfer,
Triton, multi-backend dispatch or CPU fallback: without the built library or without a CUDA
device every compute entry point raises.
"""
from __future__ import annotations

import os
from typing import Literal

import torch

from . import _lib
from .utils import (get_padded_shape_mx, get_padded_shape_nv, pad_to_block, to_blocked,  # noqa: F401
                    _attach_blocked, _detach_blocked, _version_of)

__all__ = [
    "matmul_mxf4_bf16_tn", "matmul_nvf4_bf16_tn", "fusedQuantizeMx", "fusedQuantizeNv",
    "matmul_ada_mxf4_bf16_tn", "matmul_mxf8_bf16_tn", "matmul_mxf8_bf16_nn", "backward_t_bf16",
    "backward_qt_bf16", "backward_bf16_square_double_mxfp8", "mxfp4_transpose_mxfp8", "fused_linear_fp4",
]

METHOD_QUEST, METHOD_ABSMAX = 0, 1
ROT_TRUSTED_HADAMARD = 0x100   # include/b200q.h: B200Q_ROT_TRUSTED_HADAMARD
ROT_GENERIC = 0x200            # include/b200q.h: B200Q_ROT_GENERIC (known non-Hadamard -> tensor-core rotation)
NV_SM100_CODES = 0x400         # include/b200q.h: B200Q_NV_SM100_CODES (accepted, no effect: the behaviour is the default now)
NV_ORACLE_CODES = 0x800        # include/b200q.h: B200Q_NV_ORACLE_CODES (NVFP4 abs_max H=128: the reference ORACLE's arithmetic)
KIND_MXF4, KIND_NVF4, KIND_MXF8, KIND_MXF8_NN = 0, 1, 2, 3
GEMM_STATIC_WEIGHTS = 0x100    # include/b200q.h: B200Q_GEMM_STATIC_WEIGHTS (OR into kind)


def _load_compiled_ops() -> bool:
    """torch.ops._qutlass_C / _b200q_C from the COMPILED op layer (csrc/torch_ops.cpp -> lib/b200q_torch_ops.so, torch stable
    ABI like the reference's bindings.cpp:498-540).  False -- and the Python-registered ops + ctypes calls below take over --
    only when another library already owns the `_qutlass_C` names in this process (tools that load the compiled REFERENCE
    first to race it) or when the profiling build of the C library was selected (B200Q_LIB=prof)."""
    if os.environ.get("B200Q_LIB") == "prof" or os.environ.get("B200Q_NO_COMPILED_OPS") == "1":
        return False
    path = os.path.join(os.path.dirname(_lib.LIB_PATH), "b200q_torch_ops.so")
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -m qutlass_b200.build`")
    if hasattr(torch.ops._qutlass_C, "matmul_mxf4_bf16_tn"):
        return False
    _lib.load()
    torch.ops.load_library(path)
    return True


_COMPILED_OPS = False


def _check(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _stream(t: torch.Tensor) -> int:
    # raw cudaStream_t of torch's current stream on the tensor's device (no Stream object construction)
    return torch._C._cuda_getCurrentRawStream(t.device.index)


class _DeviceGuard:
    """Cheap device guard: only switches when the tensor lives on a non-current device (the reference's
    launchers wrap every launch in a DeviceGuard, gemm.cu:164)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def _check_cuda_same(name: str, tensors) -> None:
    dev = None
    for label, t in tensors:
        _check(t.is_cuda, f"{name}: expected all tensors to be on CUDA, but {label} is on {t.device}")
        if dev is None:
            dev = t.device
        _check(t.device == dev, f"{name}: expected all tensors on the same GPU, but {label} is on {t.device} (vs {dev})")


def _check_contig(name: str, tensors) -> None:
    for label, t in tensors:
        _check(t.is_contiguous(), f"{name}: expected {label} to be contiguous")


# --------------------------------------------------------------------------------------- GEMM
def _matmul_fp4(name: str, a, b, a_sf, b_sf, alpha, kind: int, sf_dtype, min_k_bytes: int, cfg=(0, 0),
                static_weights: bool = False):
    """reference checks: qutlass/csrc/bindings.cpp:32-102 (compiled op layer: csrc/torch_ops.cpp gemm_impl; the Python copy
    below only runs when that layer is not loaded)"""
    if _COMPILED_OPS:
        return torch.ops._b200q_C.gemm_fp4(a, b, a_sf, b_sf, alpha, kind | (GEMM_STATIC_WEIGHTS if static_weights else 0),
                                           cfg[0], cfg[1])
    _check_contig(name, [("A", a), ("B", b), ("A_sf", a_sf), ("B_sf", b_sf)])
    _check_cuda_same(name, [("A", a), ("B", b), ("A_sf", a_sf), ("B_sf", b_sf), ("alpha", alpha)])
    f8 = kind in (KIND_MXF8, KIND_MXF8_NN)
    nn = kind == KIND_MXF8_NN            # A is stored [K, M] (reference: bindings.cpp:179-216)
    op_dtype, op_name = (torch.float8_e4m3fn, "float8_e4m3fn") if f8 else (torch.uint8, "uint8")
    _check(a.dtype == op_dtype, f"A must be {op_name}")
    _check(b.dtype == op_dtype, f"B must be {op_name}")
    sf_name = "float8_e4m3fn" if kind == KIND_NVF4 else "float8_e8m0fnu"
    _check(a_sf.dtype == sf_dtype, f"A_sf must be {sf_name}")
    _check(b_sf.dtype == sf_dtype, f"B_sf must be {sf_name}")
    _check(a.dim() == 2 and b.dim() == 2, "A and B must be 2D")
    if nn:
        _check(a.size(0) == b.size(1), "Inner dimensions must match for A.T @ B.T")
        _check(a.size(0) >= min_k_bytes, f"A K-dim must be >= {min_k_bytes}")
        _check(a.size(1) % 16 == 0, f"M ({a.size(1)}) must be a multiple of 16")
    else:
        _check(a.size(1) == b.size(1), "Inner dimensions must match for A @ B.T")
        _check(a.size(1) >= min_k_bytes, f"A K-dim must be >= {min_k_bytes}")
    _check(b.size(1) >= min_k_bytes, f"B K-dim must be >= {min_k_bytes}")
    _check(alpha.dtype == torch.float32 and alpha.numel() >= 1, "alpha must be a float32 tensor with one element")
    m, n, k = (a.size(1) if nn else a.size(0)), b.size(0), b.size(1) * (1 if f8 else 2)
    group = 16 if kind == KIND_NVF4 else 32
    _check(k % 32 == 0, f"K ({k}) must be a multiple of 32")
    need_a = ((m + 127) // 128) * 128 * (((k // group) + 3) // 4) * 4
    need_b = ((n + 127) // 128) * 128 * (((k // group) + 3) // 4) * 4
    _check(a_sf.numel() >= need_a, f"A_sf has {a_sf.numel()} scales, the blocked layout needs {need_a}")
    _check(b_sf.numel() >= need_b, f"B_sf has {b_sf.numel()} scales, the blocked layout needs {need_b}")
    out = torch.empty((m, n), dtype=torch.bfloat16, device=a.device)
    with _DeviceGuard(a.device):
        _lib.check(_lib.load().b200q_gemm_fp4_cfg(
            a.data_ptr(), b.data_ptr(), a_sf.data_ptr(), b_sf.data_ptr(), alpha.data_ptr(), out.data_ptr(),
            m, n, k, kind | (GEMM_STATIC_WEIGHTS if static_weights else 0), cfg[0], cfg[1], _stream(a)))
    return out


def _backend_gate(backend: str) -> None:
    if backend == "cutlass":
        return  # the name is kept for drop-in compatibility; the kernel underneath is our own tcgen05 GEMM
    if backend == "flashinfer":
        raise ImportError("flashinfer backend requested but qutlass_b200 ships a single sm_100a backend "
                          "(no multi-backend dispatch)")
    raise ValueError(f"invalid backend {backend!r}; use 'cutlass' or 'flashinfer'")


def matmul_mxf4_bf16_tn(a: torch.Tensor, b: torch.Tensor, a_sf: torch.Tensor, b_sf: torch.Tensor,
                        alpha: torch.Tensor, backend: Literal["cutlass", "flashinfer"] = "cutlass", *,
                        static_weights: bool = False) -> torch.Tensor:
    """D = bf16(alpha * dq(a) @ dq(b).T), MXFP4 (reference: qutlass/__init__.py:34-76).

    ``static_weights=True`` (extension, keyword-only): the caller guarantees ``b`` / ``b_sf`` are not produced by the
    kernels right in front of this call on the stream (weights quantised once); the GEMM then prefetches them while the
    preceding kernel drains (include/b200q.h: B200Q_GEMM_STATIC_WEIGHTS).  The default is safe for any call order."""
    _backend_gate(backend)
    if _COMPILED_OPS and not static_weights:
        return torch.ops._qutlass_C.matmul_mxf4_bf16_tn(a, b, a_sf, b_sf, alpha)      # like qutlass/__init__.py:42-43
    return _matmul_fp4("matmul_mxf4_bf16_tn", a, b, a_sf, b_sf, alpha, KIND_MXF4, torch.float8_e8m0fnu, 32,
                       static_weights=static_weights)


def matmul_nvf4_bf16_tn(a: torch.Tensor, b: torch.Tensor, a_sf: torch.Tensor, b_sf: torch.Tensor,
                        alpha: torch.Tensor, backend: Literal["cutlass", "flashinfer"] = "cutlass", *,
                        static_weights: bool = False) -> torch.Tensor:
    """D = bf16(alpha * dq(a) @ dq(b).T), NVFP4 (reference: qutlass/__init__.py:89-131).  ``static_weights``: see
    matmul_mxf4_bf16_tn."""
    _backend_gate(backend)
    if _COMPILED_OPS and not static_weights:
        return torch.ops._qutlass_C.matmul_nvf4_bf16_tn(a, b, a_sf, b_sf, alpha)
    return _matmul_fp4("matmul_nvf4_bf16_tn", a, b, a_sf, b_sf, alpha, KIND_NVF4, torch.float8_e4m3fn, 16,
                       static_weights=static_weights)


def matmul_mxf8_bf16_tn(a: torch.Tensor, b: torch.Tensor, block_scale_a: torch.Tensor, block_scale_b: torch.Tensor,
                        alpha: torch.Tensor) -> torch.Tensor:
    """D = bf16(alpha * dq(a) @ dq(b).T), MXFP8: e4m3 operands [M,K] / [N,K], e8m0 blocked scales per 32
    (reference: qutlass/__init__.py:134-139, bindings.cpp:140-176, gemm.cu:328-380) -- the first "next" row after
    the FP4 path, same tcgen05 kernel with kind::mxf8f6f4."""
    return _matmul_fp4("matmul_mxf8_bf16_tn", a, b, block_scale_a, block_scale_b, alpha, KIND_MXF8,
                       torch.float8_e8m0fnu, 32)


def matmul_mxf8_bf16_nn(a: torch.Tensor, b: torch.Tensor, block_scale_a: torch.Tensor, block_scale_b: torch.Tensor,
                        alpha: torch.Tensor) -> torch.Tensor:
    """As matmul_mxf8_bf16_tn with A stored TRANSPOSED, a = [K, M] e4m3 (M contiguous): D[m, n] = sum_k a[k, m] b[n, k]
    (reference: qutlass/__init__.py:141-146, bindings.cpp:179-216, gemm.cu:388-434; tests/mxfp8_test.py:77-96).
    ``block_scale_a`` is still the blocked scale buffer of the logical [M, K/32] matrix.  The tile of A is loaded by
    TMA as 128 K-rows x 128 M-bytes and consumed as an MN-major tcgen05 operand -- no transpose pass."""
    return _matmul_fp4("matmul_mxf8_bf16_nn", a, b, block_scale_a, block_scale_b, alpha, KIND_MXF8_NN,
                       torch.float8_e8m0fnu, 32)


# --------------------------------------------------------------------------------------- quantise
_ROT_CACHE = {}   # id(tensor) -> (weakref, _version, data_ptr, is_hadamard)
_ROT_DEAD = [0]          # inspected rotation tensors that have since been garbage-collected (= built per call by the caller)
_ROT_MAX_DEAD = 64       # after that many throw-away rotations stop synchronising for new ones: the device-side check runs


def _rotation_hint(r: torch.Tensor) -> int:
    """Exact host-side classification of the rotation matrix, cached per tensor object + version.

    The reference's API takes *any* matrix at run time (README "loaded at runtime"); the kernel can verify the
    Sylvester-Hadamard structure itself on the device (graph-safe) but that check is redundant work on every
    call.  Here the (<= 32 KB) matrix is inspected ONCE on the host -- one small D2H copy -- and the result is
    remembered for as long as that tensor object lives unmodified.  No hint (0: the device-side check runs, always
    correct) whenever the host cannot vouch for the matrix: during CUDA-graph capture, for inference-mode tensors
    (no version counter: an in-place edit would go unnoticed), and once a caller has shown that it builds a new
    rotation tensor for every call (_ROT_MAX_DEAD inspected tensors already garbage-collected -- every inspection is a
    device synchronisation; long-lived per-layer rotations never count)."""
    import weakref
    key = id(r)
    ent = _ROT_CACHE.get(key)
    ver = _version_of(r)
    if ver < 0:
        return 0
    if ent is not None and ent[0]() is r and ent[1] == ver and ent[2] == r.data_ptr():
        return ROT_TRUSTED_HADAMARD if ent[3] else ROT_GENERIC
    if r.is_cuda and torch.cuda.is_current_stream_capturing():
        return 0
    dead = [k for k, v in _ROT_CACHE.items() if v[0]() is None]
    for k in dead:
        _ROT_CACHE.pop(k, None)
    _ROT_DEAD[0] += len(dead)
    if _ROT_DEAD[0] >= _ROT_MAX_DEAD:
        return 0
    h = r.size(0)
    m = r.detach().to("cpu").view(torch.int16)                 # bit patterns (synchronises once)
    idx = torch.arange(h)
    bits = idx[:, None] & idx[None, :]
    par = torch.zeros_like(bits)
    while bits.any():
        par ^= bits & 1
        bits = bits >> 1
    c = int(m[0, 0])
    want = torch.where(par.bool(), torch.tensor(c ^ -0x8000, dtype=torch.int32), torch.tensor(c, dtype=torch.int32))
    want = ((want + 0x8000) % 0x10000 - 0x8000).to(torch.int16)
    is_h = bool(torch.equal(m, want))
    _ROT_CACHE[key] = (weakref.ref(r), ver, r.data_ptr(), is_h)
    return ROT_TRUSTED_HADAMARD if is_h else ROT_GENERIC


def _quant_checks(name: str, a, r, outs, extra=()):
    """reference checks: qutlass/csrc/bindings.cpp:218-252,292-333,335-426"""
    _check_contig(name, [("A", a), ("B", r)] + [(f"OUT{i}", o) for i, o in enumerate(outs)])
    _check_cuda_same(name, [("A", a), ("B", r)] + [(f"OUT{i}", o) for i, o in enumerate(outs)] + list(extra))
    _check(a.dtype == torch.bfloat16, "A must be bf16")
    _check(r.dtype == torch.bfloat16, "B must be bf16")
    _check(r.dim() == 2 and r.size(0) == r.size(1), "Rotation matrix must be square")
    had = r.size(0)
    _check(a.numel() % had == 0, f"A must be divisible by{had}")
    return had


def _quantize_mx_into(a, r, out, out_sf, out_sf_blocked, out_mask, method: int):
    if _COMPILED_OPS and out_sf is not None:
        _detach_blocked(out_sf)
        torch.ops._b200q_C.quantize_mx(a, r, out, out_sf, out_sf_blocked, out_mask, method | _rotation_hint(r))
        return
    had = _quant_checks("fusedQuantizeMx", a, r, [out, out_sf])
    _check(had in (32, 64, 128), f"Unsupported rotation size {had}; expected 32, 64, or 128.")
    _check(a.size(-1) % 32 == 0, "last dimension of A must be a multiple of 32")
    if out_sf is not None:
        _detach_blocked(out_sf)     # OUT_sf is overwritten through its data pointer: a blocked copy attached earlier is stale
    with _DeviceGuard(a.device):
        _lib.check(_lib.load().b200q_quantize_mx(
            a.data_ptr(), r.data_ptr(), out.data_ptr(), out_sf.data_ptr() if out_sf is not None else None,
            out_sf_blocked.data_ptr() if out_sf_blocked is not None else None,
            out_mask.data_ptr() if out_mask is not None else None,
            a.numel(), a.size(-1), had, method | _rotation_hint(r), _stream(a)))


def _quantize_nv_into(a, r, out, out_sf, out_sf_blocked, global_scale, method: int):
    if _COMPILED_OPS and out_sf is not None:
        _detach_blocked(out_sf)
        flags = method | _rotation_hint(r)
        if method == METHOD_ABSMAX and r.dim() == 2 and r.size(0) == 128 and os.environ.get("B200Q_NV128_ORACLE_CODES") == "1":
            flags |= NV_ORACLE_CODES
        torch.ops._b200q_C.quantize_nv(a, r, out, out_sf, out_sf_blocked, global_scale, flags)
        return
    had = _quant_checks("fusedQuantizeNv", a, r, [out, out_sf], extra=[("global_scale", global_scale)])
    _check(global_scale.dtype == torch.float32, "global_scale must be float")
    _check(global_scale.dim() == 1 and global_scale.size(0) == 1, "global_scale must be a scalar")
    _check(had in (16, 32, 64, 128), f"Unsupported rotation size {had}; expected 16, 32, 64, or 128.")
    _check(a.size(-1) % 32 == 0, "last dimension of A must be a multiple of 32")
    flags = method | _rotation_hint(r)
    if out_sf is not None:
        _detach_blocked(out_sf)
    if had == 128 and method == METHOD_ABSMAX and os.environ.get("B200Q_NV128_ORACLE_CODES") == "1":
        # abs_max + Hadamard-128: by default bit-compatible with the reference's sm_100-only kernel (codes from the UNROUNDED
        # scale, include/b200q.h / DESIGN.md section 4); this switch selects the arithmetic of its other kernels / test oracle
        flags |= NV_ORACLE_CODES
    with _DeviceGuard(a.device):
        _lib.check(_lib.load().b200q_quantize_nv(
            a.data_ptr(), r.data_ptr(), out.data_ptr(), out_sf.data_ptr() if out_sf is not None else None,
            out_sf_blocked.data_ptr() if out_sf_blocked is not None else None, global_scale.data_ptr(),
            a.numel(), a.size(-1), had, flags, _stream(a)))


def fusedQuantizeMx(a: torch.Tensor, b: torch.Tensor, *, method: Literal["quest", "abs_max"] = "quest",
                    return_mask: bool = False):
    """Fused rotate (x_group @ b) + MXFP4 quantise (reference: qutlass/__init__.py:149-180).

    Returns (e2m1 packed uint8 [..., K/2], e8m0 scales [pad128(rows), pad4(K/32)] row-major) exactly like
    the reference (+ the packed clip mask with return_mask=True).  The scale tensor additionally carries
    the block-scaled copy the kernel wrote in the same pass, so ``to_blocked(scales)`` is a no-op.
    """
    if method not in ("quest", "abs_max"):
        raise ValueError(f"invalid method {method!r}, must be 'quest' or 'abs_max'")
    if method == "abs_max" and return_mask:
        raise ValueError("return_mask is only supported for method 'quest'")
    padded_rows, padded_cols = get_padded_shape_mx(a)
    xh_e2m1 = torch.empty(*a.shape[:-1], a.size(-1) // 2, dtype=torch.uint8, device=a.device)
    xh_e8m0 = torch.empty(padded_rows, padded_cols, dtype=torch.float8_e8m0fnu, device=a.device)
    blocked = torch.empty(padded_rows * padded_cols, dtype=torch.float8_e8m0fnu, device=a.device)
    clip_mask = None
    if return_mask:
        clip_mask = torch.empty(*a.shape[:-1], a.size(-1) // 8, dtype=torch.uint8, device=a.device)
    _quantize_mx_into(a, b, xh_e2m1, xh_e8m0, blocked, clip_mask,
                      METHOD_QUEST if method == "quest" else METHOD_ABSMAX)
    _attach_blocked(xh_e8m0, blocked)
    if return_mask:
        return xh_e2m1, xh_e8m0, clip_mask
    return xh_e2m1, xh_e8m0


def fusedQuantizeNv(a: torch.Tensor, b: torch.Tensor, global_scale: torch.Tensor, *,
                    method: Literal["quest", "abs_max"] = "abs_max"):
    """Fused rotate + NVFP4 quantise (reference: qutlass/__init__.py:183-203)."""
    if method not in ("quest", "abs_max"):
        raise ValueError(f"invalid method {method!r}, must be 'quest' or 'abs_max'")
    padded_rows, padded_cols = get_padded_shape_nv(a)
    xh_e2m1 = torch.empty(*a.shape[:-1], a.size(-1) // 2, dtype=torch.uint8, device=a.device)
    xh_e4m3 = torch.empty(padded_rows, padded_cols, dtype=torch.float8_e4m3fn, device=a.device)
    blocked = torch.empty(padded_rows * padded_cols, dtype=torch.float8_e4m3fn, device=a.device)
    _quantize_nv_into(a, b, xh_e2m1, xh_e4m3, blocked, global_scale,
                      METHOD_QUEST if method == "quest" else METHOD_ABSMAX)
    _attach_blocked(xh_e4m3, blocked)
    return xh_e2m1, xh_e4m3


# --------------------------------------------------------------------------------------- out of scope
_FUSE_WS = {}   # (device index, stream handle) -> zeroed progress-counter workspace of the fused kernel


def _fuse_workspace(device: torch.device, stream: int, m: int) -> torch.Tensor:
    need = _lib.load().b200q_linear_fp4_workspace_bytes(m)
    key = (device.index, stream)
    ws = _FUSE_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 4096), dtype=torch.uint8, device=device)   # zeroed once; every call leaves it zeroed
        _FUSE_WS[key] = ws
    return ws


def fused_linear_fp4(x: torch.Tensor, rot: torch.Tensor, w_q: torch.Tensor, w_sf: torch.Tensor, alpha: torch.Tensor, *,
                     global_scale: torch.Tensor | None = None, method: Literal["quest", "abs_max"] = "abs_max",
                     fmt: Literal["mx", "nv"] = "mx"):
    """The per-layer forward sequence of the reference in one call (benchmarks/bench_mxfp4_sm100.py:93-104):

        xq, x_sf = fusedQuantizeMx(x, rot, method=method)          # or fusedQuantizeNv(x, rot, global_scale)
        out = matmul_mxf4_bf16_tn(xq, w_q, to_blocked(x_sf), w_sf, alpha)

    Returns (out [M, N] bf16, xq, x_sf) -- bit-identical to the two calls.  By default it IS the two launches behind one
    C-ABI call (include/b200q.h: b200q_linear_fp4); with B200Q_FUSE=1, a Hadamard rotation, K % 1024 == 0, N % 8 == 0 and
    M > 256 it runs as one persistent kernel with quantiser warps inside the GEMM (measured: not faster on a power-limited
    B200, profiles/r01_notes.md).  ``w_sf`` is the blocked (to_blocked) weight scale buffer.
    """
    if method not in ("quest", "abs_max"):
        raise ValueError(f"invalid method {method!r}, must be 'quest' or 'abs_max'")
    if fmt not in ("mx", "nv"):
        raise ValueError(f"invalid fmt {fmt!r}, must be 'mx' or 'nv'")
    nv = fmt == "nv"
    name = "fused_linear_fp4"
    extra = [("W", w_q), ("W_sf", w_sf), ("alpha", alpha)] + ([("global_scale", global_scale)] if nv else [])
    had = _quant_checks(name, x, rot, [], extra=extra)
    _check_contig(name, [("W", w_q), ("W_sf", w_sf)])
    _check(x.dim() == 2 and w_q.dim() == 2, "x and W must be 2D")
    _check(w_q.dtype == torch.uint8, "W must be uint8")
    sf_dtype = torch.float8_e4m3fn if nv else torch.float8_e8m0fnu
    _check(w_sf.dtype == sf_dtype, f"W_sf must be {'float8_e4m3fn' if nv else 'float8_e8m0fnu'}")
    _check(alpha.dtype == torch.float32 and alpha.numel() >= 1, "alpha must be a float32 tensor with one element")
    if nv:
        _check(global_scale is not None and global_scale.dtype == torch.float32 and global_scale.numel() == 1,
               "global_scale must be a float32 tensor with one element")
        _check(had in (16, 32, 64, 128), f"Unsupported rotation size {had}; expected 16, 32, 64, or 128.")
    else:
        _check(had in (32, 64, 128), f"Unsupported rotation size {had}; expected 32, 64, or 128.")
    m, k = x.shape
    n = w_q.size(0)
    _check(k % 32 == 0, f"K ({k}) must be a multiple of 32")
    _check(w_q.size(1) * 2 == k, "Inner dimensions must match for A @ B.T")
    group = 16 if nv else 32
    need_w = ((n + 127) // 128) * 128 * (((k // group) + 3) // 4) * 4
    _check(w_sf.numel() >= need_w, f"W_sf has {w_sf.numel()} scales, the blocked layout needs {need_w}")
    padded_rows, padded_cols = (get_padded_shape_nv if nv else get_padded_shape_mx)(x)
    xq = torch.empty(m, k // 2, dtype=torch.uint8, device=x.device)
    x_sf = torch.empty(padded_rows, padded_cols, dtype=sf_dtype, device=x.device)
    blocked = torch.empty(padded_rows * padded_cols, dtype=sf_dtype, device=x.device)
    out = torch.empty(m, n, dtype=torch.bfloat16, device=x.device)
    meth = (METHOD_QUEST if method == "quest" else METHOD_ABSMAX) | _rotation_hint(rot)
    if nv and had == 128 and method == "abs_max" and os.environ.get("B200Q_NV128_ORACLE_CODES") == "1":
        meth |= NV_ORACLE_CODES
    with _DeviceGuard(x.device):
        stream = _stream(x)
        capturing = torch.cuda.is_current_stream_capturing()
        # the counter workspace is keyed by stream; under graph capture the capture stream's handle is stable too
        ws = _fuse_workspace(x.device, stream, m) if not capturing or (x.device.index, stream) in _FUSE_WS else None
        _lib.check(_lib.load().b200q_linear_fp4(
            x.data_ptr(), rot.data_ptr(), xq.data_ptr(), x_sf.data_ptr(), blocked.data_ptr(), w_q.data_ptr(),
            w_sf.data_ptr(), alpha.data_ptr(), global_scale.data_ptr() if nv else None, out.data_ptr(),
            ws.data_ptr() if ws is not None else None, m, n, k, had, meth, KIND_NVF4 if nv else KIND_MXF4, stream))
    _attach_blocked(x_sf, blocked)
    return out, xq, x_sf


def _out_of_scope(name: str):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            f"qutlass_b200.{name}: outside the microscaled-FP4 forward hot path this build covers "
            "(SURVEY.md section 8 'out of scope' / 'next').")
    fn.__name__ = name
    return fn


matmul_ada_mxf4_bf16_tn = _out_of_scope("matmul_ada_mxf4_bf16_tn")      # sm_120-only prototype (gemm_ada.cu)


# --------------------------------------------------------------------------------------- backward re-quantisers
def _bwd_rot_checks(name: str, h: torch.Tensor) -> int:
    _check(h.dtype == torch.bfloat16, f"{name}: h must be bf16")
    _check(h.dim() == 2 and h.size(0) == 32 and h.size(1) == 32, f"{name}: h must be a 32 x 32 rotation matrix")
    return ROT_TRUSTED_HADAMARD if _rotation_hint(h) == ROT_TRUSTED_HADAMARD else 0


def _backward_t_bf16_into(x, h, xh_e2m1, xh_e8m0) -> None:
    name = "backward_t_bf16"
    _check_cuda_same(name, [("x", x), ("h", h), ("xh_e2m1", xh_e2m1), ("xh_e8m0", xh_e8m0)])
    _check_contig(name, [("x", x), ("h", h), ("xh_e2m1", xh_e2m1), ("xh_e8m0", xh_e8m0)])
    _check(x.dtype == torch.bfloat16, f"{name}: x must be bf16")
    _check(x.dim() >= 2, f"{name}: x must have at least 2 dimensions")
    size_m, size_n = x.size(-1), x.size(-2)
    size_b = x.numel() // max(size_m * size_n, 1)
    _check(size_n % 32 == 0, f"{name}: x.size(-2) ({size_n}) must be a multiple of 32")
    _check(size_m % 8 == 0, f"{name}: x.size(-1) ({size_m}) must be a multiple of 8")
    _check(xh_e2m1.numel() * xh_e2m1.element_size() == size_b * size_m * size_n // 2, f"{name}: xh_e2m1 has the wrong size")
    _check(xh_e8m0.numel() == size_b * size_m * size_n // 32, f"{name}: xh_e8m0 has the wrong size")
    flags = _bwd_rot_checks(name, h)
    with _DeviceGuard(x.device):
        _lib.check(_lib.load().b200q_backward_t_bf16(x.data_ptr(), h.data_ptr(), xh_e2m1.data_ptr(), xh_e8m0.data_ptr(),
                                                     size_m, size_n, size_b, flags, _stream(x)))


def backward_t_bf16(x: torch.Tensor, h: torch.Tensor, xh_e2m1: torch.Tensor = None, xh_e8m0: torch.Tensor = None):
    """MXFP4 abs-max quantisation of rotate(x.transpose(-2, -1)) without materialising the transpose
    (reference: qutlass/__init__.py:206-244, quartet_bwd_sm120.cu:237-318; pinned by tests/quartet_test.py:220-226).
    x [..., N, M] bf16 -> (e2m1 [..., M, N/2], e8m0 [..., M, N/32])."""
    if xh_e2m1 is None:
        xh_e2m1 = torch.empty(*x.shape[:-2], x.size(-1), x.size(-2) // 2, dtype=torch.float4_e2m1fn_x2, device=h.device)
    if xh_e8m0 is None:
        xh_e8m0 = torch.empty(*x.shape[:-2], x.size(-1), x.size(-2) // 32, dtype=torch.float8_e8m0fnu, device=h.device)
    assert (x.dtype == h.dtype == torch.bfloat16 and xh_e2m1.dtype == torch.float4_e2m1fn_x2
            and xh_e8m0.dtype == torch.float8_e8m0fnu)
    assert x.is_contiguous() and h.is_contiguous() and xh_e2m1.is_contiguous() and xh_e8m0.is_contiguous()
    _backward_t_bf16_into(x, h, xh_e2m1, xh_e8m0)
    return xh_e2m1, xh_e8m0


def _backward_qt_bf16_into(x_e2m1, x_e8m0, h, alpha, xh_e2m1, xh_e8m0) -> None:
    name = "backward_qt_bf16"
    ts = [("x_e2m1", x_e2m1), ("x_e8m0", x_e8m0), ("h", h), ("xh_e2m1", xh_e2m1), ("xh_e8m0", xh_e8m0)]
    _check_cuda_same(name, ts + [("alpha", alpha)])
    _check_contig(name, ts)
    _check(x_e2m1.element_size() == 1 and x_e8m0.element_size() == 1, f"{name}: x_e2m1 / x_e8m0 must be 1-byte dtypes")
    _check(alpha.dtype == torch.float32 and alpha.numel() >= 1, f"{name}: alpha must be a float32 tensor with one element")
    _check(x_e2m1.dim() >= 2, f"{name}: x_e2m1 must have at least 2 dimensions")
    size_m, size_n = x_e2m1.size(-1) * 2, x_e2m1.size(-2)
    size_b = x_e2m1.numel() // max(x_e2m1.size(-1) * size_n, 1)
    _check(size_n % 32 == 0, f"{name}: x_e2m1.size(-2) ({size_n}) must be a multiple of 32")
    _check(size_m % 32 == 0, f"{name}: 2 * x_e2m1.size(-1) ({size_m}) must be a multiple of 32")
    _check(x_e8m0.numel() == size_b * size_n * size_m // 32, f"{name}: x_e8m0 has the wrong size")
    _check(xh_e2m1.numel() * xh_e2m1.element_size() == size_b * size_m * size_n // 2, f"{name}: xh_e2m1 has the wrong size")
    _check(xh_e8m0.numel() == size_b * size_m * size_n // 32, f"{name}: xh_e8m0 has the wrong size")
    flags = _bwd_rot_checks(name, h)
    with _DeviceGuard(h.device):
        _lib.check(_lib.load().b200q_backward_qt_bf16(
            x_e2m1.data_ptr(), x_e8m0.data_ptr(), h.data_ptr(), alpha.data_ptr(), xh_e2m1.data_ptr(), xh_e8m0.data_ptr(),
            size_m, size_n, size_b, flags, _stream(h)))


def backward_qt_bf16(x_e2m1: torch.Tensor, x_e8m0: torch.Tensor, h: torch.Tensor, alpha: torch.Tensor,
                     xh_e2m1: torch.Tensor = None, xh_e8m0: torch.Tensor = None):
    """Dequantise an MXFP4 tensor, transpose, rotate and re-quantise (abs-max, scale / alpha) in one pass
    (reference: qutlass/__init__.py:247-283, quartet_bwd_sm120.cu:320-412; pinned by tests/quartet_test.py:228-239).
    x_e2m1 [..., N, M/2], x_e8m0 [..., N, M/32] -> (e2m1 [..., M, N/2], e8m0 [..., M, N/32])."""
    if xh_e2m1 is None:
        xh_e2m1 = torch.empty(*x_e2m1.shape[:-2], x_e2m1.size(-1) * 2, x_e2m1.size(-2) // 2,
                              dtype=torch.float4_e2m1fn_x2, device=h.device)
    if xh_e8m0 is None:
        xh_e8m0 = torch.empty(*x_e8m0.shape[:-2], x_e8m0.size(-1) * 32, x_e8m0.size(-2) // 32,
                              dtype=torch.float8_e8m0fnu, device=h.device)
    assert (x_e2m1.is_contiguous() and x_e8m0.is_contiguous() and h.is_contiguous() and xh_e2m1.is_contiguous()
            and xh_e8m0.is_contiguous())
    _backward_qt_bf16_into(x_e2m1, x_e8m0, h, alpha, xh_e2m1, xh_e8m0)
    return xh_e2m1, xh_e8m0


def _square_double_into(x_bf16, x_fp8, row_scales, column_scales) -> None:
    name = "backward_bf16_square_double_mxfp8"
    ts = [("x_bf16", x_bf16), ("x_fp8", x_fp8), ("row_scales", row_scales), ("column_scales", column_scales)]
    _check_cuda_same(name, ts)
    _check_contig(name, ts)
    _check(x_bf16.dtype == torch.bfloat16 and x_bf16.dim() == 2, f"{name}: x_bf16 must be a 2-D bf16 tensor")
    m, n = x_bf16.shape
    m_pad = (m + 127) // 128 * 128
    _check(n % 32 == 0, f"{name}: x_bf16.size(1) ({n}) must be a multiple of 32")
    _check(x_fp8.numel() == m_pad * n and x_fp8.element_size() == 1, f"{name}: x_fp8 must hold {m_pad} x {n} bytes")
    _check(row_scales.numel() == m_pad * n // 32, f"{name}: row_scales must hold {m_pad} x {n // 32} bytes")
    _check(column_scales.numel() == n * m_pad // 32, f"{name}: column_scales must hold {n} x {m_pad // 32} bytes")
    with _DeviceGuard(x_bf16.device):
        _lib.check(_lib.load().b200q_backward_bf16_square_double_mxfp8(
            x_bf16.data_ptr(), m, n, x_fp8.data_ptr(), row_scales.data_ptr(), column_scales.data_ptr(), _stream(x_bf16)))


def backward_bf16_square_double_mxfp8(x_bf16: torch.Tensor):
    """bf16 [m, n] -> (e4m3 [m_pad, n], row scales [m_pad, n/32], column scales [n, m_pad/32]) with one e8m0 scale per
    32 x 32 tile, usable both as a K-major and as an MN-major MXFP8 GEMM operand (reference: qutlass/__init__.py:285-294,
    quartet_bwd_sm120.cu:497-623; pinned by tests/quartet_test.py:264-291,369-378).  The reference pads x to a multiple
    of 128 rows with a host-side copy; here the kernel reads the missing rows as zero."""
    _check(x_bf16.dim() == 2, "backward_bf16_square_double_mxfp8: x_bf16 must be 2-D")
    m, n = x_bf16.shape
    m_pad = (m + 127) // 128 * 128
    x_fp8 = torch.empty(m_pad, n, dtype=torch.float8_e4m3fn, device=x_bf16.device)
    row_scales = torch.empty(m_pad, n // 32, dtype=torch.float8_e8m0fnu, device=x_bf16.device)
    column_scales = torch.empty(n, m_pad // 32, dtype=torch.float8_e8m0fnu, device=x_bf16.device)
    _square_double_into(x_bf16.contiguous(), x_fp8, row_scales, column_scales)
    return x_fp8, row_scales, column_scales


def _mxfp4_transpose_mxfp8_into(x_fp4, scales, x_fp8, shared_exps, m: int) -> None:
    name = "mxfp4_transpose_mxfp8"
    ts = [("x_fp4", x_fp4), ("scales", scales), ("x_fp8", x_fp8), ("shared_exps", shared_exps)]
    _check_cuda_same(name, ts)
    _check_contig(name, ts)
    _check(x_fp4.dim() == 2 and x_fp4.element_size() == 1, f"{name}: x_fp4 must be a 2-D 1-byte tensor")
    _check(scales.element_size() == 1, f"{name}: scales must be a 1-byte dtype")
    n = x_fp4.size(1) * 2
    m_pad = (m + 255) // 256 * 256
    _check(m <= x_fp4.size(0), f"{name}: m ({m}) exceeds x_fp4.size(0)")
    _check(n % 32 == 0, f"{name}: 2 * x_fp4.size(1) ({n}) must be a multiple of 32")
    _check(scales.numel() >= m * (n // 32), f"{name}: scales must hold at least {m} x {n // 32} bytes")
    _check(x_fp8.numel() == n * m_pad and x_fp8.element_size() == 1, f"{name}: x_fp8 must hold {n} x {m_pad} bytes")
    _check(shared_exps.numel() == n * m_pad // 32, f"{name}: shared_exps must hold {n} x {m_pad // 32} bytes")
    with _DeviceGuard(x_fp4.device):
        _lib.check(_lib.load().b200q_mxfp4_transpose_mxfp8(
            x_fp4.data_ptr(), scales.data_ptr(), m, n, x_fp8.data_ptr(), shared_exps.data_ptr(), _stream(x_fp4)))


def mxfp4_transpose_mxfp8(x_fp4: torch.Tensor, scales: torch.Tensor):
    """MXFP4 [m, n/2] + e8m0 scales [>= m, n/32] -> MXFP8 of the transpose: (e4m3 [n, m_pad], e8m0 [n, m_pad/32]),
    m_pad = m rounded up to 256 (reference: qutlass/__init__.py:296-309, quartet_bwd_sm120.cu:625-734; pinned by
    tests/quartet_test.py:294-345,380-385).  The reference pads x_fp4 with a host-side copy and overwrites the pad rows
    of ``scales`` with 1.0; here the kernel reads rows >= m as zero and leaves ``scales`` untouched."""
    _check(x_fp4.dim() == 2, "mxfp4_transpose_mxfp8: x_fp4 must be 2-D")
    m, n = x_fp4.size(0), x_fp4.size(1) * 2
    m_pad = (m + 255) // 256 * 256
    x_fp8 = torch.empty(n, m_pad, dtype=torch.float8_e4m3fn, device=x_fp4.device)
    shared_exps = torch.empty(n, m_pad // 32, dtype=torch.float8_e8m0fnu, device=x_fp4.device)
    _mxfp4_transpose_mxfp8_into(x_fp4, scales, x_fp8, shared_exps, m)
    return x_fp8, shared_exps


# --------------------------------------------------------------------------------------- torch.ops._qutlass_C
def _register_ops() -> None:
    """Python registration of the same op names and schemas as the reference's STABLE_TORCH_LIBRARY_FRAGMENT(_qutlass_C)
    (qutlass/csrc/bindings.cpp:498-507), CUDA dispatch key.  FALLBACK ONLY: normally the compiled layer
    (lib/b200q_torch_ops.so) owns these names and this function is not called."""
    try:
        lib = torch.library.Library("_qutlass_C", "FRAGMENT")
    except Exception:  # pragma: no cover
        return
    defs = {
        "matmul_mxf4_bf16_tn": "(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor",
        "matmul_nvf4_bf16_tn": "(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor",
        "matmul_mxf8_bf16_tn": "(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor",
        "matmul_mxf8_bf16_nn": "(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor",
        # schema kept for importers; like the reference on sm_100 (gemm_ada.cu is sm_120-only) the call itself fails
        "matmul_ada_mxf4_bf16_tn": "(Tensor A, Tensor B, Tensor A_sf, Tensor B_sf, Tensor alpha) -> Tensor",
        "fusedQuantizeMxQuest": "(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf) -> (Tensor, Tensor)",
        "fusedQuantizeMxAbsMax": "(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf) -> (Tensor, Tensor)",
        "fusedQuantizeNvQuest": "(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf, Tensor global_scale) -> (Tensor, Tensor)",
        "fusedQuantizeNvAbsMax": "(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf, Tensor global_scale) -> (Tensor, Tensor)",
        "fusedQuantizeMxQuestWithMask": "(Tensor A, Tensor R, Tensor OUT, Tensor OUT_sf, Tensor OUT_mask) -> (Tensor, Tensor, Tensor)",
        "backward_t_bf16": "(Tensor x, Tensor h, Tensor xh_e2m1, Tensor xh_e8m0) -> ()",
        "backward_qt_bf16": "(Tensor x_e2m1, Tensor x_e8m0, Tensor h, Tensor alpha, Tensor xh_e2m1, Tensor xh_e8m0) -> ()",
        "backward_bf16_square_double_mxfp8": "(Tensor x_bf16, Tensor x_fp8, Tensor row_scales, Tensor column_scales) -> ()",
        "mxfp4_transpose_mxfp8": "(Tensor x_fp4, Tensor scales, Tensor x_fp8, Tensor shared_exps) -> ()",
    }
    impls = {
        "matmul_mxf4_bf16_tn": lambda A, B, A_sf, B_sf, alpha: matmul_mxf4_bf16_tn(A, B, A_sf, B_sf, alpha),
        "matmul_nvf4_bf16_tn": lambda A, B, A_sf, B_sf, alpha: matmul_nvf4_bf16_tn(A, B, A_sf, B_sf, alpha),
        "matmul_mxf8_bf16_tn": lambda A, B, A_sf, B_sf, alpha: matmul_mxf8_bf16_tn(A, B, A_sf, B_sf, alpha),
        "matmul_mxf8_bf16_nn": lambda A, B, A_sf, B_sf, alpha: matmul_mxf8_bf16_nn(A, B, A_sf, B_sf, alpha),
        "matmul_ada_mxf4_bf16_tn": lambda A, B, A_sf, B_sf, alpha: matmul_ada_mxf4_bf16_tn(A, B, A_sf, B_sf, alpha),
        "fusedQuantizeMxQuest": lambda A, R, OUT, OUT_sf: (_quantize_mx_into(A, R, OUT, OUT_sf, None, None, METHOD_QUEST), (OUT, OUT_sf))[1],
        "fusedQuantizeMxAbsMax": lambda A, R, OUT, OUT_sf: (_quantize_mx_into(A, R, OUT, OUT_sf, None, None, METHOD_ABSMAX), (OUT, OUT_sf))[1],
        "fusedQuantizeNvQuest": lambda A, R, OUT, OUT_sf, gs: (_quantize_nv_into(A, R, OUT, OUT_sf, None, gs, METHOD_QUEST), (OUT, OUT_sf))[1],
        "fusedQuantizeNvAbsMax": lambda A, R, OUT, OUT_sf, gs: (_quantize_nv_into(A, R, OUT, OUT_sf, None, gs, METHOD_ABSMAX), (OUT, OUT_sf))[1],
        "fusedQuantizeMxQuestWithMask": lambda A, R, OUT, OUT_sf, OUT_mask: (_quantize_mx_into(A, R, OUT, OUT_sf, None, OUT_mask, METHOD_QUEST), (OUT, OUT_sf, OUT_mask))[1],
        "backward_t_bf16": lambda x, h, xh_e2m1, xh_e8m0: _backward_t_bf16_into(x, h, xh_e2m1, xh_e8m0),
        "backward_qt_bf16": lambda x_e2m1, x_e8m0, h, alpha, xh_e2m1, xh_e8m0: _backward_qt_bf16_into(x_e2m1, x_e8m0, h, alpha, xh_e2m1, xh_e8m0),
        "backward_bf16_square_double_mxfp8": lambda x_bf16, x_fp8, row_scales, column_scales: _square_double_into(x_bf16, x_fp8, row_scales, column_scales),
        # like the reference binding (bindings.cpp:466-479) the raw op takes x_fp4 already padded to 256 rows
        "mxfp4_transpose_mxfp8": lambda x_fp4, scales, x_fp8, shared_exps: _mxfp4_transpose_mxfp8_into(x_fp4, scales, x_fp8, shared_exps, x_fp4.size(0)),
    }
    for name, schema in defs.items():
        try:
            lib.define(name + schema)
            lib.impl(name, impls[name], "CUDA")
        except Exception:  # already defined by another copy of the library in this process
            pass
    globals()["_OPS_LIB"] = lib  # keep alive


_COMPILED_OPS = _load_compiled_ops()
if not _COMPILED_OPS:
    _register_ops()
