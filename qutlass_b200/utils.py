"""Scale-layout helpers mirroring qutlass/utils.py of the reference (same names, same results),
without Triton: the swizzle is either already done by the quantise kernel (no-op) or a small CUDA kernel.
"""
from __future__ import annotations

import torch

from . import _lib

_BLOCKED_ATTR = "_b200q_blocked"


def ceil_div(a, b):
    return (a + b - 1) // b


def get_padded_shape_mx(a: torch.Tensor):
    """reference: qutlass/utils.py:140-147"""
    rows, cols = a.numel() // a.size(-1), a.size(-1) // 32
    return ceil_div(rows, 128) * 128, ceil_div(cols, 4) * 4


def get_padded_shape_nv(a: torch.Tensor):
    """reference: qutlass/utils.py:150-157"""
    rows, cols = a.numel() // a.size(-1), a.size(-1) // 16
    return ceil_div(rows, 128) * 128, ceil_div(cols, 4) * 4


def _version_of(t: torch.Tensor) -> int:
    """Tensor version counter, or -1 for inference tensors (created under torch.inference_mode -- as the reference's
    tests do -- they do not track versions; in-place edits of such a tensor are then not detectable)."""
    try:
        return t._version
    except RuntimeError:
        return -1


def _attach_blocked(sf_rowmajor: torch.Tensor, blocked: torch.Tensor) -> None:
    """Remember the blocked copy the quantise kernel wrote next to the row-major scales it belongs to.

    Only for tensors with a working version counter: an in-place edit of the row-major scales between the quantiser and
    to_blocked (``scales[m:] = 1.0``) must invalidate the copy, and inference tensors (torch.inference_mode) do not
    track versions -- for those nothing is attached and to_blocked runs its kernel."""
    ver = _version_of(sf_rowmajor)
    if ver >= 0:
        ptr = sf_rowmajor.data_ptr()
        # + the library's write generation of that buffer: a later quantise call that writes it through the C-ABI (e.g. the
        # raw torch.ops._qutlass_C ops, which do not bump torch's version counter) makes the attached copy stale
        setattr(sf_rowmajor, _BLOCKED_ATTR, (blocked, ver, ptr, _lib.load().b200q_sf_write_generation(ptr)))


def _detach_blocked(sf_rowmajor: torch.Tensor) -> None:
    """Called by every path that overwrites a scale tensor through its data pointer (the raw torch.ops._qutlass_C
    quantise ops write OUT_sf without bumping its version)."""
    if getattr(sf_rowmajor, _BLOCKED_ATTR, None) is not None:
        try:
            delattr(sf_rowmajor, _BLOCKED_ATTR)
        except AttributeError:
            pass


def to_blocked(input_matrix: torch.Tensor, use_triton_kernel: bool = False) -> torch.Tensor:
    """Row-major scales (H, W) -> flattened block-scaled layout (reference: qutlass/utils.py:160-193).

    If `input_matrix` came straight out of fusedQuantizeMx / fusedQuantizeNv, has not been modified since (version
    counter) and this is the first to_blocked call on it, the blocked copy the quantise kernel wrote in the same pass is
    HANDED OVER (no launch; the caller now owns that buffer exclusively, exactly like the fresh tensor the reference
    returns).  In every other case -- a second call, an edited tensor, an inference-mode tensor (no version counter), a
    foreign scale tensor -- one CUDA swizzle kernel runs.  `use_triton_kernel` is accepted for drop-in compatibility
    and ignored (there is no Triton here); like the reference's Triton path, inputs need not be pre-padded.
    """
    cached = getattr(input_matrix, _BLOCKED_ATTR, None)
    if cached is not None:
        _detach_blocked(input_matrix)
        if (cached[1] >= 0 and cached[1] == _version_of(input_matrix) and cached[2] == input_matrix.data_ptr()
                and cached[3] == _lib.load().b200q_sf_write_generation(cached[2])):
            return cached[0]
    assert input_matrix.dim() == 2, "to_blocked expects a 2-D scale matrix"
    assert input_matrix.element_size() == 1, "Expected element size to be 1 byte (8 bits)"
    if not input_matrix.is_cuda:
        raise RuntimeError("to_blocked: input must be a CUDA tensor (no CPU path in qutlass_b200)")
    x = input_matrix.contiguous()
    rows, cols = x.shape
    if not use_triton_kernel:
        # the reference's torch path asserts the input is already padded (utils.py:187)
        assert (rows, cols) == (ceil_div(rows, 128) * 128, ceil_div(cols, 4) * 4)
    if hasattr(torch.ops, "_b200q_C") and hasattr(torch.ops._b200q_C, "swizzle_sf"):
        return torch.ops._b200q_C.swizzle_sf(x)        # compiled op layer (csrc/torch_ops.cpp)
    out = torch.empty(ceil_div(rows, 128) * 128 * ceil_div(cols, 4) * 4, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200q_swizzle_sf(x.data_ptr(), out.data_ptr(), rows, cols,
                                                torch._C._cuda_getCurrentRawStream(x.device.index)))
    return out


def pad_to_block(tensor, dims, blocksize):
    """reference: qutlass/utils.py:196-204"""
    pad_dims = [0 for _ in range(2 * len(tensor.shape))]
    for dim in dims:
        size = tensor.shape[dim]
        next_multiple_of_block = ((size - 1) // blocksize + 1) * blocksize
        pad_dims[-2 * dim - 1] = next_multiple_of_block - size
    return torch.nn.functional.pad(tensor, pad_dims, "constant", 0.0)
