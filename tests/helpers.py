"""Shared helpers for the GPU parity tests and tools/bringup.py (oracle-side construction of inputs)."""
from __future__ import annotations

import numpy as np
import torch

import oracle as O


def bf16_tensor_from_f32(x: np.ndarray, device="cuda") -> torch.Tensor:
    """float32 array holding bf16-representable values -> torch bf16 tensor (exact)."""
    bits = O.bf16_bits(x).astype(np.int16)
    return torch.from_numpy(bits).view(torch.bfloat16).to(device)


def bf16_bits_of(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def u8_of(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.uint8).cpu().numpy()


def random_bf16(shape, seed=0, scale=25.0, dist="randn") -> np.ndarray:
    g = torch.Generator().manual_seed(seed)
    if dist == "randn":
        x = torch.randn(*shape, generator=g) * scale
    else:
        x = torch.rand(*shape, generator=g) * scale
    return x.to(torch.bfloat16).float().numpy()


def random_fp4_operand(rows: int, k: int, kind: str, seed: int, sf_mode: str = "narrow"):
    """Random packed e2m1 [rows, k/2] + scale bytes [rows, k/group] (row-major) for GEMM tests.

    sf_mode: 'one' -> all scales 1.0; 'narrow' -> exponents within +-1 (fp32 accumulation exact);
             'wide' -> several octaves (tolerance test)."""
    rng = np.random.default_rng(seed)
    group = 32 if kind == "mx" else 16
    q = rng.integers(0, 256, size=(rows, k // 2), dtype=np.uint8)
    n_sf = k // group
    if kind == "mx":
        if sf_mode == "one":
            sf = np.full((rows, n_sf), 127, dtype=np.uint8)
        elif sf_mode == "narrow":
            sf = rng.integers(126, 129, size=(rows, n_sf)).astype(np.uint8)
        else:
            sf = rng.integers(117, 138, size=(rows, n_sf)).astype(np.uint8)
    else:
        if sf_mode == "one":
            sf = np.full((rows, n_sf), 0x38, dtype=np.uint8)  # e4m3 1.0
        elif sf_mode == "narrow":
            sf = rng.choice(np.array([0x30, 0x38, 0x40, 0x34, 0x3C], dtype=np.uint8), size=(rows, n_sf))
        else:
            sf = rng.integers(0x08, 0x70, size=(rows, n_sf)).astype(np.uint8)
    return q, sf


def blocked_sf(sf_rowmajor: np.ndarray, fill: int = 0) -> np.ndarray:
    """row-major [rows, cols] scale bytes -> padded + blocked flat layout (oracle)."""
    rows, cols = sf_rowmajor.shape
    pr, pc = O.padded_sf_shape(rows, cols)
    padded = np.full((pr, pc), fill, dtype=np.uint8)
    padded[:rows, :cols] = sf_rowmajor
    return O.to_blocked(padded)


def random_f8_operand(rows: int, k: int, seed: int, sf_mode: str = "narrow"):
    """random e4m3 bytes [rows, k] (finite) + ue8m0 scale bytes [rows, k/32]."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, size=(rows, k), dtype=np.uint8)
    q = np.where((q & 0x7F) == 0x7F, q & 0xF0, q).astype(np.uint8)      # no NaN encodings
    if sf_mode == "narrow":
        sf = rng.integers(126, 129, size=(rows, k // 32)).astype(np.uint8)
    else:
        sf = rng.integers(117, 138, size=(rows, k // 32)).astype(np.uint8)
    return q, sf


def gemm_oracle_bits(aq, asf, bq, bsf, kind: str, alpha: float = 1.0) -> np.ndarray:
    dq = {"mx": O.dequant_mx, "nv": O.dequant_nv, "f8": O.dequant_mxf8}[kind]
    return O.gemm_ref(dq(aq, asf), dq(bq, bsf), alpha)


def sf_torch(arr: np.ndarray, kind: str, device="cuda") -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    return t.view(torch.float8_e8m0fnu if kind == "mx" else torch.float8_e4m3fn)


def run_gemm(aq, asf, bq, bsf, kind: str, alpha: float = 1.0, cfg=(0, 0)) -> np.ndarray:
    """Run the CUDA GEMM through the package's C-ABI path; returns bf16 bit patterns [M, N].
    kind: 'mx' (MXFP4), 'nv' (NVFP4) or 'f8' (MXFP8: aq/bq are e4m3 bytes [rows, K])."""
    import qutlass_b200 as Q
    a = torch.from_numpy(aq).cuda()
    b = torch.from_numpy(bq).cuda()
    if kind == "f8":
        a, b = a.view(torch.float8_e4m3fn), b.view(torch.float8_e4m3fn)
    a_sf = sf_torch(blocked_sf(asf), "mx" if kind == "f8" else kind)
    b_sf = sf_torch(blocked_sf(bsf), "mx" if kind == "f8" else kind)
    al = torch.tensor([alpha], dtype=torch.float32, device="cuda")
    knd = {"mx": Q.KIND_MXF4, "nv": Q.KIND_NVF4, "f8": Q.KIND_MXF8}[kind]
    dt = torch.float8_e4m3fn if kind == "nv" else torch.float8_e8m0fnu
    out = Q._matmul_fp4("test", a, b, a_sf, b_sf, al, knd, dt, 16, cfg=cfg)
    torch.cuda.synchronize()
    return bf16_bits_of(out)


def compare_bits(got: np.ndarray, want: np.ndarray):
    """returns (mismatch_fraction, max_rel_err) between two bf16 bit arrays."""
    g = O.bf16_from_bits(got).astype(np.float64)
    w = O.bf16_from_bits(want).astype(np.float64)
    mism = float((got != want).mean())
    denom = np.maximum(np.abs(w), 1e-30)
    with np.errstate(invalid="ignore"):
        rel = np.abs(g - w) / denom
    rel = np.where(np.isfinite(rel), rel, np.where(g == w, 0.0, np.inf))
    return mism, float(rel.max()) if rel.size else 0.0
