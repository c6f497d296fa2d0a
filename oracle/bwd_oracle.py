"""numpy restatement of the reference's backward-pass re-quantisers (CPU oracle, "next" row 4 of SURVEY.md section 8f).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Each function cites the reference file:line (relative to
/root/reference) it restates.  Pinned bit-for-bit against golden vectors generated from the reference's own test
helpers (tests/quartet_test.py ``_backward_quantize_ref``, ``_backward_bf16_square_double_mxfp8``,
``_mxfp4_transpose_mxfp8``) by tests/golden/make_golden.py -> tests/golden/backward_vectors.npz.
"""
from __future__ import annotations

import numpy as np

from .fp4_oracle import (e2m1_decode, e2m1_encode, e4m3_encode, e8m0_decode, pack_e2m1, rotate, unpack_e2m1)

__all__ = ["backward_t_bf16", "backward_qt_bf16", "square_double_mxfp8", "mxfp4_transpose_mxfp8", "dequant_e8m0_shl7"]


def _f32(a):
    return np.asarray(a, dtype=np.float32)


def _quantize_absmax_groups(xh, arithmetic: str, alpha: float | None):
    """abs-max MXFP4 quantisation of rotated 32-groups xh [G, 32].

    kernel flavour (quartet_bwd_sm120.cu:303-315 / :397-410, fp32):
        s = amax [/ alpha];  s_bits &= 0x7f800000;  byte = s_bits >> 23;  q = e2m1(xh * (3 / (s [* alpha])))
      a group whose floored scale is 0 gives byte 0 and all-zero codes (what the reference's TEST oracle yields; the
      reference kernel divides by zero there).
    ref64 flavour (tests/quartet_test.py:155-175 _backward_quantize_ref, float64):
        s = 2^floor(log2(amax)) -> e8m0;  q = rtne_fp4(xh / s * 3)     (alpha: caller divides the input by alpha)
    """
    if arithmetic == "ref64":
        xh = np.asarray(xh, dtype=np.float64)
        amax = np.abs(xh).max(axis=-1)
        with np.errstate(divide="ignore"):
            e = np.floor(np.log2(amax))
        sf = np.where(amax > 0, np.clip(e + 127, 0, 254), 0).astype(np.uint8)
        scale = np.ldexp(1.0, sf.astype(np.int64) - 127)
        scaled = xh / scale[:, None] * 3.0
    else:
        xh = _f32(xh)
        amax = np.abs(xh).max(axis=-1).astype(np.float32)
        s = amax if alpha is None else _f32(amax / np.float32(alpha))
        bits = s.view(np.uint32) & np.uint32(0x7F800000)
        sf = (bits >> 23).astype(np.uint8)
        sp = bits.view(np.float32)
        den = sp if alpha is None else _f32(sp * np.float32(alpha))
        with np.errstate(divide="ignore", invalid="ignore"):
            f = np.where((sf == 0) | (sf == 255), np.float32(0), _f32(np.float32(3) / den)).astype(np.float32)
        scaled = _f32(xh * f[:, None])
    code = e2m1_encode(scaled)
    return pack_e2m1(code.reshape(-1)), sf, scaled


def backward_t_bf16(x, R, arithmetic: str = "kernel"):
    """backward_t_bf16 (qutlass/__init__.py:206-244; quartet_bwd_sm120.cu:237-318): MXFP4 abs-max quantisation of
    rotate(x^T).  x [..., N, M] (bf16-representable values), R [32, 32].
    Returns dict(q=uint8 [..., M, N/2], sf=uint8 [..., M, N/32], scaled=[..., M, N])."""
    x = np.asarray(x)
    xt = np.swapaxes(x, -1, -2)                      # [..., M, N]
    shp = xt.shape
    xh = rotate(np.ascontiguousarray(xt).reshape(-1), R, arithmetic).reshape(-1, 32)
    q, sf, scaled = _quantize_absmax_groups(xh, arithmetic, None)
    return dict(q=q.reshape(shp[:-1] + (shp[-1] // 2,)), sf=sf.reshape(shp[:-1] + (shp[-1] // 32,)),
                scaled=scaled.reshape(shp))


def dequant_e8m0_shl7(q_packed, sf) -> np.ndarray:
    """the dequantisation inside quantize_g32qt_cuda_kernel (quartet_bwd_sm120.cu:357-365): bf16(code) * bf16 whose
    bits are (scale byte << 7), i.e. 2^(e-127) with e = 0 -> 0.0 (and 255 -> inf).  float64, exact."""
    vals = e2m1_decode(unpack_e2m1(q_packed))
    b = np.asarray(sf, dtype=np.uint8).astype(np.int64)
    with np.errstate(over="ignore"):
        s = np.where(b == 0, 0.0, np.where(b == 255, np.inf, np.ldexp(1.0, b - 127)))
    shp = vals.shape
    return (vals.reshape(shp[:-1] + (-1, 32)) * s[..., None]).reshape(shp)


def backward_qt_bf16(q_packed, sf, R, alpha: float, arithmetic: str = "kernel"):
    """backward_qt_bf16 (qutlass/__init__.py:247-283; quartet_bwd_sm120.cu:320-412): dequantise MXFP4
    x [..., N, M/2] / [..., N, M/32], transpose, rotate, re-quantise with s = floor_pow2(amax / alpha),
    q = e2m1(xh * 3 / (s * alpha)).
    ref64 (tests/quartet_test.py:228-239): _backward_quantize_ref(_dq_fp4(x, alpha).T), i.e. the input divided by alpha
    BEFORE the rotation, then the plain abs-max rule."""
    if arithmetic == "ref64":
        dq = (e2m1_decode(unpack_e2m1(q_packed)).reshape(np.asarray(sf).shape + (32,)) *
              e8m0_decode(sf)[..., None]).reshape(np.asarray(q_packed).shape[:-1] + (-1,)) / float(alpha)
    else:
        dq = dequant_e8m0_shl7(q_packed, sf)
    xt = np.swapaxes(dq, -1, -2)
    shp = xt.shape
    xh = rotate(np.ascontiguousarray(xt).reshape(-1), R, arithmetic).reshape(-1, 32)
    q, osf, scaled = _quantize_absmax_groups(xh, arithmetic, None if arithmetic == "ref64" else alpha)
    return dict(q=q.reshape(shp[:-1] + (shp[-1] // 2,)), sf=osf.reshape(shp[:-1] + (shp[-1] // 32,)),
                scaled=scaled.reshape(shp))


def _e8m0_shift7(amax) -> np.ndarray:
    """encode_e8m0_shiftm8 (quartet_bwd_sm120.cu:497-503) == tests/quartet_test.py:279-285: biased exponent of amax
    minus 7 (mod 256), 127 for amax == 0."""
    a = _f32(amax)
    e = (a.view(np.uint32) >> 23) & 0xFF
    return np.where(a == 0, 127, (e.astype(np.int64) - 7) & 0xFF).astype(np.uint8)


def _to_e4m3_scaled(x, sf_bytes) -> np.ndarray:
    """e4m3_satfinite(bf16(x / 2^(byte-127))): the division by a power of two is exact (quartet_bwd_sm120.cu:582-588)."""
    s = e8m0_decode(sf_bytes)
    with np.errstate(over="ignore", invalid="ignore"):
        q = np.asarray(x, dtype=np.float64) / s
    return e4m3_encode(q.astype(np.float32))


def square_double_mxfp8(x):
    """backward_bf16_square_double_mxfp8 (qutlass/__init__.py:285-294; quartet_bwd_sm120.cu:505-623;
    tests/quartet_test.py:264-291): pad rows to a multiple of 128 with zeros; ONE ue8m0 scale per 32 x 32 tile
    (exponent(amax) - 7); x_fp8 = e4m3(x / scale); row_scales [m_pad, n/32] and column_scales [n, m_pad/32] repeat the
    tile scale for each of the 32 rows / columns.  Returns (x_fp8 u8 [m_pad, n], row_scales u8, column_scales u8)."""
    x = np.asarray(x, dtype=np.float32)
    m, n = x.shape
    m_pad = (m + 127) // 128 * 128
    xp = np.zeros((m_pad, n), dtype=np.float32)
    xp[:m] = x
    t = xp.reshape(m_pad // 32, 32, n // 32, 32)
    amax = np.abs(t).max(axis=(1, 3))
    e = _e8m0_shift7(amax)                                           # [m_pad/32, n/32]
    q = _to_e4m3_scaled(t, e[:, None, :, None]).reshape(m_pad, n)
    row = np.repeat(e, 32, axis=0)                                   # [m_pad, n/32]
    col = np.repeat(np.ascontiguousarray(e.T), 32, axis=0)           # [n, m_pad/32]
    return q, row, col


def mxfp4_transpose_mxfp8(q_packed, sf, m: int | None = None):
    """mxfp4_transpose_mxfp8 (qutlass/__init__.py:296-309; quartet_bwd_sm120.cu:627-734; tests/quartet_test.py:294-345):
    dequantise MXFP4 x [m, n/2] / scales [>= m, n/32] (bf16, exact), pad rows to a multiple of 256 with zeros, transpose,
    ue8m0 scale per 32 along m (exponent(amax) - 7), x_fp8 = e4m3(x^T / scale).
    Returns (x_fp8 u8 [n, m_pad], shared_exps u8 [n, m_pad/32])."""
    q_packed = np.asarray(q_packed, dtype=np.uint8)
    if m is None:
        m = q_packed.shape[0]
    n = q_packed.shape[1] * 2
    m_pad = (m + 255) // 256 * 256
    sf = np.asarray(sf, dtype=np.uint8).reshape(-1, n // 32)[:m]
    dq = np.zeros((m_pad, n), dtype=np.float64)
    vals = e2m1_decode(unpack_e2m1(q_packed[:m]))
    dq[:m] = (vals.reshape(m, n // 32, 32) * e8m0_decode(sf)[..., None]).reshape(m, n)
    xt = np.ascontiguousarray(dq.T)                                   # [n, m_pad]
    g = xt.reshape(n, m_pad // 32, 32)
    e = _e8m0_shift7(np.abs(g).max(axis=-1))
    q = _to_e4m3_scaled(g, e[..., None]).reshape(n, m_pad)
    return q, e
