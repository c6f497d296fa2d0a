"""numpy restatement of the reference's microscaled-FP4 hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Every function cites the
reference file:line (relative to /root/reference) whose behaviour it restates.

Two arithmetic flavours are provided for the quantisers:

* ``arithmetic="ref64"``  -- the float64 emulation the reference's tests use as
  THEIR oracle (tests/mxfp4_test.py:135-184, tests/nvfp4_test.py:132-170).
  This flavour is pinned bit-for-bit against golden vectors generated from those
  very functions (tests/golden/make_golden.py).
* ``arithmetic="kernel"`` -- the fp32 arithmetic of the reference's CUDA
  epilogues (qutlass/csrc/include/cutlass_extensions/epilogue/threadblock/
  epilogue_quant.h:460-812 for MX, :1560-2123 for NV).  The rotation is taken
  as the correctly rounded fp32 value of the exact product (the tensor-core
  summation order is not specified; it differs from this by at most a few ulp
  on rare elements, which is why the reference's tests -- and ours -- accept a
  small mismatch fraction for the quantisers and demand bit-exactness only for
  the GEMM).

The GEMM oracle is ``bf16_rne(alpha * dq(A) @ dq(B)^T)`` in float64
(tests/mxfp4_test.py:229-237).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "E2M1_VALUES", "bf16_round", "bf16_bits", "bf16_from_bits", "hadamard_matrix",
    "e2m1_encode", "e2m1_decode", "pack_e2m1", "unpack_e2m1",
    "e8m0_decode", "e4m3_encode", "e4m3_decode",
    "rotate", "quantize_mx", "quantize_nv", "padded_sf_shape", "sf_rowmajor_padded",
    "to_blocked", "from_blocked", "swizzle_offset", "dequant_mx", "dequant_nv",
    "gemm_ref", "is_sylvester_hadamard", "dequant_mxf8", "pseudoquant_mxfp8",
]

# e2m1 code -> value; code = sign<<3 | {0,.5,1,1.5,2,3,4,6}   (tests/mxfp4_test.py:92-110)
E2M1_VALUES = np.array(
    [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0],
    dtype=np.float64,
)


# --------------------------------------------------------------------------- bf16
def bf16_bits(x) -> np.ndarray:
    """fp32 -> bf16 bit pattern, round-to-nearest-even (torch .to(bfloat16))."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounding = ((u >> 16) & 1) + 0x7FFF
    out = ((u + rounding) >> 16).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, np.uint16(0x7FC0), out)
    return out


def bf16_from_bits(b) -> np.ndarray:
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_round(x) -> np.ndarray:
    """Round fp32/fp64 values to the nearest bf16, returned as float32."""
    x64 = np.asarray(x, dtype=np.float64)
    x32 = x64.astype(np.float32)
    # double rounding f64->f32->bf16 can differ from direct RNE only when the f32
    # rounding lands exactly on a bf16 tie; fix those up exactly.
    r = bf16_from_bits(bf16_bits(x32))
    tie = (x32.view(np.uint32) & 0xFFFF) == 0x8000
    if tie.any():
        lo = bf16_from_bits((x32.view(np.uint32) >> 16).astype(np.uint16))
        hi = bf16_from_bits(((x32.view(np.uint32) >> 16) + 1).astype(np.uint16))
        exact_gt = x64 > x32.astype(np.float64)
        exact_lt = x64 < x32.astype(np.float64)
        toward_hi = np.where(x32 >= 0, exact_gt, exact_lt)
        toward_lo = np.where(x32 >= 0, exact_lt, exact_gt)
        r = np.where(tie & toward_hi, hi, r)
        r = np.where(tie & toward_lo, lo, r)
    return r.astype(np.float32)


def hadamard_matrix(n: int) -> np.ndarray:
    """scipy.linalg.hadamard(n) * n**-0.5 cast to bf16 (tests/mxfp4_test.py:39-42).

    Sylvester construction: H[i, j] = (-1)**popcount(i & j).  Returned as float32
    holding bf16-representable values (32**-.5 -> 0.1767578125, 128**-.5 ->
    0.08837890625)."""
    i = np.arange(n)
    pc = np.zeros((n, n), dtype=np.int64)
    a = i[:, None] & i[None, :]
    while a.any():
        pc += a & 1
        a = a >> 1
    h = np.where(pc & 1, -1.0, 1.0) * (float(n) ** -0.5)
    return bf16_round(h)


def is_sylvester_hadamard(R: np.ndarray) -> bool:
    """True iff R == c * Sylvester-Hadamard for a single scalar c (the structure
    the CUDA kernel's butterfly fast path requires)."""
    n = R.shape[0]
    i = np.arange(n)
    a = i[:, None] & i[None, :]
    pc = np.zeros((n, n), dtype=np.int64)
    while a.any():
        pc += a & 1
        a = a >> 1
    c = R[0, 0]
    return bool(np.array_equal(np.where(pc & 1, -c, c).astype(np.float32), R.astype(np.float32)))


# --------------------------------------------------------------------------- e2m1
_E2M1_MAG = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
# decision thresholds between consecutive magnitudes and which side a tie goes to
# (round-half-to-even on the 1-bit mantissa: even codes are 0,2,4,6 = 0,1,2,4)
_E2M1_MID = np.array([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0])
_E2M1_TIE_UP = np.array([False, True, False, True, False, True, False])


def e2m1_encode(x) -> np.ndarray:
    """float -> e2m1 code (uint8 0..15), RNE, saturating to +-6.

    Restates ``cvt.rn.satfinite.e2m1x2.f32`` (epilogue_quant.h:78-97) and the
    test oracle ``_rtne_fp4`` (tests/mxfp4_test.py:45-81).  The sign bit follows
    the IEEE sign of the input (so -0.0 and tiny negatives give code 8); the
    reference's test oracle instead maps +0.0 to code 8 because of a bucketize
    quirk -- compare decoded values, not raw codes, when zeros are involved."""
    x = np.asarray(x)
    mag = np.abs(x).astype(np.float64)
    code = np.zeros(mag.shape, dtype=np.uint8)
    for mid, tie_up in zip(_E2M1_MID, _E2M1_TIE_UP):
        code += ((mag > mid) | ((mag == mid) & tie_up)).astype(np.uint8)
    sign = np.signbit(x).astype(np.uint8)
    return code | (sign << 3)


def e2m1_decode(code) -> np.ndarray:
    return E2M1_VALUES[np.asarray(code, dtype=np.int64) & 0xF]


def pack_e2m1(code) -> np.ndarray:
    """two codes / byte, element 2i in the LOW nibble (tests/mxfp4_test.py:80)."""
    code = np.asarray(code, dtype=np.uint8)
    return ((code[..., 1::2] & 0xF) << 4) | (code[..., 0::2] & 0xF)


def unpack_e2m1(packed) -> np.ndarray:
    packed = np.asarray(packed, dtype=np.uint8)
    out = np.empty(packed.shape[:-1] + (packed.shape[-1] * 2,), dtype=np.uint8)
    out[..., 0::2] = packed & 0xF
    out[..., 1::2] = packed >> 4
    return out


# --------------------------------------------------------------------------- scales
def e8m0_decode(b) -> np.ndarray:
    """ue8m0 byte -> 2**(b-127) (float64); 0xFF is NaN."""
    b = np.asarray(b, dtype=np.uint8).astype(np.int64)
    v = np.ldexp(1.0, b - 127)
    return np.where(b == 255, np.nan, v)


def e4m3_decode(b) -> np.ndarray:
    """float8_e4m3fn byte -> float64 (max 448, 0x7F/0xFF NaN)."""
    b = np.asarray(b, dtype=np.uint8).astype(np.int64)
    s = np.where(b & 0x80, -1.0, 1.0)
    e = (b >> 3) & 0xF
    m = b & 7
    v = np.where(e == 0, np.ldexp(m / 8.0, -6), np.ldexp(1.0 + m / 8.0, e - 7))
    v = np.where((e == 15) & (m == 7), np.nan, v)
    return s * v


_E4M3_POS = None


def e4m3_encode(x) -> np.ndarray:
    """fp32 -> float8_e4m3fn byte, RNE, saturate-to-finite (448).

    Restates ``__nv_fp8_e4m3(float)`` (epilogue_quant.h:1635,1676) and torch's
    ``.to(torch.float8_e4m3fn)`` for finite in-range inputs (tests/nvfp4_test.py:143)."""
    global _E4M3_POS
    if _E4M3_POS is None:
        _E4M3_POS = e4m3_decode(np.arange(0, 0x7F, dtype=np.uint8))  # 0..448 ascending
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    mag = np.minimum(np.abs(x), 448.0)
    hi = np.searchsorted(_E4M3_POS, mag, side="left").clip(0, len(_E4M3_POS) - 1)
    lo = (hi - 1).clip(0)
    d_lo = mag - _E4M3_POS[lo]
    d_hi = _E4M3_POS[hi] - mag
    pick_hi = (d_hi < d_lo) | ((d_hi == d_lo) & ((hi & 1) == 0))
    code = np.where(pick_hi, hi, lo).astype(np.uint8)
    code = np.where(np.isnan(x), np.uint8(0x7F), code)
    return code | (np.signbit(x).astype(np.uint8) << 7)


# --------------------------------------------------------------------------- rotation
def rotate(x, R, arithmetic: str = "kernel") -> np.ndarray:
    """xh[g, :] = x[g, :] (1xH) @ R (HxH, row-major [k, n]).

    Reference: fused_quantize_mx.cu:86-87 (GEMM [numel/H, H] x [H, H]), test oracle
    tests/mxfp4_test.py:139-142.  ``x`` and ``R`` hold bf16-representable values.
    ``kernel``: exact product rounded once to fp32; ``ref64``: float64."""
    R = np.asarray(R, dtype=np.float64)
    h = R.shape[0]
    x = np.asarray(x, dtype=np.float64)
    xh = (x.reshape(-1, h) @ R).reshape(x.shape)
    if arithmetic == "kernel":
        return xh.astype(np.float32)
    return xh


def _f32(a):
    return np.asarray(a, dtype=np.float32)


# --------------------------------------------------------------------------- MX quantise
def quantize_mx(x, R, method: str = "quest", arithmetic: str = "kernel"):
    """Fused rotate + MXFP4 quantise (group 32, ue8m0 scale).

    kernel flavour restates EpilogueQuantMx::op_32 (epilogue_quant.h:460-575):
      abs_max: s = amax + 1e-8f; floor to 2**e by masking 0x7f800000; q = e2m1(xh / s * 3)
      quest  : sequential fp32 sum / sum-of-squares (with FMA), mean = sum/32,
               var = sumsq/32 - mean*mean, s = var>=0 ? sqrt(var)*(2.92247856/6.)+1e-8 (in
               fp64, rounded to fp32) : 1; same floor; q = e2m1(xh / s)
      mask   : bit i of a uint32 per group = |xh_i / s| < 6   (epilogue_quant.h:1180-1196)
    ref64 flavour restates tests/mxfp4_test.py:135-184.

    Returns dict(q=packed uint8 [numel/2], sf=uint8 [numel/32], mask=uint32 [numel/32],
                 scaled=the pre-rounding scaled values)."""
    assert method in ("quest", "abs_max")
    x = np.asarray(x)
    xh = rotate(x.reshape(-1), R, arithmetic).reshape(-1, 32)
    if arithmetic == "ref64":
        if method == "quest":
            s = xh.std(axis=-1) * (2.92247856 / 6.0) + 1e-8
        else:
            s = np.abs(xh).max(axis=-1) + 1e-8
        e = np.floor(np.log2(s))
        sf = np.clip(e + 127, 0, 254).astype(np.uint8)
        scale = np.ldexp(1.0, sf.astype(np.int64) - 127)
        scaled = xh / scale[:, None]
        if method == "abs_max":
            scaled = scaled * 3
    else:
        xh = _f32(xh)
        if method == "quest":
            s1 = np.zeros(xh.shape[0], dtype=np.float32)
            s2 = np.zeros(xh.shape[0], dtype=np.float32)
            for i in range(32):
                c = xh[:, i]
                s1 = _f32(s1 + c)
                # c_sum2 += c*c contracts to one FMA
                s2 = (c.astype(np.float64) * c.astype(np.float64) + s2.astype(np.float64)).astype(np.float32)
            mean = _f32(s1 / np.float32(32))
            # var = c_sum2/32 - mean*mean  -> fma(-mean, mean, c_sum2/32)
            var = (_f32(s2 / np.float32(32)).astype(np.float64) - mean.astype(np.float64) * mean.astype(np.float64)).astype(np.float32)
            root = np.sqrt(np.maximum(var, 0).astype(np.float32)).astype(np.float32)
            s = np.where(var >= 0, (root.astype(np.float64) * (2.92247856 / 6.0) + 1e-8), 1.0).astype(np.float32)
        else:
            s = _f32(np.abs(xh).max(axis=-1) + np.float32(1e-8))
        bits = s.view(np.uint32) & np.uint32(0x7F800000)
        sf = (bits >> 23).astype(np.uint8)
        scale = bits.view(np.float32)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            scaled = _f32(xh / scale[:, None])
            if method == "abs_max":
                scaled = _f32(scaled * np.float32(3))
    code = e2m1_encode(scaled)
    mask_bits = (np.abs(scaled) < 6.0)
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    mask = (mask_bits.astype(np.uint64) * weights).sum(axis=-1).astype(np.uint32)
    return dict(q=pack_e2m1(code.reshape(-1)), sf=sf, mask=mask, scaled=scaled, xh=xh)


# --------------------------------------------------------------------------- NV quantise
def quantize_nv(x, R, global_scale: float, method: str = "abs_max", arithmetic: str = "kernel", sm100_codes=None):
    """Fused rotate + NVFP4 quantise (group 16, e4m3 scale, fp32 global scale).

    sm100_codes: None (default) = what the reference does ON sm_100, i.e. True exactly for abs_max with a 128 x 128 rotation
    (bindings.cpp:413-415 dispatches that one case to fused_quantize_nv_sm100.cu) and False otherwise; pass False to get
    the mma.sync kernels' / the reference test oracle's arithmetic for that case too (B200Q_NV_ORACLE_CODES in the C-ABI).
    sm100_codes (abs_max only): the arithmetic of the reference's sm_100-ONLY Hadamard-128 kernel
    (sm100_visitor_store_tma_warpspecialized.hpp:141-148,567-591): the stored scale is e4m3-rounded as usual but the codes
    are computed with the UNROUNDED scale, q = e2m1(xh * rcp(sfv * rcp(gs))), sfv = gs * (amax * rcp(6)).  Observed on B200:
    that kernel differs from the one below in 4.8 % of the dequantised values (profiles/r01_ref_quant_diag.jsonl).

    kernel flavour restates EpilogueQuantNv::op_16 (epilogue_quant.h:1604-1692):
      abs_max: SF = e4m3(gs * (amax * rcp(6)));  q = e2m1(xh * rcp(SF * rcp(gs)))  (0 if SF == 0)
      quest  : s = sqrt(sumsq*rcp(16) - mean^2)*(2.92247856/6.)+1e-8; SF = e4m3(s); q = e2m1(xh * rcp(SF))
    (rcp = rcp.approx.ftz, restated here as the correctly rounded reciprocal -- <= 1 ulp apart.)
    ref64 flavour restates tests/nvfp4_test.py:132-170 (abs_max only; equals the kernel
    semantics when global_scale == 6)."""
    assert method in ("quest", "abs_max")
    if sm100_codes is None:
        sm100_codes = method == "abs_max" and np.asarray(R).shape[0] == 128 and arithmetic == "kernel"
    x = np.asarray(x)
    xh = rotate(x.reshape(-1), R, arithmetic).reshape(-1, 16)
    if arithmetic == "ref64":
        assert method == "abs_max"
        s = np.abs(xh).max(axis=-1) + 1e-8
        sf = e4m3_encode(s.astype(np.float32))
        scale = e4m3_decode(sf)
        with np.errstate(divide="ignore", invalid="ignore"):
            scaled = xh / scale[:, None] * 6.0
    else:
        xh = _f32(xh)
        gs = np.float32(global_scale)
        one = np.float32(1)
        if method == "abs_max":
            amax = np.abs(xh).max(axis=-1).astype(np.float32)
            sfv = _f32(gs * _f32(amax * _f32(one / np.float32(6))))
            sf = e4m3_encode(sfv)
            sfq = sfv if sm100_codes else e4m3_decode(sf).astype(np.float32)
            with np.errstate(divide="ignore", invalid="ignore"):
                out_scale = np.where(sfq != 0, _f32(one / _f32(sfq * _f32(one / gs))), np.float32(0)).astype(np.float32)
        else:
            s1 = np.zeros(xh.shape[0], dtype=np.float32)
            s2 = np.zeros(xh.shape[0], dtype=np.float32)
            for i in range(16):
                c = xh[:, i]
                s1 = _f32(s1 + c)
                s2 = (c.astype(np.float64) * c.astype(np.float64) + s2.astype(np.float64)).astype(np.float32)
            r16 = np.float32(1.0 / 16.0)
            mean = _f32(s1 * r16)
            var = (_f32(s2 * r16).astype(np.float64) - mean.astype(np.float64) * mean.astype(np.float64)).astype(np.float32)
            with np.errstate(invalid="ignore"):
                root = np.sqrt(var).astype(np.float32)
            s = (root.astype(np.float64) * (2.92247856 / 6.0) + 1e-8).astype(np.float32)
            sf = e4m3_encode(s)
            sfq = e4m3_decode(sf).astype(np.float32)
            with np.errstate(divide="ignore", invalid="ignore"):
                out_scale = np.where(sfq > 0, _f32(one / sfq), np.float32(0)).astype(np.float32)
        scaled = _f32(xh * out_scale[:, None])
    code = e2m1_encode(scaled)
    return dict(q=pack_e2m1(code.reshape(-1)), sf=sf, scaled=scaled, xh=xh)


# --------------------------------------------------------------------------- SF layout
def padded_sf_shape(rows: int, cols: int):
    """get_padded_shape_mx / _nv (qutlass/utils.py:140-157): pad to (128, 4)."""
    return ((rows + 127) // 128) * 128, ((cols + 3) // 4) * 4


def sf_rowmajor_padded(sf_flat, rows: int, cols: int, fill: int = 0) -> np.ndarray:
    """Place the kernel's flat scale stream into the (padded_rows, padded_cols) buffer
    exactly as the reference kernels do: scale of group g is written at FLAT byte
    offset g of the padded buffer (epilogue_quant.h:509,539 ``D_sf + row``), which is
    row-major [rows, cols] whenever cols % 4 == 0 (every shape the reference tests)."""
    pr, pc = padded_sf_shape(rows, cols)
    out = np.full(pr * pc, fill, dtype=np.uint8)
    out[: rows * cols] = np.asarray(sf_flat, dtype=np.uint8).reshape(-1)
    return out.reshape(pr, pc)


def swizzle_offset(r, c, padded_cols: int):
    """Byte offset of scale (r, c) in the block-scaled ("to_blocked") layout:
    128-row x 4-col blocks of 512 B, K-blocks fastest (qutlass/utils.py:178-193,
    triton_scale_swizzle :16-77)."""
    r = np.asarray(r, dtype=np.int64)
    c = np.asarray(c, dtype=np.int64)
    return ((r // 128) * (padded_cols // 4) + c // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + (c % 4)


def to_blocked(sf2d) -> np.ndarray:
    """qutlass.utils.to_blocked (torch path, utils.py:178-193) on a padded uint8 matrix."""
    sf2d = np.asarray(sf2d, dtype=np.uint8)
    rows, cols = sf2d.shape
    assert rows % 128 == 0 and cols % 4 == 0
    nrb, ncb = rows // 128, cols // 4
    blocks = sf2d.reshape(nrb, 128, ncb, 4).transpose(0, 2, 1, 3)
    rearranged = blocks.reshape(-1, 4, 32, 4).transpose(0, 2, 1, 3).reshape(-1, 32, 16)
    return np.ascontiguousarray(rearranged).reshape(-1)


def from_blocked(flat, rows: int, cols: int) -> np.ndarray:
    """inverse of to_blocked."""
    nrb, ncb = rows // 128, cols // 4
    a = np.asarray(flat, dtype=np.uint8).reshape(nrb * ncb, 32, 4, 4).transpose(0, 2, 1, 3)
    a = a.reshape(nrb, ncb, 128, 4).transpose(0, 2, 1, 3)
    return np.ascontiguousarray(a).reshape(rows, cols)


# --------------------------------------------------------------------------- dequant / GEMM
def dequant_mx(q_packed, sf, alpha: float = 1.0) -> np.ndarray:
    """_dq_fp4 (tests/mxfp4_test.py:84-120): q [..., K/2] bytes, sf [..., K/32] ue8m0 bytes."""
    vals = e2m1_decode(unpack_e2m1(q_packed))
    s = e8m0_decode(sf)
    shp = vals.shape
    return (vals.reshape(shp[:-1] + (-1, 32)) * s[..., None]).reshape(shp) / alpha


def dequant_nv(q_packed, sf, alpha: float = 1.0) -> np.ndarray:
    """_dq_fp4 (tests/nvfp4_test.py:81-117): sf [..., K/16] e4m3 bytes."""
    vals = e2m1_decode(unpack_e2m1(q_packed))
    s = e4m3_decode(sf)
    shp = vals.shape
    return (vals.reshape(shp[:-1] + (-1, 16)) * s[..., None]).reshape(shp) / alpha


def gemm_ref(a_dq, b_dq, alpha: float = 1.0) -> np.ndarray:
    """bf16_rne(alpha * a_dq @ b_dq^T) -- the reference's bit-exact GEMM criterion
    (tests/mxfp4_test.py:229-237: float64 matmul of dequantised operands, .to(bf16)).
    Returns the bf16 BIT PATTERNS (uint16)."""
    acc = np.asarray(a_dq, dtype=np.float64) @ np.asarray(b_dq, dtype=np.float64).T
    acc = acc * float(alpha)
    return bf16_bits(bf16_round(acc))


# --------------------------------------------------------------------------- MXFP8 ("next" row: matmul_mxf8_bf16_tn)
def dequant_mxf8(q_e4m3, sf) -> np.ndarray:
    """e4m3 bytes [..., K] * ue8m0 scale per 32 -> float64 (tests/mxfp8_test.py:42 `xq * shared_exps`)."""
    vals = e4m3_decode(q_e4m3)
    s = e8m0_decode(sf)
    shp = vals.shape
    return (vals.reshape(shp[:-1] + (-1, 32)) * s[..., None]).reshape(shp)


def pseudoquant_mxfp8(x):
    """_pseudoquant_mxfp8 (tests/mxfp8_test.py:27-46): shared exponent = floor(log2(amax)) - 8 (+127 bias, 2**0 when
    the group is all zero), values = e4m3(clamp(x / 2**e, +-448)).  x holds bf16-representable values; the division
    by a power of two is exact and the bf16 quotient converts to e4m3 with one RNE."""
    x = np.asarray(x, dtype=np.float64)
    g = x.reshape(-1, 32)
    amax = np.abs(g).max(axis=-1)
    with np.errstate(divide="ignore"):
        e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) - 8 + 128, 128).astype(np.int64)
    sf = (e & 0xFF).astype(np.uint8)       # the test stores the biased value as uint8 and views it as e8m0
    scale = e8m0_decode(sf)
    q = e4m3_encode(np.clip(g / scale[:, None], -448.0, 448.0).astype(np.float32))
    return q.reshape(x.shape), sf.reshape(x.shape[:-1] + (x.shape[-1] // 32,))
