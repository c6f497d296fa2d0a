"""Size-independent properties of the oracle (hypothesis), the same invariants the GPU tests rely on at full size:
the blocked scale layout is a bijection with a closed-form inverse, the e2m1 / e4m3 codecs are idempotent, MX abs-max
quantisation is equivariant under powers of two (scales shift, codes do not move) and the GEMM criterion is linear in alpha
by powers of two."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle as O

SET = settings(max_examples=25, deadline=None)


@SET
@given(rb=st.integers(1, 3), cb=st.integers(1, 5), seed=st.integers(0, 2**31 - 1))
def test_blocked_layout_is_a_bijection_with_inverse(rb, cb, seed):
    rows, cols = rb * 128, cb * 4
    a = np.random.default_rng(seed).integers(0, 256, size=(rows, cols), dtype=np.uint8)
    flat = O.to_blocked(a)
    assert flat.shape == (rows * cols,)
    np.testing.assert_array_equal(O.from_blocked(flat, rows, cols), a)
    r = np.arange(rows)[:, None].repeat(cols, 1)
    c = np.arange(cols)[None, :].repeat(rows, 0)
    off = O.swizzle_offset(r, c, cols)
    assert sorted(off.reshape(-1).tolist()) == list(range(rows * cols))          # every byte lands exactly once
    np.testing.assert_array_equal(flat[off], a)


def test_e2m1_codec_is_idempotent_on_its_grid():
    codes = np.arange(16, dtype=np.uint8)
    vals = O.e2m1_decode(codes)
    again = O.e2m1_decode(O.e2m1_encode(vals))
    np.testing.assert_array_equal(again, vals)                                   # value-level (code 8 = -0 decodes to 0)
    np.testing.assert_array_equal(O.unpack_e2m1(O.pack_e2m1(codes)), codes)


def test_e4m3_codec_is_idempotent_on_its_grid():
    b = np.arange(256, dtype=np.uint8)
    b = b[(b & 0x7F) != 0x7F]                                                    # no NaN encodings
    v = O.e4m3_decode(b).astype(np.float32)
    np.testing.assert_array_equal(O.e4m3_decode(O.e4m3_encode(v)), O.e4m3_decode(b))


@SET
@given(k=st.integers(-20, 20), seed=st.integers(0, 2**31 - 1), had=st.sampled_from([32, 64, 128]))
def test_mx_absmax_is_equivariant_under_powers_of_two(k, seed, had):
    x = O.bf16_round(np.random.default_rng(seed).standard_normal((4, 256)).astype(np.float32) * 25)
    R = O.hadamard_matrix(had)
    a = O.quantize_mx(x, R, "abs_max")
    b = O.quantize_mx(x * np.float32(2.0 ** k), R, "abs_max")
    np.testing.assert_array_equal(b["q"], a["q"])
    np.testing.assert_array_equal(b["sf"].astype(np.int32), a["sf"].astype(np.int32) + k)


@SET
@given(seed=st.integers(0, 2**31 - 1), k=st.integers(-3, 3))
def test_gemm_criterion_is_linear_in_power_of_two_alpha(seed, k):
    rng = np.random.default_rng(seed)
    a = O.e2m1_decode(rng.integers(0, 16, size=(8, 64)).astype(np.uint8)).astype(np.float64)
    b = O.e2m1_decode(rng.integers(0, 16, size=(8, 64)).astype(np.uint8)).astype(np.float64)
    one = O.bf16_from_bits(O.gemm_ref(a, b, 1.0)).astype(np.float64)
    scaled = O.bf16_from_bits(O.gemm_ref(a, b, 2.0 ** k)).astype(np.float64)
    np.testing.assert_array_equal(scaled, one * 2.0 ** k)
