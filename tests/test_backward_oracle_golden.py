"""Pin the backward re-quantiser oracle (oracle/bwd_oracle.py) against vectors produced by the reference's own test
oracles (tests/golden/make_golden.py backward ran /root/reference/tests/quartet_test.py code on CPU)."""
import os

import numpy as np
import pytest

import oracle as O
from oracle import bwd_oracle as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bwd_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "backward_vectors.npz"))


def _rot(g):
    return O.bf16_from_bits(g["had32_bits"]).astype(np.float32)


@pytest.mark.parametrize("arith", ["ref64", "kernel"])
def test_backward_t_matches_reference_oracle(bwd_golden, arith):
    g = bwd_golden
    r = B.backward_t_bf16(O.bf16_from_bits(g["t_x_bits"]), _rot(g), arith)
    np.testing.assert_array_equal(r["sf"], g["t_e8m0"])
    # codes: the reference's _rtne_fp4 encodes +0.0 as code 8 (bucketize quirk) -> compare dequantised values
    np.testing.assert_array_equal(O.dequant_mx(r["q"], r["sf"], 3.0), g["t_dq"])
    nz = (O.unpack_e2m1(r["q"]) & 7) != 0
    np.testing.assert_array_equal(O.unpack_e2m1(r["q"])[nz], O.unpack_e2m1(g["t_e2m1"])[nz])


@pytest.mark.parametrize("arith", ["ref64", "kernel"])
def test_backward_qt_matches_reference_oracle(bwd_golden, arith):
    g = bwd_golden
    r = B.backward_qt_bf16(g["qt_in_e2m1"], g["qt_in_e8m0"], _rot(g), 3.0, arith)
    np.testing.assert_array_equal(r["sf"], g["qt_e8m0"])
    np.testing.assert_array_equal(O.dequant_mx(r["q"], r["sf"], 3.0), g["qt_dq"])


def test_square_double_mxfp8_matches_reference_oracle(bwd_golden):
    g = bwd_golden
    q, row, col = B.square_double_mxfp8(O.bf16_from_bits(g["sq_x_bits"]))
    np.testing.assert_array_equal(q, g["sq_fp8"])
    np.testing.assert_array_equal(row, g["sq_row"])
    np.testing.assert_array_equal(col, g["sq_col"])
    # the reference test's own input (tests/quartet_test.py:369-378): arange(0, 256) rows
    xa = np.tile(np.arange(256, dtype=np.float32)[None, :], (130, 1))
    q, row, col = B.square_double_mxfp8(xa)
    np.testing.assert_array_equal(q, g["sqa_fp8"])
    np.testing.assert_array_equal(row, g["sqa_row"])
    np.testing.assert_array_equal(col, g["sqa_col"])


def test_mxfp4_transpose_mxfp8_matches_reference_oracle(bwd_golden):
    g = bwd_golden
    q, e = B.mxfp4_transpose_mxfp8(g["tr_fp4"], g["tr_scales"])
    np.testing.assert_array_equal(q, g["tr_fp8"])
    np.testing.assert_array_equal(e, g["tr_exps"])


def test_zero_group_and_pow2_boundary():
    R = O.hadamard_matrix(32)
    x = np.zeros((32, 8), dtype=np.float32)
    r = B.backward_t_bf16(x, R)
    assert (r["sf"] == 0).all() and (r["q"] == 0).all()
    q, row, col = B.square_double_mxfp8(np.zeros((32, 32), dtype=np.float32))
    assert (row == 127).all() and (col == 127).all() and (q == 0).all() and q.shape == (128, 32)
