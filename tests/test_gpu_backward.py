"""GPU parity tests of the backward re-quantisers, through the C-ABI, against the CPU
oracle (oracle/bwd_oracle.py), the golden vectors generated from the reference's own test oracles
(tests/golden/backward_vectors.npz <- /root/reference/tests/quartet_test.py) and size-independent properties.

Tolerances: scale bytes and dequantised values may differ from the oracle on at most 1e-4 of the elements for the two
rotating kernels (fp32 summation order at rounding boundaries, the bar tests/mxfp4_test.py:221 sets for the forward
quantiser; the reference's own test demands equality and holds it on randn data, tests/quartet_test.py:224,237-239);
the two MXFP8 re-quantisers are exact integer/byte work: bit-exact (torch.testing.assert_close on fp8 in
tests/quartet_test.py:375-385)."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from oracle import bwd_oracle as B
import helpers as H

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA required", allow_module_level=True)

import qutlass_b200 as Q  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bwd_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "backward_vectors.npz"))


@pytest.fixture(autouse=True, params=["oneshot", "pipelined", "tensorcore", "library-rule"])
def bwd_form(request, b200q_env):
    """every test of this module runs against all forms of the transposing kernels: one CTA per tile (backward.cu), the
    persistent double-buffered form (B200Q_BWD_PIPE=1; same compute and store code: the same bytes) and, for
    backward_t_bf16 / backward_qt_bf16, the tcgen05 kernels (backward_tc.cu, B200Q_BWD_T_TC=1 / B200Q_BWD_QT_TC=1: the tile goes to the tensor core as an MN-major
    operand; same products, fp32 accumulation in the tensor core's order -- inside the same 1e-4 bars, and exactly equal to
    the forward tcgen05 quantiser on the transpose)."""
    if request.param == "library-rule":       # no switch set: the size-dependent choice the library makes on its own
        for name in ("B200Q_BWD_PIPE", "B200Q_BWD_T_TC", "B200Q_BWD_QT_TC"):
            b200q_env(name, None)
        return request.param
    b200q_env("B200Q_BWD_PIPE", "1" if request.param == "pipelined" else "0")
    b200q_env("B200Q_BWD_T_TC", "1" if request.param == "tensorcore" else "0")
    b200q_env("B200Q_BWD_QT_TC", "1" if request.param == "tensorcore" else "0")
    return request.param


def _had():
    return H.bf16_tensor_from_f32(O.hadamard_matrix(32))


def _u8(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ----------------------------------------------------------------------------- backward_t_bf16
def test_backward_t_golden(bwd_golden):
    g = bwd_golden
    x = torch.from_numpy(g["t_x_bits"].view(np.int16)).view(torch.bfloat16).cuda()
    q, sf = Q.backward_t_bf16(x, _had())
    torch.cuda.synchronize()
    assert q.shape == (2, 96, 32) and q.dtype == torch.float4_e2m1fn_x2
    assert sf.shape == (2, 96, 2) and sf.dtype == torch.float8_e8m0fnu
    np.testing.assert_array_equal(H.u8_of(sf), g["t_e8m0"])
    np.testing.assert_array_equal(O.dequant_mx(H.u8_of(q), H.u8_of(sf), 3.0), g["t_dq"])


@pytest.mark.parametrize("shape", [(1, 128, 128), (3, 160, 72), (1, 32, 8), (2, 1024, 2048), (1, 96, 4104)])
def test_backward_t_vs_oracle(shape):
    b, n, m = shape
    x = H.random_bf16(shape, seed=n + m)
    R = O.hadamard_matrix(32)
    ref = B.backward_t_bf16(x, R)
    q, sf = Q.backward_t_bf16(H.bf16_tensor_from_f32(x), _had())
    torch.cuda.synchronize()
    assert tuple(q.shape) == (b, m, n // 2) and tuple(sf.shape) == (b, m, n // 32)
    assert (H.u8_of(sf) != ref["sf"]).mean() <= 1e-4
    dq = O.dequant_mx(H.u8_of(q), H.u8_of(sf))
    assert (dq != O.dequant_mx(ref["q"], ref["sf"])).mean() <= 1e-4
    # the reference's own criterion (tests/quartet_test.py:220-226): scales equal to the fp64 oracle's
    ref64 = B.backward_t_bf16(x, R, "ref64")
    assert (H.u8_of(sf) != ref64["sf"]).mean() <= 1e-4


@pytest.mark.parametrize("rot", ["identity", "random"])
def test_backward_t_generic_rotation(rot):
    """any runtime 32 x 32 matrix is accepted (the reference multiplies by whatever it is given); non-Hadamard
    matrices take the generic fp32 x @ R path."""
    rng = np.random.default_rng(1)
    R = np.eye(32, dtype=np.float32) if rot == "identity" else O.bf16_round(rng.standard_normal((32, 32)) * 0.2)
    x = H.random_bf16((2, 96, 136), seed=9)
    ref = B.backward_t_bf16(x, R)
    q, sf = Q.backward_t_bf16(H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R))
    torch.cuda.synchronize()
    assert (H.u8_of(sf) != ref["sf"]).mean() <= 1e-3
    assert (O.dequant_mx(H.u8_of(q), H.u8_of(sf)) != O.dequant_mx(ref["q"], ref["sf"])).mean() <= 1e-3


def test_backward_t_equals_forward_quantiser_on_the_transpose():
    """size-independent property at a full-size shape: backward_t_bf16(x) == fusedQuantizeMx(x^T, abs_max) -- the forward
    kernel adds 1e-8 to the abs-max before flooring, which cannot change the exponent of randn * 25 data."""
    torch.manual_seed(0)
    x = torch.randn(2, 4096, 4096, dtype=torch.bfloat16, device="cuda") * 25.0
    h = _had()
    q, sf = Q.backward_t_bf16(x, h)
    qf, sff = Q.fusedQuantizeMx(x.transpose(-2, -1).contiguous(), h, method="abs_max")
    torch.cuda.synchronize()
    assert torch.equal(sf.view(torch.uint8).reshape(-1), sff.view(torch.uint8).reshape(-1)[: sf.numel()])
    assert torch.equal(q.view(torch.uint8), qf)
    # graph capture + replay gives the same bytes
    out_q, out_sf = torch.empty_like(q), torch.empty_like(sf)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        Q.backward_t_bf16(x, h, out_q, out_sf)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out_q.view(torch.uint8), q.view(torch.uint8)) and torch.equal(out_sf.view(torch.uint8), sf.view(torch.uint8))


# ----------------------------------------------------------------------------- backward_qt_bf16
def test_backward_qt_golden(bwd_golden):
    g = bwd_golden
    al = torch.tensor([3.0], device="cuda")
    q, sf = Q.backward_qt_bf16(_u8(g["qt_in_e2m1"]), _u8(g["qt_in_e8m0"]).view(torch.float8_e8m0fnu), _had(), al)
    torch.cuda.synchronize()
    assert q.shape == (2, 96, 32) and sf.shape == (2, 96, 2)
    np.testing.assert_array_equal(H.u8_of(sf), g["qt_e8m0"])
    np.testing.assert_array_equal(O.dequant_mx(H.u8_of(q), H.u8_of(sf), 3.0), g["qt_dq"])


@pytest.mark.parametrize("shape,alpha", [((1, 128, 128), 3.0), ((2, 96, 160), 1.0), ((1, 32, 32), 3.0),
                                         ((2, 1024, 1024), 3.0), ((1, 160, 4128), 0.75)])
def test_backward_qt_vs_oracle(shape, alpha):
    b, n, m = shape
    rng = np.random.default_rng(n * 7 + m)
    xq = rng.integers(0, 256, size=(b, n, m // 2), dtype=np.uint8)
    xs = rng.integers(120, 135, size=(b, n, m // 32)).astype(np.uint8)
    R = O.hadamard_matrix(32)
    ref = B.backward_qt_bf16(xq, xs, R, alpha)
    al = torch.tensor([alpha], device="cuda")
    q, sf = Q.backward_qt_bf16(_u8(xq), _u8(xs).view(torch.float8_e8m0fnu), _had(), al)
    torch.cuda.synchronize()
    assert tuple(q.shape) == (b, m, n // 2) and tuple(sf.shape) == (b, m, n // 32)
    assert (H.u8_of(sf) != ref["sf"]).mean() <= 1e-4
    assert (O.dequant_mx(H.u8_of(q), H.u8_of(sf)) != O.dequant_mx(ref["q"], ref["sf"])).mean() <= 1e-4


def test_backward_qt_reference_recipe():
    """the reference's own sequence (tests/quartet_test.py:228-239): abs_max forward quantisation, then
    backward_qt_bf16(alpha = 3) must equal the fp64 oracle of the dequantised transpose EXACTLY (few-bit inputs make
    every fp32 sum exact); plus the property backward_qt(alpha = 1) == backward_t(dequantised bf16)."""
    torch.manual_seed(0)
    x = torch.randn(2, 512, 1024, dtype=torch.bfloat16, device="cuda") * 25.0
    h = _had()
    xq, xs = Q.fusedQuantizeMx(x, h, method="abs_max")
    xs = xs.view(torch.uint8).reshape(-1)[: 2 * 512 * 32].reshape(2, 512, 32).view(torch.float8_e8m0fnu)
    q, sf = Q.backward_qt_bf16(xq, xs, h, torch.tensor([3.0], device="cuda"))
    torch.cuda.synchronize()
    ref = B.backward_qt_bf16(H.u8_of(xq), H.u8_of(xs), O.hadamard_matrix(32), 3.0, "ref64")
    np.testing.assert_array_equal(H.u8_of(sf), ref["sf"])
    np.testing.assert_array_equal(O.dequant_mx(H.u8_of(q), H.u8_of(sf), 3.0), O.dequant_mx(ref["q"], ref["sf"], 3.0))
    # alpha = 1: same as quantising the (exactly representable) dequantised tensor with backward_t_bf16
    dq = torch.from_numpy(O.dequant_mx(H.u8_of(xq), H.u8_of(xs)).astype(np.float32)).cuda().to(torch.bfloat16)
    q1, sf1 = Q.backward_qt_bf16(xq, xs, h, torch.tensor([1.0], device="cuda"))
    q2, sf2 = Q.backward_t_bf16(dq, h)
    torch.cuda.synchronize()
    assert torch.equal(sf1.view(torch.uint8), sf2.view(torch.uint8))
    assert torch.equal(q1.view(torch.uint8), q2.view(torch.uint8))


# ----------------------------------------------------------------------------- backward_bf16_square_double_mxfp8
def test_square_double_golden(bwd_golden):
    g = bwd_golden
    x = torch.from_numpy(g["sq_x_bits"].view(np.int16)).view(torch.bfloat16).cuda()
    y, row, col = Q.backward_bf16_square_double_mxfp8(x)
    torch.cuda.synchronize()
    assert y.dtype == torch.float8_e4m3fn and row.dtype == torch.float8_e8m0fnu and col.dtype == torch.float8_e8m0fnu
    np.testing.assert_array_equal(H.u8_of(y), g["sq_fp8"])
    np.testing.assert_array_equal(H.u8_of(row), g["sq_row"])
    np.testing.assert_array_equal(H.u8_of(col), g["sq_col"])
    # the reference test's own input (tests/quartet_test.py:369-378)
    xa = torch.arange(0, 256, dtype=torch.bfloat16, device="cuda")[None, :].repeat(130, 1)
    y, row, col = Q.backward_bf16_square_double_mxfp8(xa)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(H.u8_of(y), g["sqa_fp8"])
    np.testing.assert_array_equal(H.u8_of(row), g["sqa_row"])
    np.testing.assert_array_equal(H.u8_of(col), g["sqa_col"])


@pytest.mark.parametrize("shape", [(128, 128), (2694, 256), (100, 32), (384, 416), (4096, 4096)])
def test_square_double_vs_oracle(shape):
    m, n = shape
    x = H.random_bf16(shape, seed=m + n, scale=3.0)
    x[: min(m, 32), :32] = 0                        # an all-zero tile
    x[min(m - 1, 40), 40 % n] = 2.0 ** -20          # tiny values: exponent - 7 wraps like the reference's uint8 arithmetic
    want = B.square_double_mxfp8(x)
    got = Q.backward_bf16_square_double_mxfp8(H.bf16_tensor_from_f32(x))
    torch.cuda.synchronize()
    for g_, w_ in zip(got, want):
        assert tuple(g_.shape) == w_.shape
        np.testing.assert_array_equal(H.u8_of(g_), w_)


# ----------------------------------------------------------------------------- mxfp4_transpose_mxfp8
def test_mxfp4_transpose_golden(bwd_golden):
    g = bwd_golden
    y, e = Q.mxfp4_transpose_mxfp8(_u8(g["tr_fp4"]), _u8(g["tr_scales"]).view(torch.float8_e8m0fnu))
    torch.cuda.synchronize()
    assert y.shape == (128, 256) and e.shape == (128, 8)
    np.testing.assert_array_equal(H.u8_of(y), g["tr_fp8"])
    np.testing.assert_array_equal(H.u8_of(e), g["tr_exps"])


@pytest.mark.parametrize("shape", [(256, 256), (2694, 256), (100, 32), (512, 4128), (4096, 4096)])
def test_mxfp4_transpose_vs_oracle(shape):
    m, n = shape
    rng = np.random.default_rng(m + 3 * n)
    fp4 = rng.integers(0, 256, size=(m, n // 2), dtype=np.uint8)
    fp4[: min(m, 32), 0] = 0
    sc = rng.integers(118, 136, size=(m, n // 32)).astype(np.uint8)
    want_q, want_e = B.mxfp4_transpose_mxfp8(fp4, sc)
    y, e = Q.mxfp4_transpose_mxfp8(_u8(fp4), _u8(sc).view(torch.float8_e8m0fnu))
    torch.cuda.synchronize()
    assert tuple(y.shape) == want_q.shape and tuple(e.shape) == want_e.shape
    np.testing.assert_array_equal(H.u8_of(e), want_e)
    np.testing.assert_array_equal(H.u8_of(y), want_q)


def test_fp8_requant_pipeline_reference_recipe():
    """the reference's _fp8_requant_test (tests/quartet_test.py:365-408): bf16 -> square-block MXFP8 (A, used MN-major
    through its COLUMN scales) and MXFP4 -> transposed MXFP8 (B), multiplied with matmul_mxf8_bf16_nn.  The GEMM must be
    bit-exact against the oracle on the dequantised operands and close to bf16.T @ bf16."""
    m, n = 2694, 256
    bf16 = torch.arange(0, n, dtype=torch.bfloat16, device="cuda")[None, :].repeat(m, 1)
    a_fp8, a_row, a_col = Q.backward_bf16_square_double_mxfp8(bf16)              # [2816, 256], col scales [256, 88]
    fp4, scales = Q.fusedQuantizeMx(bf16, torch.eye(32, dtype=torch.bfloat16, device="cuda"), method="abs_max")
    b_fp8, b_exps = Q.mxfp4_transpose_mxfp8(fp4, scales)                           # [256, 2816], [256, 88]
    torch.cuda.synchronize()
    want = B.mxfp4_transpose_mxfp8(H.u8_of(fp4), H.u8_of(scales).reshape(-1)[: m * (n // 32)].reshape(m, n // 32))
    np.testing.assert_array_equal(H.u8_of(b_fp8), want[0])
    np.testing.assert_array_equal(H.u8_of(b_exps), want[1])
    al = torch.tensor([1.0], device="cuda")
    out = Q.matmul_mxf8_bf16_nn(a_fp8, b_fp8, Q.to_blocked(a_col), Q.to_blocked(b_exps), al)
    torch.cuda.synchronize()
    a_dq = O.dequant_mxf8(np.ascontiguousarray(H.u8_of(a_fp8).T), H.u8_of(a_col))   # logical A [M = 256, K = 2816]
    b_dq = O.dequant_mxf8(H.u8_of(b_fp8), H.u8_of(b_exps))                          # [N = 256, K = 2816]
    # abs_max MXFP4 carries the x3 factor: B is 3x the true values
    mism, rel = H.compare_bits(H.bf16_bits_of(out), O.gemm_ref(a_dq, b_dq, 1.0))
    assert rel <= 2.0 ** -7, (mism, rel)
    ref = (bf16.float().T @ bf16.float()) * 3.0
    sim = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=-1).item()
    assert sim > 0.99, sim


def test_backward_ops_registered_and_errors():
    h = _had()
    x = torch.randn(64, 64, dtype=torch.bfloat16, device="cuda")
    q = torch.empty(64, 32, dtype=torch.float4_e2m1fn_x2, device="cuda")
    sf = torch.empty(64, 2, dtype=torch.float8_e8m0fnu, device="cuda")
    torch.ops._qutlass_C.backward_t_bf16(x, h, q, sf)
    q2, sf2 = Q.backward_t_bf16(x, h)
    torch.cuda.synchronize()
    assert torch.equal(q.view(torch.uint8), q2.view(torch.uint8)) and torch.equal(sf.view(torch.uint8), sf2.view(torch.uint8))
    with pytest.raises(RuntimeError, match="multiple of 32"):
        Q.backward_t_bf16(torch.randn(48, 64, dtype=torch.bfloat16, device="cuda"), h,
                          torch.empty(64, 24, dtype=torch.float4_e2m1fn_x2, device="cuda"),
                          torch.empty(64, 1, dtype=torch.float8_e8m0fnu, device="cuda"))
    with pytest.raises(RuntimeError, match="32 x 32"):
        Q.backward_t_bf16(x, torch.eye(64, dtype=torch.bfloat16, device="cuda"))
    with pytest.raises(RuntimeError, match="multiple of 32"):
        Q.backward_bf16_square_double_mxfp8(torch.randn(128, 48, dtype=torch.bfloat16, device="cuda"))
