"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI
(libb200q.so via qutlass_b200), against the CPU oracle on identical seeded inputs, against the
committed golden fixtures generated from the reference's own test helpers, and -- at BASELINE.json's
full sizes -- through size-independent properties.

Tolerances (written here, from north_star and the reference's own tests):
  * GEMM: bit-exact vs bf16(fp64 matmul of the dequantised operands) wherever fp32 accumulation is exact
    (tests/mxfp4_test.py:237 `out.equal(...)`); otherwise |err| <= 2**-7 relative (north_star).
  * quantise (floating point, rounding-boundary sensitive): mismatch fraction of dequantised values
    <= 1e-4 for MX (tests/mxfp4_test.py:221) and <= 1e-2 for NV (reference bar is 1e-1, nvfp4_test.py:205).
"""
import numpy as np
import pytest
import torch

import oracle as O
import helpers as H

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA required", allow_module_level=True)

import qutlass_b200 as Q  # noqa: E402
from qutlass_b200 import _lib  # noqa: E402

REL_TOL = 2.0 ** -7


def _flat_sf(sf_t, rows, cols):
    return H.u8_of(sf_t).reshape(-1)[: rows * cols].reshape(rows, cols)


# ----------------------------------------------------------------------------- quantise
@pytest.mark.parametrize("had", [32, 64, 128])
@pytest.mark.parametrize("method", ["abs_max", "quest"])
def test_quantize_mx_vs_oracle(had, method):
    rows, k = 384, 2048
    x = H.random_bf16((rows, k), seed=had)
    R = O.hadamard_matrix(had)
    ref = O.quantize_mx(x, R, method)
    out = Q.fusedQuantizeMx(H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R), method=method,
                            return_mask=(method == "quest"))
    torch.cuda.synchronize()
    cols = k // 32
    assert out[1].shape == (384, 64) and out[1].dtype == torch.float8_e8m0fnu
    assert out[0].shape == (rows, k // 2) and out[0].dtype == torch.uint8
    sf = _flat_sf(out[1], rows, cols)
    assert (sf != ref["sf"].reshape(rows, cols)).mean() <= 1e-4
    dq = O.dequant_mx(H.u8_of(out[0]), sf)
    dq_ref = O.dequant_mx(ref["q"].reshape(rows, -1), ref["sf"].reshape(rows, cols))
    assert (dq != dq_ref).mean() <= 1e-4
    if method == "quest":
        mask = H.u8_of(out[2]).reshape(-1).view(np.uint32)
        assert (mask != ref["mask"]).mean() <= 1e-4
    # the blocked copy written by the same kernel == to_blocked(row-major) (no-op path) == CUDA swizzle kernel
    blk_noop = H.u8_of(Q.to_blocked(out[1]))
    np.testing.assert_array_equal(blk_noop, H.blocked_sf(sf))
    fresh = out[1].clone()          # loses the attached buffer -> runs the swizzle kernel
    np.testing.assert_array_equal(H.u8_of(Q.to_blocked(fresh, use_triton_kernel=True)), blk_noop)


@pytest.mark.parametrize("had", [16, 32, 64, 128])
@pytest.mark.parametrize("method", ["abs_max", "quest"])
@pytest.mark.parametrize("gs", [1.0, 6.0])
def test_quantize_nv_vs_oracle(had, method, gs):
    rows, k = 256, 1024
    x = H.random_bf16((rows, k), seed=100 + had)
    R = O.hadamard_matrix(had)
    ref = O.quantize_nv(x, R, gs, method)
    gst = torch.tensor([gs], dtype=torch.float32, device="cuda")
    q, sf_t = Q.fusedQuantizeNv(H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R), gst, method=method)
    torch.cuda.synchronize()
    cols = k // 16
    assert sf_t.dtype == torch.float8_e4m3fn and sf_t.shape == (256, 64)
    sf = _flat_sf(sf_t, rows, cols)
    assert (sf != ref["sf"].reshape(rows, cols)).mean() <= 1e-3
    dq = O.dequant_nv(H.u8_of(q), sf)
    dq_ref = O.dequant_nv(ref["q"].reshape(rows, -1), ref["sf"].reshape(rows, cols))
    assert (dq != dq_ref).mean() <= 1e-2
    np.testing.assert_array_equal(H.u8_of(Q.to_blocked(sf_t)), H.blocked_sf(sf))


@pytest.mark.parametrize("had", [32, 64, 128])
@pytest.mark.parametrize("method", ["quest", "abs_max"])
def test_quantize_mx_golden_reference_vectors(golden, had, method):
    """GPU kernel vs the reference's fp64 test oracle output (golden), reference's own bar."""
    x = O.bf16_from_bits(golden["mx_x_bits"])
    R = O.bf16_from_bits(golden[f"had{had}_bits"])
    tag = f"mx_h{had}_{'quest' if method == 'quest' else 'absmax'}"
    q, sf_t = Q.fusedQuantizeMx(H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R), method=method)
    torch.cuda.synchronize()
    rows, k = x.shape
    dq = O.dequant_mx(H.u8_of(q), _flat_sf(sf_t, rows, k // 32), alpha=1.0 if method == "quest" else 3.0)
    assert (dq != golden[tag + "_dq"]).mean() <= 1e-4


@pytest.mark.parametrize("had", [16, 32, 64, 128])
def test_quantize_nv_golden_reference_vectors(golden, had, monkeypatch):
    x = O.bf16_from_bits(golden["mx_x_bits"])
    R = O.bf16_from_bits(golden[f"had{had}_bits"])
    gst = torch.tensor([6.0], dtype=torch.float32, device="cuda")
    q, sf_t = Q.fusedQuantizeNv(H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R), gst)
    torch.cuda.synchronize()
    rows, k = x.shape
    dq = O.dequant_nv(H.u8_of(q), _flat_sf(sf_t, rows, k // 16), alpha=6.0)
    # golden = the reference's fp64 TEST ORACLE.  For Hadamard-128 the default follows the reference's sm_100 KERNEL, which
    # sits 4.8 % from that oracle (the reference's own bar there is 1e-1, tests/nvfp4_test.py:204-205); the oracle-arithmetic
    # switch brings it back under 1e-2 like every other size
    assert (dq != golden[f"nv_h{had}_dq"]).mean() <= (1e-1 if had == 128 else 1e-2)
    if had == 128:
        monkeypatch.setenv("B200Q_NV128_ORACLE_CODES", "1")
        q, sf_t = Q.fusedQuantizeNv(H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R), gst)
        torch.cuda.synchronize()
        dq = O.dequant_nv(H.u8_of(q), _flat_sf(sf_t, rows, k // 16), alpha=6.0)
        assert (dq != golden[f"nv_h{had}_dq"]).mean() <= 1e-2


def test_quantize_edge_cases():
    R32 = O.hadamard_matrix(32)
    Rt = H.bf16_tensor_from_f32(R32)
    # all-zero input: scale floor(1e-8) and all-zero codes; single row (M=1); leading batch dims; ragged K=96
    z = torch.zeros(1, 4096, dtype=torch.bfloat16, device="cuda")
    q, sf = Q.fusedQuantizeMx(z, Rt, method="abs_max")
    torch.cuda.synchronize()
    assert sf.shape == (128, 128)
    ref = O.quantize_mx(np.zeros((1, 4096), np.float32), R32, "abs_max")
    np.testing.assert_array_equal(_flat_sf(sf, 1, 128), ref["sf"].reshape(1, 128))
    assert (O.e2m1_decode(O.unpack_e2m1(H.u8_of(q))) == 0).all()
    # padding of the blocked buffer is zero-filled (pad rows never produce NaN scales)
    blk = H.u8_of(Q.to_blocked(sf))
    want = H.blocked_sf(ref["sf"].reshape(1, 128))
    np.testing.assert_array_equal(blk, want)
    # batch dims flattened like the reference (utils.py:141)
    x = H.random_bf16((2, 5, 96), seed=5)
    q, sf = Q.fusedQuantizeMx(H.bf16_tensor_from_f32(x), Rt, method="quest")
    torch.cuda.synchronize()
    assert q.shape == (2, 5, 48) and sf.shape == (128, 4)
    ref = O.quantize_mx(x, R32, "quest")
    dq = O.dequant_mx(H.u8_of(q).reshape(10, 48), _flat_sf(sf, 10, 3))
    assert (dq != O.dequant_mx(ref["q"].reshape(10, 48), ref["sf"].reshape(10, 3))).mean() <= 1e-4
    # arbitrary (non-Hadamard) runtime rotation: identity, as quartet_test.py:380 passes
    eye = torch.eye(32, dtype=torch.bfloat16, device="cuda")
    x = H.random_bf16((64, 256), seed=6)
    q, sf = Q.fusedQuantizeMx(H.bf16_tensor_from_f32(x), eye, method="abs_max")
    torch.cuda.synchronize()
    ref = O.quantize_mx(x, np.eye(32, dtype=np.float32), "abs_max")
    dq = O.dequant_mx(H.u8_of(q), _flat_sf(sf, 64, 8))
    np.testing.assert_array_equal(dq, O.dequant_mx(ref["q"].reshape(64, 128), ref["sf"].reshape(64, 8)))


@pytest.mark.parametrize("kind", ["mx", "nv"])
@pytest.mark.parametrize("had", [32, 64, 128])
def test_quantize_device_check_equals_trusted_hint(kind, had):
    """C-ABI: with B200Q_ROT_TRUSTED_HADAMARD the kernel skips its device-side structure check; without it the
    kernel verifies R itself.  Both must give identical bytes (the Python API passes the hint after an exact,
    cached host-side inspection of R; a CUDA-graph capture of an unseen R takes the device-check path)."""
    lib = _lib.load()
    rows, k = 200, 1024
    x = H.bf16_tensor_from_f32(H.random_bf16((rows, k), seed=had + 7))
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(had))
    group = 32 if kind == "mx" else 16
    outs = []
    st = torch.cuda.current_stream().cuda_stream
    gs = torch.tensor([1.0], device="cuda")
    for flag in (0, Q.ROT_TRUSTED_HADAMARD):
        q = torch.zeros(rows, k // 2, dtype=torch.uint8, device="cuda")
        sf = torch.zeros(256 * (k // group), dtype=torch.uint8, device="cuda")
        blk = torch.zeros(256 * (k // group), dtype=torch.uint8, device="cuda")
        if kind == "mx":
            rc = lib.b200q_quantize_mx(x.data_ptr(), R.data_ptr(), q.data_ptr(), sf.data_ptr(), blk.data_ptr(), None,
                                       rows * k, k, had, Q.METHOD_QUEST | flag, st)
        else:
            rc = lib.b200q_quantize_nv(x.data_ptr(), R.data_ptr(), q.data_ptr(), sf.data_ptr(), blk.data_ptr(),
                                       gs.data_ptr(), rows * k, k, had, Q.METHOD_ABSMAX | flag, st)
        assert rc == 0, lib.b200q_last_error()
        torch.cuda.synchronize()
        outs.append((q.cpu(), sf.cpu(), blk.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert Q._rotation_hint(R) == Q.ROT_TRUSTED_HADAMARD
    assert Q._rotation_hint(torch.eye(had, dtype=torch.bfloat16, device="cuda")) == Q.ROT_GENERIC


@pytest.mark.parametrize("kind", ["mx", "nv"])
@pytest.mark.parametrize("had", [32, 64, 128])
def test_quantize_arbitrary_rotation_on_tensor_cores(kind, had):
    """a non-Hadamard runtime rotation (the API allows any matrix) takes the mma.sync kernel via the Python hint and the
    butterfly kernel's scalar fallback without it; both must agree with the oracle."""
    rows, k = 96, 1024
    x = H.random_bf16((rows, k), seed=had)
    R = O.bf16_round(np.random.default_rng(had).standard_normal((had, had)).astype(np.float32) * had ** -0.5)
    xt, Rt = H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R)
    if kind == "mx":
        ref = O.quantize_mx(x, R, "abs_max")
        q, sf = Q.fusedQuantizeMx(xt, Rt, method="abs_max")
        dq = O.dequant_mx(H.u8_of(q), _flat_sf(sf, rows, k // 32))
        dq_ref = O.dequant_mx(ref["q"].reshape(rows, -1), ref["sf"].reshape(rows, -1))
    else:
        ref = O.quantize_nv(x, R, 1.0, "abs_max")
        q, sf = Q.fusedQuantizeNv(xt, Rt, torch.tensor([1.0], device="cuda"))
        dq = O.dequant_nv(H.u8_of(q), _flat_sf(sf, rows, k // 16))
        dq_ref = O.dequant_nv(ref["q"].reshape(rows, -1), ref["sf"].reshape(rows, -1))
    torch.cuda.synchronize()
    assert (dq != dq_ref).mean() <= (1e-3 if kind == "mx" else 1e-2)


@pytest.mark.parametrize("kind,had", [("mx", 32), ("mx", 64), ("mx", 128), ("nv", 16), ("nv", 32), ("nv", 64), ("nv", 128)])
@pytest.mark.parametrize("method", ["abs_max", "quest"])
@pytest.mark.parametrize("shape", [(384, 2048), (132, 160), (1000, 4096)])
def test_quantize_tcgen05_kernel_vs_oracle(kind, had, method, shape, b200q_env):
    """The tcgen05 rotation kernel (quantize_tc.cu; default for large inputs, forced here with B200Q_QUANT_TC=1) against
    the oracle: codes, row-major scales, the blocked copy and the clip mask.  (132, 160): 5 / 10 scales per row, so a
    128-element row of the kernel's flat view straddles matrix rows (byte-granular blocked stores) and the last tile
    is partial; (1000, 4096): 8 tiles per CTA-round with a ragged row count."""
    b200q_env("B200Q_QUANT_TC", "1")
    rows, k = shape
    if (rows * k) % had:
        pytest.skip("numel % H")
    x = H.random_bf16((rows, k), seed=had + rows)
    R = O.hadamard_matrix(had)
    xt, Rt = H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R)
    if kind == "mx":
        group, tol_sf, tol = 32, 1e-4, 1e-4
        ref = O.quantize_mx(x, R, method)
        out = Q.fusedQuantizeMx(xt, Rt, method=method, return_mask=(method == "quest"))
        deq = O.dequant_mx
    else:
        group, tol_sf, tol = 16, 1e-3, 1e-2
        ref = O.quantize_nv(x, R, 6.0, method)
        out = Q.fusedQuantizeNv(xt, Rt, torch.tensor([6.0], device="cuda"), method=method)
        deq = O.dequant_nv
    torch.cuda.synchronize()
    cols = k // group
    sf = _flat_sf(out[1], rows, cols)
    assert (sf != ref["sf"].reshape(rows, cols)).mean() <= tol_sf
    dq = deq(H.u8_of(out[0]), sf)
    assert (dq != deq(ref["q"].reshape(rows, -1), ref["sf"].reshape(rows, cols))).mean() <= tol
    if kind == "mx" and method == "quest":
        assert (H.u8_of(out[2]).reshape(-1).view(np.uint32) != ref["mask"]).mean() <= 1e-4
    np.testing.assert_array_equal(H.u8_of(Q.to_blocked(out[1])), H.blocked_sf(sf))


def test_quantize_tcgen05_kernel_is_the_large_input_default_and_matches_the_butterfly_kernel(b200q_env):
    """4096 x 4096, Hadamard-128 (bench.py's activation tensor) dispatches to the tcgen05 kernel; forcing the butterfly
    kernel gives the same bytes up to summation-order rounding (<= 1e-5 of the codes), any non-symmetric runtime rotation
    agrees with the oracle (the rotation is an MN-major tensor-core operand: a transposed R would show here)."""
    x = torch.randn(4096, 4096, dtype=torch.bfloat16, device="cuda", generator=torch.Generator("cuda").manual_seed(3)) * 25
    Rt = H.bf16_tensor_from_f32(O.hadamard_matrix(128))
    q1, sf1 = Q.fusedQuantizeMx(x, Rt, method="abs_max")
    b1 = Q.to_blocked(sf1).clone()
    b200q_env("B200Q_QUANT_TC", "0")
    q0, sf0 = Q.fusedQuantizeMx(x, Rt, method="abs_max")
    b0 = Q.to_blocked(sf0)
    torch.cuda.synchronize()
    assert (q0 != q1).float().mean().item() <= 1e-5
    assert (sf0.view(torch.uint8) != sf1.view(torch.uint8)).float().mean().item() <= 1e-5
    assert (b0.view(torch.uint8) != b1.view(torch.uint8)).float().mean().item() <= 1e-5
    b200q_env("B200Q_QUANT_TC", "1")
    for had in (32, 64, 128):
        xs = H.random_bf16((256, 1024), seed=had)
        R = O.bf16_round(np.random.default_rng(had).standard_normal((had, had)).astype(np.float32) * had ** -0.5)
        ref = O.quantize_mx(xs, R, "quest")
        q, sf = Q.fusedQuantizeMx(H.bf16_tensor_from_f32(xs), H.bf16_tensor_from_f32(R), method="quest")
        torch.cuda.synchronize()
        dq = O.dequant_mx(H.u8_of(q), _flat_sf(sf, 256, 32))
        assert (dq != O.dequant_mx(ref["q"].reshape(256, -1), ref["sf"].reshape(256, -1))).mean() <= 1e-3


def test_error_behaviour():
    R = torch.eye(32, dtype=torch.bfloat16, device="cuda")
    x = torch.zeros(4, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        Q.fusedQuantizeMx(x, R, method="nope")
    with pytest.raises(ValueError):
        Q.fusedQuantizeMx(x, R, method="abs_max", return_mask=True)
    with pytest.raises(RuntimeError, match="A must be bf16"):
        Q.fusedQuantizeMx(x.float(), R)
    with pytest.raises(RuntimeError, match="Unsupported rotation size"):
        Q.fusedQuantizeMx(x, torch.eye(8, dtype=torch.bfloat16, device="cuda"))
    with pytest.raises(RuntimeError, match="square"):
        Q.fusedQuantizeMx(x, torch.zeros(32, 16, dtype=torch.bfloat16, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA"):
        Q.fusedQuantizeMx(x.cpu(), R)
    a = torch.zeros(4, 32, dtype=torch.uint8, device="cuda")
    sf = torch.zeros(512, dtype=torch.float8_e8m0fnu, device="cuda")
    al = torch.ones(1, device="cuda")
    with pytest.raises(RuntimeError, match="A_sf must be float8_e8m0fnu"):
        Q.matmul_mxf4_bf16_tn(a, a, sf.view(torch.float8_e4m3fn), sf, al)
    with pytest.raises(RuntimeError, match="Inner dimensions"):
        Q.matmul_mxf4_bf16_tn(a, torch.zeros(4, 64, dtype=torch.uint8, device="cuda"), sf, sf, al)
    with pytest.raises(ValueError):
        Q.matmul_mxf4_bf16_tn(a, a, sf, sf, al, backend="bogus")
    with pytest.raises(ImportError):
        Q.matmul_mxf4_bf16_tn(a, a, sf, sf, al, backend="flashinfer")
    with pytest.raises(NotImplementedError):
        Q.matmul_ada_mxf4_bf16_tn(a, a, sf, sf, al)     # sm_120-only prototype: errors on sm_100 in the reference too
    with pytest.raises(AssertionError):
        Q.backward_t_bf16(a, a)                         # uint8 input: the reference asserts the dtypes (__init__.py:226)
    with pytest.raises(RuntimeError, match="A must be float8_e4m3fn"):
        Q.matmul_mxf8_bf16_nn(a, a, sf, sf, al)
    with pytest.raises(RuntimeError, match="A must be float8_e4m3fn"):
        Q.matmul_mxf8_bf16_tn(a, a, sf, sf, al)
    # C-ABI error convention: negative code + message
    lib = _lib.load()
    assert lib.b200q_gemm_fp4(None, None, None, None, None, None, 1, 1, 32, 0, None) < 0
    assert b"null" in lib.b200q_last_error()


# ----------------------------------------------------------------------------- GEMM
# (cta_group, block_n); (0, 0) = the planner's choice; cta_group 4 = CTA pairs in clusters of four with the A tiles
# multicast between the two pairs (opt-in experiment: bit-identical, measured slower -- profiles/r01_notes.md)
CFGS = [(0, 0), (1, 64), (1, 128), (1, 192), (1, 256), (2, 128), (2, 192), (2, 256), (4, 192), (4, 256)]


@pytest.mark.parametrize("kind", ["mx", "nv"])
@pytest.mark.parametrize("cfg", CFGS)
@pytest.mark.parametrize("shape", [(128, 128, 256), (256, 512, 1024), (504, 504, 2048), (1, 504, 4096),
                                   (16, 1000, 2176), (130, 72, 96), (5, 100, 64)])
def test_gemm_bit_exact_vs_oracle(kind, cfg, shape):
    m, n, k = shape
    aq, asf = H.random_fp4_operand(m, k, kind, seed=m + 1, sf_mode="narrow")
    bq, bsf = H.random_fp4_operand(n, k, kind, seed=n + 2, sf_mode="narrow")
    want = H.gemm_oracle_bits(aq, asf, bq, bsf, kind, 1.0)
    got = H.run_gemm(aq, asf, bq, bsf, kind, 1.0, cfg=cfg)
    mism, rel = H.compare_bits(got, want)
    if kind == "mx":
        assert mism == 0.0, (mism, rel)       # pow2 scales within +-1: fp32 accumulation is exact
    else:
        assert rel <= REL_TOL and mism <= 1e-3, (mism, rel)


@pytest.mark.parametrize("kind", ["mx", "nv"])
def test_gemm_wide_dynamic_range_within_tolerance(kind):
    m, n, k = 300, 1000, 2176
    aq, asf = H.random_fp4_operand(m, k, kind, seed=7, sf_mode="wide")
    bq, bsf = H.random_fp4_operand(n, k, kind, seed=8, sf_mode="wide")
    dq = O.dequant_mx if kind == "mx" else O.dequant_nv
    a_dq, b_dq = dq(aq, asf), dq(bq, bsf)
    want = a_dq @ b_dq.T
    got = O.bf16_from_bits(H.run_gemm(aq, asf, bq, bsf, kind, 1.0)).astype(np.float64)
    # fp32 accumulation: error bounded relative to the magnitude of the terms, not of a cancelling sum
    scale = np.abs(a_dq) @ np.abs(b_dq).T
    assert (np.abs(got - want) <= REL_TOL * np.abs(want) + 2.0 ** -20 * scale).all()


def test_gemm_config0_golden(golden):
    """BASELINE.json configs[0] (M=N=K=256 MXFP4 abs_max) -- the reference-generated golden output."""
    got = H.run_gemm(golden["c0_a_q"], golden["c0_a_s"], golden["c0_b_q"], golden["c0_b_s"], "mx", 1.0)
    np.testing.assert_array_equal(got, golden["c0_out_bits"])


def test_gemm_nv_golden(golden):
    got = H.run_gemm(golden["nvg_a_q"], golden["nvg_a_s"], golden["nvg_b_q"], golden["nvg_b_s"], "nv", 1.0)
    np.testing.assert_array_equal(got, golden["nvg_out_bits"])


@pytest.mark.parametrize("had,method,shape", [(32, "abs_max", (1, 504, 4096)), (64, "abs_max", (1, 504, 4096)),
                                              (128, "abs_max", (1, 504, 4096)), (32, "quest", (504, 504, 2048)),
                                              (128, "quest", (504, 504, 2048))])
def test_reference_style_quantize_then_gemm_mx(had, method, shape):
    """mirrors tests/mxfp4_test.py:208-269: randn*25 -> fusedQuantizeMx -> to_blocked -> matmul, bit-exact
    against the fp64 matmul of the dequantised operands."""
    m, n, k = shape
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(had))
    a = H.bf16_tensor_from_f32(H.random_bf16((m, k), seed=21))
    b = H.bf16_tensor_from_f32(H.random_bf16((n, k), seed=22))
    a_q, a_sf = Q.fusedQuantizeMx(a, R, method=method)
    b_q, b_sf = Q.fusedQuantizeMx(b, R, method=method)
    alpha = torch.tensor([1.0], device="cuda")
    out = Q.matmul_mxf4_bf16_tn(a_q, b_q, Q.to_blocked(a_sf, use_triton_kernel=True),
                                Q.to_blocked(b_sf, use_triton_kernel=True), alpha)
    torch.cuda.synchronize()
    a_dq = O.dequant_mx(H.u8_of(a_q), H.u8_of(a_sf)[:m, : k // 32])
    b_dq = O.dequant_mx(H.u8_of(b_q), H.u8_of(b_sf)[:n, : k // 32])
    np.testing.assert_array_equal(H.bf16_bits_of(out), O.gemm_ref(a_dq, b_dq, 1.0))


@pytest.mark.parametrize("had", [16, 128])
def test_reference_style_quantize_then_gemm_nv(had):
    """mirrors tests/nvfp4_test.py:190-224 (smaller n)."""
    m, n, k = 504, 1024, 4096
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(had))
    gs = torch.tensor([6.0], device="cuda")
    a = H.bf16_tensor_from_f32(H.random_bf16((m, k), seed=31))
    b = H.bf16_tensor_from_f32(H.random_bf16((n, k), seed=32))
    a_q, a_sf = Q.fusedQuantizeNv(a, R, gs)
    b_q, b_sf = Q.fusedQuantizeNv(b, R, gs)
    alpha = torch.tensor([1.0], device="cuda")
    out = Q.matmul_nvf4_bf16_tn(a_q, b_q, Q.to_blocked(a_sf, True).view(-1, k // 16),
                                Q.to_blocked(b_sf, True).view(-1, k // 16), alpha)
    torch.cuda.synchronize()
    a_dq = O.dequant_nv(H.u8_of(a_q), H.u8_of(a_sf)[:m, : k // 16])
    b_dq = O.dequant_nv(H.u8_of(b_q), H.u8_of(b_sf)[:n, : k // 16])
    mism, rel = H.compare_bits(H.bf16_bits_of(out), O.gemm_ref(a_dq, b_dq, 1.0))
    assert rel <= REL_TOL and mism <= 1e-3, (mism, rel)


def test_llama_shapes_small_batch():
    """subset of tests/mxfp4_test.py:272-299 (rand*25, quest, batch 1 and 16)."""
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(64))
    for (k, n) in ((4096, 4096), (4096, 14336), (8192, 1024)):
        for m in (1, 16):
            a = H.bf16_tensor_from_f32(H.random_bf16((m, k), seed=k + m, dist="rand"))
            b = H.bf16_tensor_from_f32(H.random_bf16((n, k), seed=n + 3, dist="rand"))
            a_q, a_sf = Q.fusedQuantizeMx(a, R, method="quest")
            b_q, b_sf = Q.fusedQuantizeMx(b, R, method="quest")
            out = Q.matmul_mxf4_bf16_tn(a_q, b_q, Q.to_blocked(a_sf), Q.to_blocked(b_sf),
                                        torch.tensor([1.0], device="cuda"))
            torch.cuda.synchronize()
            a_dq = O.dequant_mx(H.u8_of(a_q), H.u8_of(a_sf)[:m, : k // 32])
            b_dq = O.dequant_mx(H.u8_of(b_q), H.u8_of(b_sf)[:n, : k // 32])
            np.testing.assert_array_equal(H.bf16_bits_of(out), O.gemm_ref(a_dq, b_dq, 1.0))


@torch.inference_mode()
def test_under_inference_mode_like_the_reference_tests():
    """the reference's tests run under @torch.inference_mode (tests/mxfp4_test.py:209): tensors created there have
    no version counter -- the cached rotation hint and the attached blocked scales must still work."""
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(64))
    a = torch.randn(40, 512, dtype=torch.bfloat16, device="cuda") * 25
    b = torch.randn(72, 512, dtype=torch.bfloat16, device="cuda") * 25
    a_q, a_sf = Q.fusedQuantizeMx(a, R, method="abs_max")
    b_q, b_sf = Q.fusedQuantizeMx(b, R, method="abs_max")
    out = Q.matmul_mxf4_bf16_tn(a_q, b_q, Q.to_blocked(a_sf, use_triton_kernel=True), Q.to_blocked(b_sf, True),
                                torch.tensor([1.0], device="cuda"))
    torch.cuda.synchronize()
    a_dq = O.dequant_mx(H.u8_of(a_q), H.u8_of(a_sf)[:40, :16])
    b_dq = O.dequant_mx(H.u8_of(b_q), H.u8_of(b_sf)[:72, :16])
    np.testing.assert_array_equal(H.bf16_bits_of(out), O.gemm_ref(a_dq, b_dq, 1.0))


def test_torch_ops_schema_path():
    """external integrations call torch.ops._qutlass_C directly (bindings.cpp:498-507)."""
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(32))
    a = H.bf16_tensor_from_f32(H.random_bf16((128, 256), seed=41))
    out = torch.empty(128, 128, dtype=torch.uint8, device="cuda")
    sf = torch.empty(128, 8, dtype=torch.float8_e8m0fnu, device="cuda")
    r = torch.ops._qutlass_C.fusedQuantizeMxAbsMax(a, R, out, sf)
    q2, sf2 = Q.fusedQuantizeMx(a, R, method="abs_max")
    torch.cuda.synchronize()
    assert r[0].data_ptr() == out.data_ptr()
    np.testing.assert_array_equal(H.u8_of(out), H.u8_of(q2))
    np.testing.assert_array_equal(H.u8_of(sf), H.u8_of(sf2))
    blk = Q.to_blocked(sf, True)
    al = torch.ones(1, device="cuda")
    d1 = torch.ops._qutlass_C.matmul_mxf4_bf16_tn(out, out, blk, blk, al)
    d2 = Q.matmul_mxf4_bf16_tn(q2, q2, Q.to_blocked(sf2), Q.to_blocked(sf2), al)
    assert torch.equal(d1, d2)


def test_cuda_graph_capture():
    """the reference benchmarks time everything under CUDA-graph capture (bench_mxfp4_sm100.py:216-226)."""
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(128))
    a = H.bf16_tensor_from_f32(H.random_bf16((256, 1024), seed=51))
    b = H.bf16_tensor_from_f32(H.random_bf16((384, 1024), seed=52))
    b_q, b_sf = Q.fusedQuantizeMx(b, R, method="abs_max")
    b_blk = Q.to_blocked(b_sf)
    al = torch.ones(1, device="cuda")

    def run():
        a_q, a_sf = Q.fusedQuantizeMx(a, R, method="abs_max")
        return Q.matmul_mxf4_bf16_tn(a_q, b_q, Q.to_blocked(a_sf, True), b_blk, al)

    eager = run()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = run()
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)


def test_linear_host_entry_point():
    """b200q_linear_fp4_host (what bench.py's e2e leg calls) == the composed device path."""
    m, n, k, had = 200, 640, 1024, 64
    lib = _lib.load()
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(had))
    w = H.bf16_tensor_from_f32(H.random_bf16((n, k), seed=61))
    wq, wsf = Q.fusedQuantizeMx(w, R, method="abs_max")
    wblk = Q.to_blocked(wsf)
    x_host = torch.from_numpy(O.bf16_bits(H.random_bf16((m, k), seed=62)).astype(np.int16)).view(torch.bfloat16).pin_memory()
    d_host = torch.empty(m, n, dtype=torch.bfloat16).pin_memory()
    ws = torch.empty(lib.b200q_linear_workspace_bytes(m, n, k, 0), dtype=torch.uint8, device="cuda")
    al = torch.ones(1, device="cuda")
    _lib.check(lib.b200q_linear_fp4_host(x_host.data_ptr(), R.data_ptr(), wq.data_ptr(), wblk.data_ptr(), al.data_ptr(),
                                         None, d_host.data_ptr(), ws.data_ptr(), m, n, k, had, 0,
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    xq, xsf = Q.fusedQuantizeMx(x_host.cuda(), R, method="abs_max")
    want = Q.matmul_mxf4_bf16_tn(xq, wq, Q.to_blocked(xsf), wblk, al)
    torch.cuda.synchronize()
    assert torch.equal(d_host.cuda(), want)


@pytest.mark.parametrize("m,fmt", [(1100, "mx"), (640, "nv")])
def test_linear_host_entry_point_multi_slab(m, fmt):
    """several pipelined row slabs (128 + 384 + 512 + tail; b200q_linear_host_slabs) == one quantise + one GEMM over all rows."""
    n, k, had = 384, 512, 32
    lib = _lib.load()
    kind = 0 if fmt == "mx" else 1
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(had))
    w = H.bf16_tensor_from_f32(H.random_bf16((n, k), seed=63))
    gs = torch.tensor([2.0], device="cuda")
    wq, wsf = Q.fusedQuantizeMx(w, R, method="abs_max") if fmt == "mx" else Q.fusedQuantizeNv(w, R, gs, method="abs_max")
    wblk = Q.to_blocked(wsf)
    x_host = torch.from_numpy(O.bf16_bits(H.random_bf16((m, k), seed=64)).astype(np.int16)).view(torch.bfloat16).pin_memory()
    d_host = torch.zeros(m, n, dtype=torch.bfloat16).pin_memory()
    ws = torch.empty(lib.b200q_linear_workspace_bytes(m, n, k, kind), dtype=torch.uint8, device="cuda")
    al = torch.tensor([1.0 / 9.0], device="cuda")
    for _ in range(2):      # back to back: the second call reuses the helper streams / events
        _lib.check(lib.b200q_linear_fp4_host(x_host.data_ptr(), R.data_ptr(), wq.data_ptr(), wblk.data_ptr(), al.data_ptr(),
                                             gs.data_ptr(), d_host.data_ptr(), ws.data_ptr(), m, n, k, had, kind,
                                             torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    if fmt == "mx":
        xq, xsf = Q.fusedQuantizeMx(x_host.cuda(), R, method="abs_max")
        want = Q.matmul_mxf4_bf16_tn(xq, wq, Q.to_blocked(xsf), wblk, al)
    else:
        xq, xsf = Q.fusedQuantizeNv(x_host.cuda(), R, gs, method="abs_max")
        want = Q.matmul_nvf4_bf16_tn(xq, wq, Q.to_blocked(xsf), wblk, al)
    torch.cuda.synchronize()
    assert torch.equal(d_host.cuda(), want)


# ----------------------------------------------------------------------------- fused quantise + GEMM (one persistent kernel)
def _fused_case(m, n, k, had, method, fmt, seed):
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(had))
    x = H.bf16_tensor_from_f32(H.random_bf16((m, k), seed=seed))
    w = H.bf16_tensor_from_f32(H.random_bf16((n, k), seed=seed + 1))
    al = torch.tensor([1.0 / 9.0], device="cuda")
    gs = torch.tensor([3.0], device="cuda")
    if fmt == "mx":
        wq, wsf = Q.fusedQuantizeMx(w, R, method="abs_max")
        xq, xsf = Q.fusedQuantizeMx(x, R, method=method)
        want = Q.matmul_mxf4_bf16_tn(xq, wq, Q.to_blocked(xsf), Q.to_blocked(wsf), al)
    else:
        wq, wsf = Q.fusedQuantizeNv(w, R, gs, method="abs_max")
        xq, xsf = Q.fusedQuantizeNv(x, R, gs, method=method)
        want = Q.matmul_nvf4_bf16_tn(xq, wq, Q.to_blocked(xsf), Q.to_blocked(wsf), al)
    torch.cuda.synchronize()
    return R, x, wq, Q.to_blocked(wsf), al, gs, xq, xsf, want


@pytest.mark.parametrize("fmt,had,method", [("mx", 128, "abs_max"), ("mx", 64, "quest"), ("mx", 32, "abs_max"),
                                            ("nv", 16, "abs_max"), ("nv", 128, "quest"), ("nv", 64, "abs_max")])
@pytest.mark.parametrize("shape", [(512, 512, 1024), (1000, 1544, 2048), (300, 4096, 1024), (2048, 2304, 4096)])
def test_fused_linear_equals_two_calls(fmt, had, method, shape, b200q_env):
    """b200q_linear_fp4 (quantiser warps inside the persistent GEMM) must reproduce fusedQuantize* followed by matmul_*
    bit for bit: codes, both scale layouts and the bf16 output."""
    m, n, k = shape
    assert _lib.load().b200q_linear_fp4_launches(m, n, k, had, 1 | Q.ROT_TRUSTED_HADAMARD, 0) >= 2   # default: two launches
    b200q_env("B200Q_FUSE", "1")
    # the fused kernel's quantiser warps run the butterfly arithmetic; large standalone inputs would otherwise take the
    # tcgen05 kernel, which differs from it in fp32 summation order (<= 1e-5 of the codes, see the tcgen05 tests)
    b200q_env("B200Q_QUANT_TC", "0")
    R, x, wq, wblk, al, gs, xq, xsf, want = _fused_case(m, n, k, had, method, fmt, seed=m + n)
    lib = _lib.load()
    meth = (0 if method == "quest" else 1) | Q.ROT_TRUSTED_HADAMARD
    assert lib.b200q_linear_fp4_launches(m, n, k, had, meth, 0 if fmt == "mx" else 1) == 1      # really the fused kernel
    for _ in range(3):    # repeated calls reuse the self-cleaning counter workspace
        out, xq2, xsf2 = Q.fused_linear_fp4(x, R, wq, wblk, al, global_scale=gs if fmt == "nv" else None, method=method, fmt=fmt)
        torch.cuda.synchronize()
        assert torch.equal(xq2, xq)
        rows, cols = m, k // (32 if fmt == "mx" else 16)
        assert torch.equal(xsf2.view(torch.uint8)[:rows, :cols], xsf.view(torch.uint8)[:rows, :cols])
        # the blocked copy written by the fused kernel (handed over by to_blocked) == swizzle of the real scales, padding zero-filled
        np.testing.assert_array_equal(H.u8_of(Q.to_blocked(xsf2)), H.blocked_sf(H.u8_of(xsf).reshape(-1, xsf.shape[-1])[:rows, :cols]))
        assert torch.equal(out, want)
    for ws in Q._FUSE_WS.values():
        assert int(ws.view(torch.int32).abs().sum()) == 0           # left zeroed


def test_fused_linear_fallback_and_graph(b200q_env):
    """shapes the fused kernel does not take (K % 1024 != 0, small M) run as two launches with the same results; the
    fused kernel is CUDA-graph capturable once its workspace exists."""
    lib = _lib.load()
    b200q_env("B200Q_FUSE", "1")
    for (m, n, k) in ((512, 512, 1536), (128, 1024, 1024)):
        R, x, wq, wblk, al, gs, xq, xsf, want = _fused_case(m, n, k, 64, "abs_max", "mx", seed=7)
        assert lib.b200q_linear_fp4_launches(m, n, k, 64, 1 | Q.ROT_TRUSTED_HADAMARD, 0) == 2
        out, xq2, _ = Q.fused_linear_fp4(x, R, wq, wblk, al)
        torch.cuda.synchronize()
        assert torch.equal(out, want) and torch.equal(xq2, xq)
    m, n, k = 768, 1024, 2048
    R, x, wq, wblk, al, gs, xq, xsf, want = _fused_case(m, n, k, 128, "abs_max", "mx", seed=9)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        Q.fused_linear_fp4(x, R, wq, wblk, al)       # warm-up on the capture stream: creates its workspace
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out, _, _ = Q.fused_linear_fp4(x, R, wq, wblk, al)
    for _ in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want)


def test_fused_linear_full_size(b200q_env):
    """config 1 (4096 x 14336 x 4096) through the fused kernel, 20 back-to-back calls: identical to the two-kernel path."""
    m, n, k = 4096, 14336, 4096
    b200q_env("B200Q_FUSE", "1")
    b200q_env("B200Q_QUANT_TC", "0")     # bit-for-bit reference = the butterfly arithmetic the fused kernel runs
    R, x, wq, wblk, al, gs, xq, xsf, want = _fused_case(m, n, k, 128, "abs_max", "mx", seed=3)
    for i in range(20):
        out, xq2, _ = Q.fused_linear_fp4(x, R, wq, wblk, al)
    torch.cuda.synchronize()
    assert torch.equal(xq2, xq) and torch.equal(out, want)
    b200q_env("B200Q_FUSE", None)
    out, xq2, _ = Q.fused_linear_fp4(x, R, wq, wblk, al)          # default path: two launches behind the same call
    torch.cuda.synchronize()
    assert torch.equal(xq2, xq) and torch.equal(out, want)


# ----------------------------------------------------------------------------- MXFP8 ("next" row of the scope table)
@pytest.mark.parametrize("cfg", [(0, 0), (1, 128), (1, 256), (2, 128), (2, 192), (2, 256)])
@pytest.mark.parametrize("shape", [(128, 128, 128), (256, 512, 1024), (504, 504, 2048), (16, 1000, 2176), (1, 504, 4096),
                                   (130, 72, 96)])
def test_mxfp8_gemm_vs_oracle(cfg, shape):
    m, n, k = shape
    aq, asf = H.random_f8_operand(m, k, seed=m + 11)
    bq, bsf = H.random_f8_operand(n, k, seed=n + 12)
    want = H.gemm_oracle_bits(aq, asf, bq, bsf, "f8", 1.0)
    got = H.run_gemm(aq, asf, bq, bsf, "f8", 1.0, cfg=cfg)
    a_dq, b_dq = O.dequant_mxf8(aq, asf), O.dequant_mxf8(bq, bsf)
    scale = np.abs(a_dq) @ np.abs(b_dq).T
    g = O.bf16_from_bits(got).astype(np.float64)
    w = a_dq @ b_dq.T
    assert (np.abs(g - w) <= REL_TOL * np.abs(w) + 2.0 ** -20 * scale).all()
    assert (got != want).mean() <= 2e-2      # fp32 accumulation vs fp64: last-bit bf16 differences only


def test_mxfp8_golden_and_reference_style(golden):
    """golden vectors from the reference's _pseudoquant_mxfp8, then the reference's own test recipe
    (tests/mxfp8_test.py:60-78: rand*25 -> pseudoquant -> to_blocked(.., True) -> matmul_mxf8_bf16_tn,
    assert_close(atol=1e-1, rtol=1e-1)) at a Llama shape with batch 16."""
    got = H.run_gemm(golden["f8_a_q"], golden["f8_a_s"], golden["f8_b_q"], golden["f8_b_s"], "f8", 1.0)
    mism, rel = H.compare_bits(got, golden["f8_out64_bits"])
    assert rel <= REL_TOL and mism <= 1e-2
    m, n, k = 16, 4096, 4096
    a = H.random_bf16((m, k), seed=81, dist="rand")
    b = H.random_bf16((n, k), seed=82, dist="rand")
    a_q, a_s = O.pseudoquant_mxfp8(a)
    b_q, b_s = O.pseudoquant_mxfp8(b)
    a_sf = torch.from_numpy(a_s).cuda().view(torch.float8_e8m0fnu)       # un-padded [M, K/32], like the reference test
    b_sf = torch.from_numpy(b_s).cuda().view(torch.float8_e8m0fnu)
    out = Q.matmul_mxf8_bf16_tn(torch.from_numpy(a_q).cuda().view(torch.float8_e4m3fn),
                                torch.from_numpy(b_q).cuda().view(torch.float8_e4m3fn),
                                Q.to_blocked(a_sf, True), Q.to_blocked(b_sf, True), torch.tensor([1.0], device="cuda"))
    torch.cuda.synchronize()
    ref = torch.from_numpy((O.dequant_mxf8(a_q, a_s) @ O.dequant_mxf8(b_q, b_s).T)).to(torch.bfloat16)
    torch.testing.assert_close(out.cpu(), ref, atol=1e-1, rtol=1e-1)
    assert (H.bf16_bits_of(out) != O.gemm_ref(O.dequant_mxf8(a_q, a_s), O.dequant_mxf8(b_q, b_s))).mean() <= 1e-2


@pytest.mark.parametrize("cfg", [(0, 0), (1, 128), (1, 256), (2, 128), (2, 192), (2, 256)])
@pytest.mark.parametrize("shape", [(128, 128, 128), (256, 512, 1024), (496, 504, 2048), (16, 1000, 2176), (144, 72, 96),
                                   (1024, 1536, 512)])
def test_mxfp8_nn_gemm(cfg, shape):
    """matmul_mxf8_bf16_nn: A stored [K, M] (reference tests/mxfp8_test.py:77-96 builds it as a_e4m3.T.contiguous()).
    Same arithmetic as the tn kernel on the same logical operands -> bit-identical to it, and within fp32-accumulation
    distance of the oracle."""
    m, n, k = shape
    aq, asf = H.random_f8_operand(m, k, seed=m + 21)
    bq, bsf = H.random_f8_operand(n, k, seed=n + 22)
    tn = H.run_gemm(aq, asf, bq, bsf, "f8", 1.0, cfg=cfg)
    a_t = torch.from_numpy(np.ascontiguousarray(aq.T)).cuda().view(torch.float8_e4m3fn)          # [K, M]
    b = torch.from_numpy(bq).cuda().view(torch.float8_e4m3fn)
    al = torch.ones(1, device="cuda")
    out = Q._matmul_fp4("matmul_mxf8_bf16_nn", a_t, b, H.sf_torch(H.blocked_sf(asf), "mx"), H.sf_torch(H.blocked_sf(bsf), "mx"),
                        al, Q.KIND_MXF8_NN, torch.float8_e8m0fnu, 32, cfg=cfg)
    torch.cuda.synchronize()
    got = H.bf16_bits_of(out)
    assert np.array_equal(got, tn)
    want = H.gemm_oracle_bits(aq, asf, bq, bsf, "f8", 1.0)
    assert (got != want).mean() <= 2e-2


def test_mxfp8_nn_reference_style():
    """the reference's own nn recipe (tests/mxfp8_test.py:77-96): randn*25 -> pseudoquant -> to_blocked(.., True) ->
    a_e4m3.T.contiguous().view(k, m) -> matmul_mxf8_bf16_nn, assert_close(atol=1e-1, rtol=1e-1); Llama-7B shape, batch 16."""
    m, n, k = 16, 4096, 4096
    a = H.random_bf16((m, k), seed=91)
    b = H.random_bf16((n, k), seed=92)
    a_q, a_s = O.pseudoquant_mxfp8(a)
    b_q, b_s = O.pseudoquant_mxfp8(b)
    a_sf = torch.from_numpy(a_s).cuda().view(torch.float8_e8m0fnu)
    b_sf = torch.from_numpy(b_s).cuda().view(torch.float8_e8m0fnu)
    a_e4m3 = torch.from_numpy(a_q).cuda().view(torch.float8_e4m3fn)
    a_e4m3 = a_e4m3.T.contiguous().view((k, m))
    out = Q.matmul_mxf8_bf16_nn(a_e4m3, torch.from_numpy(b_q).cuda().view(torch.float8_e4m3fn),
                                Q.to_blocked(a_sf, True), Q.to_blocked(b_sf, True), torch.tensor([1.0], device="cuda"))
    torch.cuda.synchronize()
    ref = torch.from_numpy((O.dequant_mxf8(a_q, a_s) @ O.dequant_mxf8(b_q, b_s).T)).to(torch.bfloat16)
    torch.testing.assert_close(out.cpu(), ref, atol=1e-1, rtol=1e-1)
    out2 = torch.ops._qutlass_C.matmul_mxf8_bf16_nn(a_e4m3, torch.from_numpy(b_q).cuda().view(torch.float8_e4m3fn),
                                                    Q.to_blocked(a_sf, True), Q.to_blocked(b_sf, True), torch.tensor([1.0], device="cuda"))
    assert torch.equal(out, out2)


# ----------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("kind", ["mx", "nv"])
def test_full_size_properties(kind):
    """BASELINE.json configs[1]/[2]: M=4096, N=14336, K=4096 -- too big for a full CPU oracle, so:
    (a) a random sample of rows against the oracle, (b) tile-independence (row subset == same rows of the
    full product), (c) alpha linearity (x2 is exact in bf16), (d) every kernel configuration agrees bit for bit."""
    m, n, k = 4096, 14336, 4096
    aq, asf = H.random_fp4_operand(m, k, kind, seed=71, sf_mode="narrow")
    bq, bsf = H.random_fp4_operand(n, k, kind, seed=72, sf_mode="narrow")
    full = H.run_gemm(aq, asf, bq, bsf, kind, 1.0)
    rows = np.random.default_rng(0).choice(m, size=48, replace=False)
    want = H.gemm_oracle_bits(aq[rows], asf[rows], bq, bsf, kind, 1.0)
    mism, rel = H.compare_bits(full[rows], want)
    if kind == "mx":
        assert mism == 0.0
    else:
        assert rel <= REL_TOL and mism <= 1e-3
    sub = H.run_gemm(aq[256:384], asf[256:384], bq, bsf, kind, 1.0)
    np.testing.assert_array_equal(sub, full[256:384])
    twice = H.run_gemm(aq, asf, bq, bsf, kind, 2.0)
    np.testing.assert_array_equal(O.bf16_from_bits(twice), 2.0 * O.bf16_from_bits(full))
    for cfg in ((1, 128), (1, 192), (1, 256), (2, 128), (2, 192), (2, 256), (4, 192), (4, 256)):
        np.testing.assert_array_equal(H.run_gemm(aq, asf, bq, bsf, kind, 1.0, cfg=cfg), full)


def test_quantize_full_size_idempotent_layout():
    """config 3 size (M=16384, K=4096): blocked scales == swizzle(row-major scales); codes decode to finite
    values whose re-quantisation with the same scale is a fixed point (idempotence)."""
    m, k = 16384, 4096
    x = torch.randn(m, k, dtype=torch.bfloat16, device="cuda") * 25
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(128))
    q, sf = Q.fusedQuantizeMx(x, R, method="quest")
    blk = Q.to_blocked(sf)
    sw = Q.to_blocked(sf.clone(), True)
    torch.cuda.synchronize()
    assert torch.equal(blk.view(torch.uint8), sw.view(torch.uint8))
    codes = O.unpack_e2m1(H.u8_of(q[:64]))
    np.testing.assert_array_equal(O.e2m1_encode(O.e2m1_decode(codes)) & 7, codes & 7)


# ----------------------------------------------------------------------------- round 2: call orders / shapes not covered before
@pytest.mark.parametrize("fmt", ["mx", "nv"])
@pytest.mark.parametrize("shape", [(16, 384, 512), (160, 1024, 1024), (4096, 14336, 4096)])
def test_quantise_quantise_matmul_back_to_back_without_host_sync(fmt, shape):
    """the reference's own pattern `quantise(a); quantise(b); matmul(...)` with NOTHING between the
    kernels -- alpha and every other device tensor exist beforehand, so all launches are chained by programmatic dependent
    launch.  The GEMM must not read b_q / b_sf (written by the kernel right in front of it) before that kernel is done:
    the result must equal the fully synchronised sequence, on every iteration, with garbage-prefilled buffers."""
    m, n, k = shape
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(128 if k % 128 == 0 else 32))
    gs = torch.tensor([1.0], device="cuda")
    alpha = torch.tensor([1.0 / 7.0], device="cuda")
    g = torch.Generator("cuda").manual_seed(m + n)
    fq = (lambda t: Q.fusedQuantizeMx(t, R, method="abs_max")) if fmt == "mx" else (lambda t: Q.fusedQuantizeNv(t, R, gs, method="abs_max"))
    mm = Q.matmul_mxf4_bf16_tn if fmt == "mx" else Q.matmul_nvf4_bf16_tn
    for it in range(6):
        a = torch.randn(m, k, dtype=torch.bfloat16, device="cuda", generator=g) * 25
        b = torch.randn(n, k, dtype=torch.bfloat16, device="cuda", generator=g) * 25
        # reference result: every step synchronised
        aq, asf = fq(a); torch.cuda.synchronize()
        bq, bsf = fq(b); torch.cuda.synchronize()
        want = mm(aq, bq, Q.to_blocked(asf), Q.to_blocked(bsf), alpha); torch.cuda.synchronize()
        # poison the caching allocator's free blocks so a premature read sees garbage, not last iteration's identical bytes
        del aq, asf, bq, bsf
        junk = [torch.full((n, k // 2), 0x77, dtype=torch.uint8, device="cuda") for _ in range(3)]
        del junk
        torch.cuda.synchronize()
        aq, asf = fq(a)
        bq, bsf = fq(b)
        got = mm(aq, bq, Q.to_blocked(asf), Q.to_blocked(bsf), alpha)
        torch.cuda.synchronize()
        assert torch.equal(got, want), (fmt, shape, it, (got != want).float().mean().item())


def test_static_weights_flag_is_bit_identical_when_the_promise_holds():
    m, n, k = 300, 1000, 2048
    aq, asf = H.random_fp4_operand(m, k, "mx", seed=7, sf_mode="wide")
    bq, bsf = H.random_fp4_operand(n, k, "mx", seed=8, sf_mode="wide")
    a, b = torch.from_numpy(aq).cuda(), torch.from_numpy(bq).cuda()
    a_sf, b_sf = H.sf_torch(H.blocked_sf(asf), "mx"), H.sf_torch(H.blocked_sf(bsf), "mx")
    al = torch.tensor([0.5], device="cuda")
    torch.cuda.synchronize()
    d0 = Q.matmul_mxf4_bf16_tn(a, b, a_sf, b_sf, al)
    d1 = Q.matmul_mxf4_bf16_tn(a, b, a_sf, b_sf, al, static_weights=True)
    torch.cuda.synchronize()
    assert torch.equal(d0, d1)


@pytest.mark.parametrize("kind", ["mx", "nv"])
def test_config4_llama70b_ffn_shape(kind):
    """BASELINE.json configs[4]: N=28672, K=8192 (CTA-pair plan, 32 k-tiles), M=16384 -- a random sample of rows against
    the oracle, and every 2/4/8-way row shard (qutlass_b200.sharding.shard_rows) computed on its own equals the same rows
    of the full product bit for bit (what the multi-GPU run relies on)."""
    from qutlass_b200.sharding import shard_rows
    m, n, k = 16384, 28672, 8192
    aq, asf = H.random_fp4_operand(m, k, kind, seed=171, sf_mode="narrow")
    bq, bsf = H.random_fp4_operand(n, k, kind, seed=172, sf_mode="narrow")
    a, b = torch.from_numpy(aq).cuda(), torch.from_numpy(bq).cuda()
    b_sf = H.sf_torch(H.blocked_sf(bsf), kind)
    al = torch.tensor([1.0], device="cuda")
    knd, dt = (Q.KIND_MXF4, torch.float8_e8m0fnu) if kind == "mx" else (Q.KIND_NVF4, torch.float8_e4m3fn)
    full = Q._matmul_fp4("t", a, b, H.sf_torch(H.blocked_sf(asf), kind), b_sf, al, knd, dt, 16)
    rows = np.sort(np.random.default_rng(4).choice(m, size=24, replace=False))
    want = H.gemm_oracle_bits(aq[rows], asf[rows], bq, bsf, kind, 1.0)
    got = H.bf16_bits_of(full[torch.from_numpy(rows).cuda()])
    mism, rel = H.compare_bits(got, want)
    if kind == "mx":
        assert mism == 0.0, (mism, rel)
    else:
        assert rel <= REL_TOL and mism <= 1e-3, (mism, rel)
    for world in (2, 8):
        for rank in (0, world - 1):
            s0, r = shard_rows(m, world, rank)
            part = Q._matmul_fp4("t", a[s0:s0 + r], b, H.sf_torch(H.blocked_sf(asf[s0:s0 + r]), kind), b_sf, al, knd, dt, 16)
            assert torch.equal(part, full[s0:s0 + r]), (world, rank)


def test_gemm_output_larger_than_2_31_elements():
    """The reference's own benchmark sweeps M up to 65536 at N = 57344 (benchmarks/bench_mxfp4_sm100.py:176-193,264): 3.76e9
    outputs, beyond 32-bit element indices.  K is kept small so the check stays cheap: the last rows against the oracle, and
    a row block computed on its own equals the same rows of the big product."""
    m, n, k = 40960, 57344, 256
    assert m * n > 2 ** 31
    aq, asf = H.random_fp4_operand(m, k, "mx", seed=271, sf_mode="narrow")
    bq, bsf = H.random_fp4_operand(n, k, "mx", seed=272, sf_mode="narrow")
    a, b = torch.from_numpy(aq).cuda(), torch.from_numpy(bq).cuda()
    b_sf = H.sf_torch(H.blocked_sf(bsf), "mx")
    al = torch.tensor([1.0], device="cuda")
    full = Q._matmul_fp4("t", a, b, H.sf_torch(H.blocked_sf(asf), "mx"), b_sf, al, Q.KIND_MXF4, torch.float8_e8m0fnu, 16)
    rows = np.array([0, 1, 20479, 37448, 40958, 40959])
    want = H.gemm_oracle_bits(aq[rows], asf[rows], bq, bsf, "mx", 1.0)
    np.testing.assert_array_equal(H.bf16_bits_of(full[torch.from_numpy(rows).cuda()]), want)
    s0 = 40960 - 512
    part = Q._matmul_fp4("t", a[s0:], b, H.sf_torch(H.blocked_sf(asf[s0:]), "mx"), b_sf, al, Q.KIND_MXF4, torch.float8_e8m0fnu, 16)
    assert torch.equal(part, full[s0:])


def test_tensor_map_cache_hits_on_repeated_calls_and_never_serves_a_stale_map():
    """Host overhead: the second call with the same buffers re-uses all five tensor maps; a different
    buffer at a recycled address with a different shape gets its own encoding (the key is the full descriptor content)."""
    import ctypes
    lib = _lib.load()
    h, mi = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    aq, asf = H.random_fp4_operand(256, 512, "mx", seed=1)
    bq, bsf = H.random_fp4_operand(256, 512, "mx", seed=2)
    want = H.gemm_oracle_bits(aq, asf, bq, bsf, "mx", 1.0)
    a, b = torch.from_numpy(aq).cuda(), torch.from_numpy(bq).cuda()
    a_sf, b_sf = H.sf_torch(H.blocked_sf(asf), "mx"), H.sf_torch(H.blocked_sf(bsf), "mx")
    al = torch.tensor([1.0], device="cuda")
    out = torch.empty(256, 256, dtype=torch.bfloat16, device="cuda")
    call = lambda: _lib.check(lib.b200q_gemm_fp4(a.data_ptr(), b.data_ptr(), a_sf.data_ptr(), b_sf.data_ptr(), al.data_ptr(),
                                                 out.data_ptr(), 256, 256, 512, 0, torch.cuda.current_stream().cuda_stream))
    call()
    lib.b200q_debug_tmap_cache_stats(ctypes.byref(h), ctypes.byref(mi))
    call()
    lib.b200q_debug_tmap_cache_stats(ctypes.byref(h), ctypes.byref(mi))
    assert h.value == 5 and mi.value == 0, (h.value, mi.value)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(H.bf16_bits_of(out), want)
    # same addresses, different logical shape (K halves): must not reuse the old maps
    want2 = H.gemm_oracle_bits(aq[:, :128], asf[:, :8], bq[:, :128], bsf[:, :8], "mx", 1.0)
    a2, b2 = a[:, :128].contiguous(), b[:, :128].contiguous()
    a.copy_(torch.zeros_like(a)); b.copy_(torch.zeros_like(b))
    a.view(-1)[: a2.numel()].copy_(a2.view(-1)); b.view(-1)[: b2.numel()].copy_(b2.view(-1))
    a_sf2, b_sf2 = H.sf_torch(H.blocked_sf(asf[:, :8]), "mx"), H.sf_torch(H.blocked_sf(bsf[:, :8]), "mx")
    _lib.check(lib.b200q_gemm_fp4(a.data_ptr(), b.data_ptr(), a_sf2.data_ptr(), b_sf2.data_ptr(), al.data_ptr(),
                                  out.data_ptr(), 256, 256, 256, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(H.bf16_bits_of(out), want2)


def test_to_blocked_hand_over_and_invalidation():
    """the blocked copy written by the quantiser is handed over ONCE, only while the row-major tensor is
    unmodified; an in-place edit, a second call, a raw-op overwrite or an inference-mode tensor all take the swizzle kernel
    and see the CURRENT bytes."""
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(32))
    x = torch.randn(200, 256, dtype=torch.bfloat16, device="cuda") * 25
    q, sf = Q.fusedQuantizeMx(x, R, method="abs_max")
    first = Q.to_blocked(sf)
    second = Q.to_blocked(sf)
    torch.cuda.synchronize()
    assert first.data_ptr() != second.data_ptr()
    # same REAL scales (the pad rows of the row-major tensor are uninitialised memory, like the reference's torch.empty;
    # only the copy written by the quantiser has them zero-filled)
    unb = lambda t: O.from_blocked(H.u8_of(t), 256, 8)[:200]
    np.testing.assert_array_equal(unb(first), unb(second))
    np.testing.assert_array_equal(H.u8_of(first), H.blocked_sf(H.u8_of(sf)[:200]))
    q, sf = Q.fusedQuantizeMx(x, R, method="abs_max")
    sf.view(torch.uint8)[200:256] = 127                      # the reference tests' `scales[m:m_up] = 1.0` pattern
    blk = Q.to_blocked(sf)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(H.u8_of(blk), H.blocked_sf(H.u8_of(sf)))      # incl. the edited pad rows: the CURRENT bytes
    # raw op writes OUT_sf through its data pointer: an attached copy from an earlier quantisation must not survive
    q, sf = Q.fusedQuantizeMx(x, R, method="abs_max")
    torch.ops._qutlass_C.fusedQuantizeMxAbsMax(x * 4, R, q, sf)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(H.u8_of(Q.to_blocked(sf)), H.blocked_sf(H.u8_of(sf)))
    with torch.inference_mode():
        xi = torch.randn(200, 256, dtype=torch.bfloat16, device="cuda") * 25
        q, sf = Q.fusedQuantizeMx(xi, R.clone(), method="abs_max")
        sf.view(torch.uint8)[200:256] = 127
        blk = Q.to_blocked(sf)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(H.u8_of(blk), H.blocked_sf(H.u8_of(sf)))


# ----------------------------------------------------------------------------- decode kernel (gemm_decode.cu)
@pytest.mark.parametrize("kind", ["mx", "nv"])
@pytest.mark.parametrize("shape", [(1, 504, 4096), (16, 14336, 4096), (7, 1000, 2048), (32, 2816, 1024), (17, 640, 256),
                                   (16, 128 * 150, 512), (3, 100, 768), (64, 1024, 2048), (48, 640, 1024)])
def test_decode_kernel_bit_identical_to_the_general_kernel_and_exact_vs_oracle(kind, shape):
    """The swapped-operand weight-streaming kernel (configuration (1, 16), the default for M <= 32) computes the same
    products in the same k order as gemm_fp4_kernel: bit-identical to it on wide-dynamic-range operands, bit-exact against
    the fp64 oracle where fp32 accumulation is exact.  Shapes: N % 128 != 0 (ragged last tile), N % 8 != 0, M = 1 / 17 / 32
    (16- and 32-row MMA shapes), more tiles than SMs (two accumulators alternate), a single k-tile."""
    m, n, k = shape
    for mode in ("narrow", "wide"):
        aq, asf = H.random_fp4_operand(m, k, kind, seed=m + 31, sf_mode=mode)
        bq, bsf = H.random_fp4_operand(n, k, kind, seed=n + 32, sf_mode=mode)
        got = H.run_gemm(aq, asf, bq, bsf, kind, 1.0 / 3.0, cfg=(1, 16))
        np.testing.assert_array_equal(got, H.run_gemm(aq, asf, bq, bsf, kind, 1.0 / 3.0, cfg=(1, 128)))
        np.testing.assert_array_equal(got, H.run_gemm(aq, asf, bq, bsf, kind, 1.0 / 3.0))          # ... and to the planner's choice
        if mode == "narrow":
            want = H.gemm_oracle_bits(aq, asf, bq, bsf, kind, 1.0)
            mism, rel = H.compare_bits(H.run_gemm(aq, asf, bq, bsf, kind, 1.0, cfg=(1, 16)), want)
            if kind == "mx":
                assert mism == 0.0, (mism, rel)
            else:
                assert rel <= REL_TOL and mism <= 1e-3, (mism, rel)


@pytest.mark.parametrize("kind", ["mx", "nv"])
@pytest.mark.parametrize("shape", [(16, 14336, 4096), (1, 28672, 4096), (32, 2816, 1024), (7, 1000, 2048)])
@pytest.mark.parametrize("pace", [0, 40, 600])
def test_decode_kernel_paced_weight_stream_is_bit_identical(kind, shape, pace, b200q_env):
    """B200Q_DECODE_PACE: the weight stages of a decode-kernel CTA are requested on a clock (and prefetched into L2 by an idle
    warp) instead of all at once -- the same loads into the same ring, so the same bytes.  The library's own rule (the
    reference result here) paces default launches; 0 = everything at once, 40 = gates that have always expired already, 600 =
    slower than the stream.  (1, 28672, 4096): 224 tiles, two per CTA for half of the grid (the clock runs across tiles)."""
    m, n, k = shape
    aq, asf = H.random_fp4_operand(m, k, kind, seed=m + 31, sf_mode="wide")
    bq, bsf = H.random_fp4_operand(n, k, kind, seed=n + 32, sf_mode="wide")
    want = H.run_gemm(aq, asf, bq, bsf, kind, 1.0 / 3.0, cfg=(1, 16))
    b200q_env("B200Q_DECODE_PACE", str(pace))
    for _ in range(2):
        np.testing.assert_array_equal(H.run_gemm(aq, asf, bq, bsf, kind, 1.0 / 3.0, cfg=(1, 16)), want)


def test_decode_kernel_is_rejected_where_it_does_not_apply():
    aq, asf = H.random_fp4_operand(65, 512, "mx", seed=1)
    bq, bsf = H.random_fp4_operand(256, 512, "mx", seed=2)
    with pytest.raises(Exception, match="decode kernel"):
        H.run_gemm(aq, asf, bq, bsf, "mx", 1.0, cfg=(1, 16))


@pytest.mark.parametrize("fmt", ["mx", "nv"])
def test_decode_path_in_a_cuda_graph_with_quantiser_in_front(fmt):
    """quantise -> (to_blocked no-op) -> decode GEMM captured in one CUDA graph (the reference benchmark's loop at batch 1),
    replayed with changing activations: equals the eager, synchronised result every time."""
    m, n, k = 4, 1536, 2048
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(128))
    gs = torch.tensor([1.0], device="cuda")
    al = torch.tensor([0.25], device="cuda")
    fq = (lambda t: Q.fusedQuantizeMx(t, R, method="abs_max")) if fmt == "mx" else (lambda t: Q.fusedQuantizeNv(t, R, gs, method="abs_max"))
    mm = Q.matmul_mxf4_bf16_tn if fmt == "mx" else Q.matmul_nvf4_bf16_tn
    w = torch.randn(n, k, dtype=torch.bfloat16, device="cuda") * 25
    wq, wsf = fq(w)
    wblk = Q.to_blocked(wsf)
    x = torch.randn(m, k, dtype=torch.bfloat16, device="cuda") * 25
    fq(x)      # the rotation matrix is classified outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            xq, xsf = fq(x)
            out = mm(xq, wq, Q.to_blocked(xsf), wblk, al, static_weights=True)
    for it in range(4):
        x.copy_(torch.randn(m, k, dtype=torch.bfloat16, device="cuda") * 25)
        g.replay()
        torch.cuda.synchronize()
        xq2, xsf2 = fq(x)
        torch.cuda.synchronize()
        want = mm(xq2, wq, Q.to_blocked(xsf2), wblk, al)
        torch.cuda.synchronize()
        assert torch.equal(out, want), it


@pytest.mark.parametrize("fmt,had,method", [("mx", 128, "abs_max"), ("mx", 64, "quest"), ("mx", 32, "abs_max"),
                                            ("nv", 16, "abs_max"), ("nv", 128, "quest"), ("nv", 128, "abs_max"), ("nv", 64, "abs_max")])
@pytest.mark.parametrize("shape", [(16, 1536, 2048), (1, 512, 1024), (32, 640, 4096), (7, 1000, 1024), (16, 128 * 150, 1024)])
def test_decode_step_in_one_launch_equals_two_calls(fmt, had, method, shape, b200q_env):
    """b200q_linear_fp4 for M <= 32 is ONE launch -- every CTA of the weight-streaming kernel rotates
    and quantises the activations itself -- and reproduces fusedQuantize* followed by matmul_* bit for bit: the bf16 output,
    the codes, the row-major scales and the blocked copy (written by CTA 0)."""
    m, n, k = shape
    R, x, wq, wblk, al, gs, xq, xsf, want = _fused_case(m, n, k, had, method, fmt, seed=m + n + had)
    lib = _lib.load()
    meth = (0 if method == "quest" else 1) | Q.ROT_TRUSTED_HADAMARD
    assert lib.b200q_linear_fp4_launches(m, n, k, had, meth, 0 if fmt == "mx" else 1) == 2     # default: two launches (faster)
    b200q_env("B200Q_FUSE_DECODE", "1")
    assert lib.b200q_linear_fp4_launches(m, n, k, had, meth, 0 if fmt == "mx" else 1) == 1
    for _ in range(2):
        out, xq2, xsf2 = Q.fused_linear_fp4(x, R, wq, wblk, al, global_scale=gs if fmt == "nv" else None, method=method, fmt=fmt)
        torch.cuda.synchronize()
        assert torch.equal(out, want), (out != want).float().mean().item()
        assert torch.equal(xq2, xq)
        rows, cols = m, k // (32 if fmt == "mx" else 16)
        assert torch.equal(xsf2.view(torch.uint8)[:rows, :cols], xsf.view(torch.uint8)[:rows, :cols])
        np.testing.assert_array_equal(H.u8_of(Q.to_blocked(xsf2)), H.blocked_sf(H.u8_of(xsf).reshape(-1, xsf.shape[-1])[:rows, :cols]))


def test_decode_step_one_launch_in_a_cuda_graph_and_switch(b200q_env):
    m, n, k = 8, 2048, 4096
    R, x, wq, wblk, al, gs, xq, xsf, want = _fused_case(m, n, k, 128, "abs_max", "mx", seed=77)
    b200q_env("B200Q_FUSE_DECODE", "1")
    assert _lib.load().b200q_linear_fp4_launches(m, n, k, 128, 1 | Q.ROT_TRUSTED_HADAMARD, 0) == 1
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            out, _, _ = Q.fused_linear_fp4(x, R, wq, wblk, al)
    for _ in range(3):
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want)
    b200q_env("B200Q_FUSE_DECODE", None)
    assert _lib.load().b200q_linear_fp4_launches(m, n, k, 128, 1 | Q.ROT_TRUSTED_HADAMARD, 0) == 2
    out2, _, _ = Q.fused_linear_fp4(x, R, wq, wblk, al)
    torch.cuda.synchronize()
    assert torch.equal(out2, want)


def test_rotation_hint_stops_synchronising_for_throw_away_rotations():
    """a caller that builds a new rotation tensor for every call must not pay a device sync per call for
    ever: after 64 inspected-and-collected tensors new ones get no hint (the kernel's own device-side check runs) -- and the
    results stay identical."""
    x = torch.randn(64, 1024, dtype=torch.bfloat16, device="cuda") * 25
    base = H.bf16_tensor_from_f32(O.hadamard_matrix(64))
    want = Q.fusedQuantizeMx(x, base, method="abs_max")
    for i in range(80):
        Rn = base.clone()
        got = Q.fusedQuantizeMx(x, Rn, method="abs_max")
        del Rn
    torch.cuda.synchronize()
    assert Q._ROT_DEAD[0] >= Q._ROT_MAX_DEAD
    assert Q._rotation_hint(base.clone()) == 0
    assert torch.equal(got[0], want[0]) and torch.equal(got[1].view(torch.uint8)[:64], want[1].view(torch.uint8)[:64])
