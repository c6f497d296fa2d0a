"""Pin the CPU oracle against vectors produced by the reference's own test helpers
(tests/golden/make_golden.py ran /root/reference/tests/{mxfp4,nvfp4}_test.py code)."""
import numpy as np
import pytest

import oracle as O


def test_e2m1_rounding_matches_reference_rtne(golden):
    x = golden["rtne_in"]
    code = O.e2m1_encode(x)
    # reference maps +0.0 to code 8 (bucketize quirk): compare decoded values + packed bytes mod sign-of-zero
    np.testing.assert_array_equal(O.e2m1_decode(code), golden["rtne_val"])
    ref_codes = O.unpack_e2m1(golden["rtne_packed"])
    nz = (code & 7) != 0
    np.testing.assert_array_equal(code[nz], ref_codes[nz])
    assert ((ref_codes[~nz] & 7) == 0).all()


def test_pack_low_nibble_first():
    c = np.array([1, 2, 3, 15], dtype=np.uint8)
    np.testing.assert_array_equal(O.pack_e2m1(c), np.array([0x21, 0xF3], dtype=np.uint8))
    np.testing.assert_array_equal(O.unpack_e2m1(O.pack_e2m1(c)), c)


@pytest.mark.parametrize("h", [32, 64, 128])
def test_hadamard_matches_reference(golden, h):
    np.testing.assert_array_equal(O.bf16_bits(O.hadamard_matrix(h)), golden[f"had{h}_bits"])
    assert O.is_sylvester_hadamard(O.hadamard_matrix(h))
    assert not O.is_sylvester_hadamard(np.eye(h, dtype=np.float32))


@pytest.mark.parametrize("h", [32, 64, 128])
@pytest.mark.parametrize("method", ["quest", "abs_max"])
def test_mx_ref64_bit_exact(golden, h, method):
    x = O.bf16_from_bits(golden["mx_x_bits"])
    R = O.bf16_from_bits(golden[f"had{h}_bits"])
    tag = f"mx_h{h}_{'quest' if method == 'quest' else 'absmax'}"
    r = O.quantize_mx(x, R, method, arithmetic="ref64")
    np.testing.assert_array_equal(r["sf"].reshape(golden[tag + "_e8m0"].shape), golden[tag + "_e8m0"])
    dq = O.dequant_mx(r["q"].reshape(golden[tag + "_e2m1"].shape),
                      r["sf"].reshape(golden[tag + "_e8m0"].shape),
                      alpha=1.0 if method == "quest" else 3.0)
    np.testing.assert_array_equal(dq, golden[tag + "_dq"])
    # clip mask: reference packs bit i of byte j = element 8j+i  == our uint32 little-endian
    mask_bytes = r["mask"].view(np.uint8).reshape(golden[tag + "_mask"].shape)
    np.testing.assert_array_equal(mask_bytes, golden[tag + "_mask"])


@pytest.mark.parametrize("h", [32, 64, 128])
@pytest.mark.parametrize("method", ["quest", "abs_max"])
def test_mx_kernel_arithmetic_within_reference_tolerance(golden, h, method):
    """fp32 'kernel' flavour vs the reference's fp64 test oracle: the reference's own
    acceptance bar is mismatch fraction <= 1e-4 (tests/mxfp4_test.py:220-221)."""
    x = O.bf16_from_bits(golden["mx_x_bits"])
    R = O.bf16_from_bits(golden[f"had{h}_bits"])
    tag = f"mx_h{h}_{'quest' if method == 'quest' else 'absmax'}"
    r = O.quantize_mx(x, R, method, arithmetic="kernel")
    dq = O.dequant_mx(r["q"].reshape(golden[tag + "_e2m1"].shape),
                      r["sf"].reshape(golden[tag + "_e8m0"].shape),
                      alpha=1.0 if method == "quest" else 3.0)
    assert (dq != golden[tag + "_dq"]).mean() <= 1e-4


@pytest.mark.parametrize("h", [16, 32, 64, 128])
def test_nv_ref64_bit_exact(golden, h):
    x = O.bf16_from_bits(golden["mx_x_bits"])
    R = O.bf16_from_bits(golden[f"had{h}_bits"])
    r = O.quantize_nv(x, R, 6.0, "abs_max", arithmetic="ref64")
    tag = f"nv_h{h}"
    np.testing.assert_array_equal(r["sf"].reshape(golden[tag + "_e4m3"].shape), golden[tag + "_e4m3"])
    dq = O.dequant_nv(r["q"].reshape(golden[tag + "_e2m1"].shape),
                      r["sf"].reshape(golden[tag + "_e4m3"].shape), alpha=6.0)
    np.testing.assert_array_equal(dq, golden[tag + "_dq"])


@pytest.mark.parametrize("h", [16, 32, 64, 128])
def test_nv_kernel_arithmetic_within_reference_tolerance(golden, h):
    """reference bar for NV: mismatch fraction <= 1e-1 (tests/nvfp4_test.py:204-205);
    we hold the oracle's kernel flavour to 1e-2."""
    x = O.bf16_from_bits(golden["mx_x_bits"])
    R = O.bf16_from_bits(golden[f"had{h}_bits"])
    tag = f"nv_h{h}"
    # the arithmetic of the reference's mma.sync kernels (for h = 128: what it runs on sm_120) is within 1e-2 of its test oracle
    r = O.quantize_nv(x, R, 6.0, "abs_max", arithmetic="kernel", sm100_codes=False)
    dq = O.dequant_nv(r["q"].reshape(golden[tag + "_e2m1"].shape),
                      r["sf"].reshape(golden[tag + "_e4m3"].shape), alpha=6.0)
    assert (dq != golden[tag + "_dq"]).mean() <= 1e-2
    # the default flavour = what the reference runs ON sm_100: for h = 128 its sm_100-only kernel (codes from the unrounded
    # scale) sits 4.8 % away from that oracle -- inside the reference's own 1e-1 bar, and the reason the bar is that loose
    r = O.quantize_nv(x, R, 6.0, "abs_max", arithmetic="kernel")
    dq = O.dequant_nv(r["q"].reshape(golden[tag + "_e2m1"].shape),
                      r["sf"].reshape(golden[tag + "_e4m3"].shape), alpha=6.0)
    assert (dq != golden[tag + "_dq"]).mean() <= (1e-1 if h == 128 else 1e-2)


@pytest.mark.parametrize("shape", [(128, 4), (256, 8), (384, 12), (128, 128)])
def test_to_blocked_matches_reference(golden, shape):
    r, c = shape
    sf = golden[f"blk_{r}x{c}_in"]
    blk = O.to_blocked(sf)
    np.testing.assert_array_equal(blk, golden[f"blk_{r}x{c}_out"])
    rr, cc = np.meshgrid(np.arange(r), np.arange(c), indexing="ij")
    np.testing.assert_array_equal(blk[O.swizzle_offset(rr, cc, c)], sf)
    np.testing.assert_array_equal(O.from_blocked(blk, r, c), sf)


def test_padded_shapes(golden):
    assert tuple(golden["padded_mx_3x200x4096"]) == O.padded_sf_shape(600, 4096 // 32)
    assert tuple(golden["padded_nv_3x200x4096"]) == O.padded_sf_shape(600, 4096 // 16)
    assert tuple(golden["padded_mx_1x96"]) == O.padded_sf_shape(1, 3)
    assert tuple(golden["padded_nv_1x96"]) == O.padded_sf_shape(1, 6)


def test_config0_gemm_256_cpu(golden):
    """BASELINE.json configs[0]: M=N=K=256 MXFP4 abs_max, emulated quantise + matmul on CPU."""
    a = O.bf16_from_bits(golden["c0_a_bits"])
    b = O.bf16_from_bits(golden["c0_b_bits"])
    R = O.hadamard_matrix(32)
    qa = O.quantize_mx(a, R, "abs_max", arithmetic="ref64")
    qb = O.quantize_mx(b, R, "abs_max", arithmetic="ref64")
    np.testing.assert_array_equal(qa["sf"].reshape(256, 8), golden["c0_a_s"])
    np.testing.assert_array_equal(qb["sf"].reshape(256, 8), golden["c0_b_s"])
    a_dq = O.dequant_mx(golden["c0_a_q"], golden["c0_a_s"])
    b_dq = O.dequant_mx(golden["c0_b_q"], golden["c0_b_s"])
    np.testing.assert_array_equal(O.dequant_mx(qa["q"].reshape(256, 128), qa["sf"].reshape(256, 8)), a_dq)
    out = O.gemm_ref(a_dq, b_dq, 1.0)
    np.testing.assert_array_equal(out, golden["c0_out_bits"])
    # fp32 accumulation gives the same bf16 here
    np.testing.assert_array_equal(golden["c0_out32_bits"], golden["c0_out_bits"])


def test_nv_gemm_golden(golden):
    a_dq = O.dequant_nv(golden["nvg_a_q"], golden["nvg_a_s"])
    b_dq = O.dequant_nv(golden["nvg_b_q"], golden["nvg_b_s"])
    np.testing.assert_array_equal(O.gemm_ref(a_dq, b_dq, 1.0), golden["nvg_out_bits"])


def test_e4m3_roundtrip_and_saturation():
    b = np.arange(256, dtype=np.uint8)
    v = O.e4m3_decode(b)
    ok = ~np.isnan(v)
    enc = O.e4m3_encode(v[ok].astype(np.float32))
    # -0 / +0 both valid encodings of zero
    np.testing.assert_array_equal(O.e4m3_decode(enc), v[ok])
    assert O.e4m3_encode(np.float32(1e6)) == 0x7E  # 448
    assert O.e4m3_encode(np.float32(0.0)) == 0


def test_e8m0_decode():
    assert O.e8m0_decode(np.uint8(127)) == 1.0
    assert O.e8m0_decode(np.uint8(0)) == 2.0 ** -127
    assert np.isnan(O.e8m0_decode(np.uint8(255)))


def test_mxfp8_pseudoquant_and_gemm_golden(golden):
    """'next' row (matmul_mxf8_bf16_tn): the oracle's MXFP8 pseudo-quantiser and GEMM criterion against vectors from
    the reference's _pseudoquant_mxfp8 (tests/mxfp8_test.py:27-78)."""
    for t in ("a", "b"):
        x = O.bf16_from_bits(golden[f"f8_{t}_bits"])
        q, s = O.pseudoquant_mxfp8(x)
        np.testing.assert_array_equal(q, golden[f"f8_{t}_q"])
        np.testing.assert_array_equal(s, golden[f"f8_{t}_s"])
    out = O.gemm_ref(O.dequant_mxf8(golden["f8_a_q"], golden["f8_a_s"]), O.dequant_mxf8(golden["f8_b_q"], golden["f8_b_s"]))
    np.testing.assert_array_equal(out, golden["f8_out64_bits"])


def test_sm100_nv_quirk_flavour_reproduces_the_observed_divergence():
    """The reference's sm_100-only NVFP4 abs_max Hadamard-128 kernel computes the codes with the scale BEFORE its e4m3
    rounding (sm100_visitor_store_tma_warpspecialized.hpp:141-148,567-591).  On a B200 that kernel differed from ours (= the
    reference's other kernels / its test oracle) in 4.76 % of the dequantised values and 9.2 % of the code bytes with identical
    scale bytes (profiles/r01_ref_quant_diag.jsonl).  The oracle's `sm100_codes` flavour restates the quirk: on data of the
    same distribution it must move the same fraction of values -- the diagnosis, checked on CPU."""
    import helpers as H
    rows, k = 64, 4096
    x = H.random_bf16((rows, k), seed=1134)
    R = O.hadamard_matrix(128)
    a = O.quantize_nv(x, R, 6.0, "abs_max", sm100_codes=False)
    b = O.quantize_nv(x, R, 6.0, "abs_max")                      # default = the reference's sm_100 dispatch for H = 128
    assert np.array_equal(b["q"], O.quantize_nv(x, R, 6.0, "abs_max", sm100_codes=True)["q"])
    np.testing.assert_array_equal(a["sf"], b["sf"])
    cols = k // 16
    da = O.dequant_nv(a["q"].reshape(rows, -1), a["sf"].reshape(rows, cols))
    db = O.dequant_nv(b["q"].reshape(rows, -1), b["sf"].reshape(rows, cols))
    assert 0.04 <= float((da != db).mean()) <= 0.055
    assert 0.08 <= float((a["q"] != b["q"]).mean()) <= 0.105
