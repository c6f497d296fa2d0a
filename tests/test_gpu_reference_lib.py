"""GPU parity against the REAL reference kernels (run with -m gpu on a B200).

oracle/build_ref.py compiles the unmodified reference sources (/root/reference/qutlass/csrc/*.cu + its vendored
CUTLASS) for sm_100a in the build container; the resulting oracle/_ref/qutlass_ref_C.so travels to the GPU box.  It
registers the reference's ops in torch's `_qutlass_C` library -- the namespace our drop-in registers too -- so it runs in
a CHILD process (oracle/ref_gpu.py); tensors cross as files.  Same seeded inputs on both sides:

  * quantisers: the reference's own bar against its test oracle is a mismatch fraction <= 1e-4 for MX
    (tests/mxfp4_test.py:221) and <= 1e-1 for NV (tests/nvfp4_test.py:205); ours against the same oracle is <= 3e-6,
    so ours against the reference's KERNEL must stay within 2e-4 (MX) / 1e-2 (NV) of the dequantised values;
  * GEMM on IDENTICAL quantised operands: both sides are one fp32 tcgen05 accumulation chain over K in the same order,
    alpha in fp32, one RNE to bf16 -> bit-exact.

Observed on a B200 (profiles/r01_ref_parity_first_run.log, r01_ref_quant_diag.jsonl, r02_first_call.md): both GEMMs
bit-identical to the reference's CUTLASS kernels on every shape; every quantiser case IDENTICAL in dequantised values and
scale bytes (one +0 / -0 code in a million differs).  NVFP4 abs_max with Hadamard-128 is the one case the reference
dispatches on sm_100 to a kernel of its own (bindings.cpp:413-415 -> fused_quantize_nv_sm100.cu) that stores the
e4m3-ROUNDED scale but computes the codes with the UNROUNDED one
(cutlass_extensions/epilogue/fusion/sm100_visitor_store_tma_warpspecialized.hpp:141-148,567-591), unlike its own mma.sync
kernels for H = 16/32/64 and for sm_120 (epilogue_quant.h:1664-1692) and unlike its test oracle (tests/nvfp4_test.py:132-170)
-- which is why the reference's NVFP4 bar is 1e-1.  Round 2 decision: this library is sm_100-only, so it reproduces THAT
kernel by default (same 2e-4 / 1e-2 bars as every other case); B200Q_NV128_ORACLE_CODES=1 / B200Q_NV_ORACLE_CODES selects the
oracle / mma.sync arithmetic (scales identical, 4.8 % of the dequantised values differ).  See DESIGN.md section 4.
Infrastructure trouble (library not built, child cannot start) is a skip, never a failure.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

import oracle as O
import helpers as H
from oracle import ref_gpu

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA required", allow_module_level=True)

import qutlass_b200 as Q  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference(cases, timeout=240):
    """run `cases` through the reference library in a child process; returns its list of result dicts"""
    if not ref_gpu.available():
        pytest.skip("oracle/_ref/qutlass_ref_C.so not built (python oracle/build_ref.py in the build container)")
    with tempfile.TemporaryDirectory() as wd:
        torch.save({"cases": cases}, os.path.join(wd, "job.pt"))
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_gpu.py"), "parity", wd],
                               capture_output=True, text=True, timeout=timeout)
        except subprocess.TimeoutExpired:
            pytest.skip("reference child process timed out")
        if r.returncode != 0 or not os.path.exists(os.path.join(wd, "out.pt")):
            pytest.skip("reference child process failed: " + (r.stderr or r.stdout)[-400:])
        return torch.load(os.path.join(wd, "out.pt"))["results"]


def _bf16_cpu(x: np.ndarray) -> torch.Tensor:
    return H.bf16_tensor_from_f32(x, device="cpu")


# 512 rows: our butterfly kernel; 2048 rows at H = 128: our tcgen05 rotation kernel.  On the reference's side the two
# Had-128 abs_max cases run ITS sm_100 tcgen05 kernels (fused_quantize_{mx,nv}_sm100.cu), all others its mma.sync kernels.
QUANT_CASES = [("mx", "abs_max", 32, 1.0, 512), ("mx", "abs_max", 128, 1.0, 512), ("mx", "quest", 64, 1.0, 512),
               ("mx", "quest", 128, 1.0, 512), ("mx", "abs_max", 128, 1.0, 2048), ("nv", "abs_max", 16, 6.0, 512),
               ("nv", "abs_max", 128, 6.0, 512), ("nv", "quest", 64, 6.0, 512)]


def test_quantisers_match_the_reference_kernels():
    k = 4096
    cases, ours = [], []
    for fmt, method, had, gs, rows in QUANT_CASES:
        x = H.random_bf16((rows, k), seed=1000 + had + len(cases))         # randn * 25, like the reference's tests
        R = O.hadamard_matrix(had)
        cases.append({"op": "quantize", "fmt": fmt, "method": method, "x": _bf16_cpu(x), "R": _bf16_cpu(R), "gs": gs})
        xt, Rt = H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R)
        if fmt == "mx":
            q, sf = Q.fusedQuantizeMx(xt, Rt, method=method)
        else:
            q, sf = Q.fusedQuantizeNv(xt, Rt, torch.tensor([gs], device="cuda"), method=method)
        torch.cuda.synchronize()
        ours.append((H.u8_of(q), H.u8_of(sf)))
    ref = _reference(cases)
    report = []
    for (fmt, method, had, gs, rows), (q, sf), r in zip(QUANT_CASES, ours, ref):
        cols = k // (32 if fmt == "mx" else 16)
        sf_o = sf.reshape(-1, sf.shape[-1])[:rows, :cols]
        sf_r = r["sf"].numpy()[:rows, :cols]
        dq = O.dequant_mx if fmt == "mx" else O.dequant_nv
        mism = float((dq(q, sf_o) != dq(r["q"].numpy().reshape(rows, -1), sf_r)).mean())
        sf_mism = float((sf_o != sf_r).mean())
        # NVFP4 Had-128 abs_max: the reference's sm_100-only kernel; reproduced by default, so it gets the tight bar too
        tight_nv = (fmt, method, had) == ("nv", "abs_max", 128)
        bar = 2e-4 if (fmt == "mx" or tight_nv) else 1e-2
        report.append((fmt, method, had, rows, mism, sf_mism, bar))
    bad = [r for r in report if r[4] > r[6] or r[5] > 1e-3]
    assert not bad, report


@pytest.mark.parametrize("fmt", ["mx", "nv"])
def test_gemm_is_bit_identical_to_the_reference_kernel(fmt):
    """the reference's bit-exact shapes (tests/mxfp4_test.py:223-237,255-269; nvfp4_test.py:207-224) + a Llama FFN slice,
    on operands quantised by OUR kernels, multiplied by both GEMMs"""
    had = 32
    R = H.bf16_tensor_from_f32(O.hadamard_matrix(had))
    gs = torch.tensor([1.0], device="cuda")
    cases, ours = [], []
    for i, (m, n, k) in enumerate([(1, 504, 4096), (504, 504, 2048), (16, 14336, 4096), (512, 1024, 4096)]):
        a = H.bf16_tensor_from_f32(H.random_bf16((m, k), seed=70 + i))
        b = H.bf16_tensor_from_f32(H.random_bf16((n, k), seed=80 + i))
        if fmt == "mx":
            (aq, asf), (bq, bsf) = Q.fusedQuantizeMx(a, R, method="abs_max"), Q.fusedQuantizeMx(b, R, method="abs_max")
            mm = Q.matmul_mxf4_bf16_tn
        else:
            (aq, asf), (bq, bsf) = Q.fusedQuantizeNv(a, R, gs, method="abs_max"), Q.fusedQuantizeNv(b, R, gs, method="abs_max")
            mm = Q.matmul_nvf4_bf16_tn
        a_blk, b_blk = Q.to_blocked(asf), Q.to_blocked(bsf)
        alpha = 1.0 / 9.0
        d = mm(aq, bq, a_blk, b_blk, torch.tensor([alpha], device="cuda"))
        torch.cuda.synchronize()
        ours.append(H.bf16_bits_of(d))
        cases.append({"op": "gemm", "fmt": fmt, "a": aq.cpu(), "b": bq.cpu(), "a_sf": a_blk.view(torch.uint8).cpu(),
                      "b_sf": b_blk.view(torch.uint8).cpu(), "alpha": alpha})
    ref = _reference(cases)
    for got, r, c in zip(ours, ref, cases):
        want = r["d"].numpy().view(np.uint16)
        mism, rel = H.compare_bits(got, want)
        assert mism == 0.0, (fmt, tuple(c["a"].shape), tuple(c["b"].shape), mism, rel)


# ----------------------------------------------------------------------------- wider comparisons
# First executed by the driver at the end of round 1 (all four XPASSed) and again at the start of round 2
# (profiles/r02_first_call.md): strict since.
def test_remaining_quantiser_sizes_match_the_reference_kernels():
    k, rows = 2048, 256
    todo = [("mx", "quest", 32, 1.0), ("mx", "abs_max", 64, 1.0), ("nv", "abs_max", 32, 6.0), ("nv", "abs_max", 64, 6.0),
            ("nv", "quest", 16, 6.0), ("nv", "quest", 32, 6.0), ("nv", "quest", 128, 6.0), ("nv", "abs_max", 64, 1.0)]
    cases, ours = [], []
    for i, (fmt, method, had, gs) in enumerate(todo):
        x = H.random_bf16((rows, k), seed=2000 + i)
        R = O.hadamard_matrix(had)
        cases.append({"op": "quantize", "fmt": fmt, "method": method, "x": _bf16_cpu(x), "R": _bf16_cpu(R), "gs": gs})
        xt, Rt = H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R)
        q, sf = (Q.fusedQuantizeMx(xt, Rt, method=method) if fmt == "mx" else
                 Q.fusedQuantizeNv(xt, Rt, torch.tensor([gs], device="cuda"), method=method))
        torch.cuda.synchronize()
        ours.append((H.u8_of(q), H.u8_of(sf)))
    ref = _reference(cases)
    report = []
    for (fmt, method, had, gs), (q, sf), r in zip(todo, ours, ref):
        cols = k // (32 if fmt == "mx" else 16)
        sf_o = sf.reshape(-1, sf.shape[-1])[:rows, :cols]
        sf_r = r["sf"].numpy()[:rows, :cols]
        dq = O.dequant_mx if fmt == "mx" else O.dequant_nv
        mism = float((dq(q, sf_o) != dq(r["q"].numpy().reshape(rows, -1), sf_r)).mean())
        report.append((fmt, method, had, gs, mism, float((sf_o != sf_r).mean())))
    assert all(m <= (2e-4 if f == "mx" else 1e-2) and s <= 1e-3 for f, _, _, _, m, s in report), report


def test_clip_mask_matches_the_reference_kernel():
    """fusedQuantizeMx(method="quest", return_mask=True): codes, scales and the packed clip mask (Hadamard-32: the only size
    the reference's mask kernel supports, bindings.cpp:277-286)"""
    rows, k, had = 256, 2048, 32
    x = H.random_bf16((rows, k), seed=2100)
    R = O.hadamard_matrix(had)
    q, sf, mask = Q.fusedQuantizeMx(H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R), method="quest", return_mask=True)
    torch.cuda.synchronize()
    r = _reference([{"op": "quantize_mask", "x": _bf16_cpu(x), "R": _bf16_cpu(R)}])[0]
    cols = k // 32
    sf_o = H.u8_of(sf).reshape(-1, sf.shape[-1])[:rows, :cols]
    sf_r = r["sf"].numpy()[:rows, :cols]
    assert float((sf_o != sf_r).mean()) <= 1e-4
    assert float((O.dequant_mx(H.u8_of(q), sf_o) != O.dequant_mx(r["q"].numpy().reshape(rows, -1), sf_r)).mean()) <= 2e-4
    assert float((H.u8_of(mask).reshape(-1) != r["mask"].numpy().reshape(-1)).mean()) <= 2e-4


@pytest.mark.parametrize("nn", [False, True])
def test_mxfp8_gemm_is_bit_identical_to_the_reference_kernel(nn):
    cases, ours = [], []
    for i, (m, n, k) in enumerate([(16, 1024, 4096), (496, 512, 2048), (1024, 1536, 512)]):
        aq, asf = H.random_f8_operand(m, k, seed=300 + i)
        bq, bsf = H.random_f8_operand(n, k, seed=400 + i)
        a_blk, b_blk = H.blocked_sf(asf), H.blocked_sf(bsf)
        a_np = np.ascontiguousarray(aq.T) if nn else aq                      # nn: A stored [K, M]
        a = torch.from_numpy(a_np).cuda().view(torch.float8_e4m3fn)
        b = torch.from_numpy(bq).cuda().view(torch.float8_e4m3fn)
        mm = Q.matmul_mxf8_bf16_nn if nn else Q.matmul_mxf8_bf16_tn
        d = mm(a, b, H.sf_torch(a_blk, "mx"), H.sf_torch(b_blk, "mx"), torch.tensor([0.5], device="cuda"))
        torch.cuda.synchronize()
        ours.append(H.bf16_bits_of(d))
        cases.append({"op": "gemm_f8", "nn": nn, "a": torch.from_numpy(a_np), "b": torch.from_numpy(bq),
                      "a_sf": torch.from_numpy(a_blk), "b_sf": torch.from_numpy(b_blk), "alpha": 0.5})
    ref = _reference(cases)
    for got, r, c in zip(ours, ref, cases):
        mism, rel = H.compare_bits(got, r["d"].numpy().view(np.uint16))
        assert mism == 0.0, (nn, tuple(c["a"].shape), tuple(c["b"].shape), mism, rel)


@pytest.mark.parametrize("rows", [512, 2048, 64])
def test_nv128_abs_max_default_is_the_reference_sm100_kernel_and_the_switch_selects_the_oracle(rows, monkeypatch):
    """NVFP4 abs_max Hadamard-128: by default bit-compatible with the reference's sm_100-only kernel (codes from the
    unrounded scale) on every one of our kernels -- 512 rows: butterfly, 2048 rows: tcgen05, 64 rows with a random
    (non-Hadamard) rotation: mma.sync -- and B200Q_NV128_ORACLE_CODES=1 gives the reference ORACLE's arithmetic instead."""
    k, had, gs = 4096, 128, 6.0
    x = H.random_bf16((rows, k), seed=1134 + rows)
    R = O.hadamard_matrix(had)
    if rows == 64:
        R = O.bf16_round(np.random.default_rng(5).standard_normal((had, had)).astype(np.float32) * had ** -0.5)
    cols = k // 16
    xt, Rt, gst = H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R), torch.tensor([gs], device="cuda")
    q, sf = Q.fusedQuantizeNv(xt, Rt, gst, method="abs_max")
    torch.cuda.synchronize()
    ref = _reference([{"op": "quantize", "fmt": "nv", "method": "abs_max", "x": _bf16_cpu(x), "R": _bf16_cpu(R), "gs": gs}])[0]
    sf_o = H.u8_of(sf).reshape(-1, sf.shape[-1])[:rows, :cols]
    sf_r = ref["sf"].numpy()[:rows, :cols]
    # scale bytes: identical up to fp32 summation order of the rotation (butterfly vs tensor-core accumulation: an amax that
    # sits on an e4m3 rounding boundary can land one code apart -- 3 of 131072 observed)
    assert float((sf_o != sf_r).mean()) <= 1e-4
    dq_ref = O.dequant_nv(ref["q"].numpy().reshape(rows, -1), sf_r)
    assert float((O.dequant_nv(H.u8_of(q), sf_o) != dq_ref).mean()) <= 2e-4
    want = O.quantize_nv(x, R, gs, "abs_max")                         # the oracle's default follows the sm_100 dispatch
    assert float((O.dequant_nv(want["q"].reshape(rows, -1), want["sf"].reshape(rows, cols)) != dq_ref).mean()) <= 1e-2
    monkeypatch.setenv("B200Q_NV128_ORACLE_CODES", "1")
    q2, sf2 = Q.fusedQuantizeNv(xt, Rt, gst, method="abs_max")
    torch.cuda.synchronize()
    assert float((H.u8_of(sf2).reshape(-1, sf2.shape[-1])[:rows, :cols] != sf_r).mean()) <= 1e-4
    want2 = O.quantize_nv(x, R, gs, "abs_max", sm100_codes=False)
    dq2 = O.dequant_nv(H.u8_of(q2), sf_o)
    assert float((dq2 != O.dequant_nv(want2["q"].reshape(rows, -1), want2["sf"].reshape(rows, cols))).mean()) <= 1e-2
    assert 0.03 <= float((dq2 != dq_ref).mean()) <= 0.06             # the documented 4.8 % divergence
