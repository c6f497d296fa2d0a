#!/usr/bin/env python
"""Runs the UNMODIFIED reference kernels (oracle/_ref/qutlass_ref_C.so, built by oracle/build_ref.py from
/root/reference) on a GPU -- in its OWN process.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference registers its ops in the torch library
`_qutlass_C` (qutlass/csrc/bindings.cpp:498-526) -- the same namespace qutlass_b200 registers its drop-in ops in -- so
the two cannot live in one process.  tests/test_gpu_reference_lib.py and bench.py start this file as a child:

  python oracle/ref_gpu.py parity <workdir>     reads <workdir>/job.pt, writes <workdir>/out.pt
  python oracle/ref_gpu.py bench <mx|nv> <M> <N> <K> <had> <steps>       prints ONE JSON line

The calls below are the reference's Python wrappers restated as plain op calls (qutlass/__init__.py:34-43,89-98,
149-203: output allocation + `qutlass_CUDA.<op>`); /root/reference itself does not exist on the GPU box, only the
compiled library travels.  This file never imports qutlass_b200.
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "qutlass_ref_C.so")


def available() -> bool:
    return os.path.exists(LIB)


def _load():
    import torch
    torch.ops.load_library(LIB)
    return torch, torch.ops._qutlass_C


def _padded(rows: int, cols: int):
    return (rows + 127) // 128 * 128, (cols + 3) // 4 * 4          # qutlass/utils.py:136-157


def _quantize(torch, ops, fmt, method, x, R, gs):
    """fusedQuantizeMx / fusedQuantizeNv of the reference (qutlass/__init__.py:149-203)."""
    group = 32 if fmt == "mx" else 16
    rows, k = x.numel() // x.size(-1), x.size(-1)
    pr, pc = _padded(rows, k // group)
    q = torch.empty(*x.shape[:-1], k // 2, dtype=torch.uint8, device=x.device)
    sf = torch.empty(pr, pc, dtype=torch.float8_e8m0fnu if fmt == "mx" else torch.float8_e4m3fn, device=x.device)
    if fmt == "mx":
        op = ops.fusedQuantizeMxQuest if method == "quest" else ops.fusedQuantizeMxAbsMax
        op(x, R, q, sf)
    else:
        op = ops.fusedQuantizeNvQuest if method == "quest" else ops.fusedQuantizeNvAbsMax
        op(x, R, q, sf, gs)
    return q, sf


def _to_blocked(torch, sf):
    """row-major padded scales [R, C] (R % 128 == 0, C % 4 == 0) -> flat block-scaled layout, from the layout formula
    off(r, c) = ((r/128)*(C/4) + c/4)*512 + (r%32)*16 + ((r%128)/32)*4 + c%4  (SURVEY.md section 8a, row a4)."""
    r, c = sf.shape
    u = sf.view(torch.uint8).view(r // 128, 4, 32, c // 4, 4)          # [rb, r/32 % 4, r % 32, cb, c % 4]
    return u.permute(0, 3, 2, 1, 4).contiguous().view(-1).view(sf.dtype)


def run_parity(workdir: str) -> None:
    torch, ops = _load()
    job = torch.load(os.path.join(workdir, "job.pt"))
    dev = torch.device("cuda", 0)
    out = []
    for case in job["cases"]:
        if case["op"] == "quantize":
            x, R = case["x"].to(dev), case["R"].to(dev)
            gs = torch.tensor([case.get("gs", 1.0)], dtype=torch.float32, device=dev)
            q, sf = _quantize(torch, ops, case["fmt"], case["method"], x, R, gs)
            torch.cuda.synchronize()
            out.append({"q": q.cpu(), "sf": sf.view(torch.uint8).cpu()})
        elif case["op"] == "quantize_mask":
            # fusedQuantizeMx(..., method="quest", return_mask=True) (qutlass/__init__.py:166-172)
            x, R = case["x"].to(dev), case["R"].to(dev)
            rows, k = x.numel() // x.size(-1), x.size(-1)
            pr, pc = _padded(rows, k // 32)
            q = torch.empty(*x.shape[:-1], k // 2, dtype=torch.uint8, device=dev)
            sf = torch.empty(pr, pc, dtype=torch.float8_e8m0fnu, device=dev)
            mask = torch.empty(*x.shape[:-1], k // 8, dtype=torch.uint8, device=dev)
            ops.fusedQuantizeMxQuestWithMask(x, R, q, sf, mask)
            torch.cuda.synchronize()
            out.append({"q": q.cpu(), "sf": sf.view(torch.uint8).cpu(), "mask": mask.cpu()})
        elif case["op"] == "gemm_f8":
            # matmul_mxf8_bf16_tn / _nn (qutlass/__init__.py:134-146): e4m3 operands, e8m0 blocked scales
            a = case["a"].to(dev).view(torch.float8_e4m3fn)
            b = case["b"].to(dev).view(torch.float8_e4m3fn)
            a_sf = case["a_sf"].to(dev).view(torch.float8_e8m0fnu)
            b_sf = case["b_sf"].to(dev).view(torch.float8_e8m0fnu)
            alpha = torch.tensor([case["alpha"]], dtype=torch.float32, device=dev)
            op = ops.matmul_mxf8_bf16_nn if case.get("nn") else ops.matmul_mxf8_bf16_tn
            d = op(a, b, a_sf, b_sf, alpha)
            torch.cuda.synchronize()
            out.append({"d": d.view(torch.int16).cpu()})
        elif case["op"] == "gemm":
            sf_dt = torch.float8_e8m0fnu if case["fmt"] == "mx" else torch.float8_e4m3fn
            a, b = case["a"].to(dev), case["b"].to(dev)
            a_sf, b_sf = case["a_sf"].to(dev).view(sf_dt), case["b_sf"].to(dev).view(sf_dt)
            alpha = torch.tensor([case["alpha"]], dtype=torch.float32, device=dev)
            op = ops.matmul_mxf4_bf16_tn if case["fmt"] == "mx" else ops.matmul_nvf4_bf16_tn
            d = op(a, b, a_sf, b_sf, alpha)
            torch.cuda.synchronize()
            out.append({"d": d.view(torch.int16).cpu()})
        else:
            raise ValueError(case["op"])
    torch.save({"results": out}, os.path.join(workdir, "out.pt"))


def run_bench(fmt: str, M: int, N: int, K: int, had: int, steps: int) -> None:
    """The reference's own benchmark iteration (benchmarks/bench_mxfp4_sm100.py:93-106) on synthetic data of the bench
    shape: GEMM alone, quantise alone and quantise + GEMM (the to_blocked launch between them is NOT timed: the
    reference does it with a Triton kernel that lives in its Python package -- this favours the reference)."""
    torch, ops = _load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    idx = torch.arange(had)
    bits = idx[:, None] & idx[None, :]
    par = torch.zeros_like(bits)
    while bits.any():
        par ^= bits & 1
        bits = bits >> 1
    R = ((1.0 - 2.0 * par.double()) * had ** -0.5).to(torch.bfloat16).to(dev)
    gs = torch.tensor([1.0], dtype=torch.float32, device=dev)
    alpha = torch.tensor([1.0], dtype=torch.float32, device=dev)
    NS = 4                                               # rotating sets: footprint > L2, like bench.py
    w = torch.randn(N, K, dtype=torch.bfloat16, device=dev)
    wq, wsf = _quantize(torch, ops, fmt, "abs_max", w, R, gs)
    wqs = [wq.clone() for _ in range(NS)]
    wsfs = [_to_blocked(torch, wsf).clone() for _ in range(NS)]
    del w
    acts = [torch.randn(M, K, dtype=torch.bfloat16, device=dev) for _ in range(NS)]
    aq0, asf0 = _quantize(torch, ops, fmt, "abs_max", acts[0], R, gs)
    aqs = [aq0.clone() for _ in range(NS)]
    asf_rm = [asf0.clone() for _ in range(NS)]
    asf_blk = [_to_blocked(torch, asf0).clone() for _ in range(NS)]
    gemm_op = ops.matmul_mxf4_bf16_tn if fmt == "mx" else ops.matmul_nvf4_bf16_tn
    q_op = ops.fusedQuantizeMxAbsMax if fmt == "mx" else ops.fusedQuantizeNvAbsMax

    def quant(i):
        s = i % NS
        if fmt == "mx":
            q_op(acts[s], R, aqs[s], asf_rm[s])
        else:
            q_op(acts[s], R, aqs[s], asf_rm[s], gs)

    def gemm(i):
        s = i % NS
        return gemm_op(aqs[s], wqs[s], asf_blk[s], wsfs[s], alpha)

    def step(i):
        quant(i)
        return gemm(i)

    def timed(fn, n, warm=10):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    flops = 2.0 * M * N * K
    us_gemm = timed(gemm, steps)
    us_quant = timed(quant, steps)
    us_step = timed(step, steps)
    print(json.dumps({
        "impl": "reference CUTLASS kernels (oracle/_ref/qutlass_ref_C.so, unmodified sources, child process)",
        "kind": fmt, "M": M, "N": N, "K": K, "had": had, "steps": steps,
        "gemm_us": us_gemm, "gemm_tflops": flops / us_gemm / 1e6,
        "quantize_us": us_quant,
        "step_us_without_to_blocked": us_step, "step_tflops_without_to_blocked": flops / us_step / 1e6,
        "note": "eager op calls (the reference allocates D per call), CUDA events, rotating 4 buffer sets; to_blocked not timed",
    }), flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "parity":
        run_parity(sys.argv[2])
    elif len(sys.argv) >= 8 and sys.argv[1] == "bench":
        run_bench(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]))
    else:
        sys.exit(__doc__)
