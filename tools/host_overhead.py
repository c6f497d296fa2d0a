#!/usr/bin/env python
"""Eager per-call cost of the public API at decode size (M=1, N=14336, K=4096) for WHICHEVER `qutlass` package PYTHONPATH
resolves -- run it twice on the same box:
    PYTHONPATH=<repo> python tools/host_overhead.py ours            (the drop-in: qutlass -> qutlass_b200)
    PYTHONPATH=<repo>/oracle/_ref/ref_pkg:<repo>/oracle/ref_suite_shims python tools/host_overhead.py reference
Two figures per call sequence: `eager_us` = wall time per call of a tight loop of 2000 calls (max of host and GPU time,
what a non-graph caller sees) and `host_us` = wall time per call while CAPTURING into a CUDA graph (nothing executes on the
GPU during capture: pure host cost -- argument checks, allocation, dispatch, tensor-map encoding, launch)."""
import json, sys, time
import torch
from scipy.linalg import hadamard
import qutlass
from qutlass import matmul_mxf4_bf16_tn, fusedQuantizeMx
from qutlass.utils import to_blocked

label = sys.argv[1] if len(sys.argv) > 1 else "?"
dev = torch.device("cuda")
M, N, K = 1, 14336, 4096
H = torch.tensor(hadamard(128) * 128 ** -0.5, dtype=torch.bfloat16, device=dev)
a = torch.randn(M, K, dtype=torch.bfloat16, device=dev)
b = torch.randn(N, K, dtype=torch.bfloat16, device=dev)
wq, wsf = fusedQuantizeMx(b, H, method="abs_max")
wblk = to_blocked(wsf, True)
aq, asf = fusedQuantizeMx(a, H, method="abs_max")
ablk = to_blocked(asf, True)
alpha = torch.tensor([1.0], device=dev)


def seq_gemm():
    return matmul_mxf4_bf16_tn(aq, wq, ablk, wblk, alpha)


def seq_quant():
    return fusedQuantizeMx(a, H, method="abs_max")


def seq_full():
    q, s = fusedQuantizeMx(a, H, method="abs_max")
    return matmul_mxf4_bf16_tn(q, wq, to_blocked(s, True), wblk, alpha)


def eager(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def host_only(fn, n=300):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            dt = time.perf_counter() - t0
    return dt / n * 1e6


out = {"impl": label, "M": M, "N": N, "K": K, "package": getattr(qutlass, "__file__", "?")}
for name, fn in (("gemm", seq_gemm), ("quantize", seq_quant), ("quantize+to_blocked+gemm", seq_full)):
    out[name] = {"eager_us": round(eager(fn), 2), "host_us": round(host_only(fn), 2)}
print(json.dumps(out), flush=True)
