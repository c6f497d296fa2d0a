#!/usr/bin/env python
"""Head-to-head M sweep on one B200: OUR kernels (C-ABI) vs the reference's own kernels (oracle/_ref/qutlass_ref_C.so =
the unmodified reference compiled by oracle/build_ref.py), same process, same buffers, CUDA-graph replay like the
reference's benchmarks (benchmarks/bench_mxfp4_sm100.py:213-226).  The reference library is loaded FIRST so its ops own the
`_qutlass_C` namespace; qutlass_b200's Python functions call the C-ABI directly and do not need their registration.

Per (kind, M): GEMM alone ("ideal"), quantise alone, quantise + GEMM ("actual").  For the reference the "actual" step also
runs a to_blocked stand-in (the torch view/permute/contiguous of oracle/ref_gpu.py -- one copy kernel, standing in for the
Triton kernel of its Python package, which does not travel); ours writes the blocked scales in the quantiser.
Writes gpurun_out/ref_msweep.md and prints JSON lines.  Diagnostic / profiling tool (uses oracle/: not product code)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref_gpu
torch.ops.load_library(ref_gpu.LIB)
ops = torch.ops._qutlass_C
import qutlass_b200 as Q          # noqa: E402  (its op registration is skipped: the names are taken)
from qutlass_b200 import _lib     # noqa: E402
lib = _lib.load()
dev = torch.device("cuda")


def graph_time(fn, iters):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


def hadamard(h):
    idx = torch.arange(h)
    bits = idx[:, None] & idx[None, :]
    par = torch.zeros_like(bits)
    while bits.any():
        par ^= bits & 1
        bits = bits >> 1
    return ((1.0 - 2.0 * par.double()) * h ** -0.5).to(torch.bfloat16).to(dev)


rows_out = []


def sweep(kind, N, K, Ms, had=128):
    R = hadamard(had)
    gs = torch.ones(1, device=dev)
    alpha = torch.ones(1, device=dev)
    w = torch.randn(N, K, dtype=torch.bfloat16, device=dev)
    if kind == "mx":
        wq, wsf = Q.fusedQuantizeMx(w, R, method="abs_max")
        ours_q = lambda x: Q.fusedQuantizeMx(x, R, method="abs_max")                      # noqa: E731
        ours_mm, ref_mm, ref_qop = Q.matmul_mxf4_bf16_tn, ops.matmul_mxf4_bf16_tn, ops.fusedQuantizeMxAbsMax
    else:
        wq, wsf = Q.fusedQuantizeNv(w, R, gs, method="abs_max")
        ours_q = lambda x: Q.fusedQuantizeNv(x, R, gs, method="abs_max")                  # noqa: E731
        ours_mm, ref_mm, ref_qop = Q.matmul_nvf4_bf16_tn, ops.matmul_nvf4_bf16_tn, ops.fusedQuantizeNvAbsMax
    wblk = Q.to_blocked(wsf)
    del w
    for M in Ms:
        x = torch.randn(M, K, dtype=torch.bfloat16, device=dev)
        aq, asf = ours_q(x)
        ablk = Q.to_blocked(asf)
        rq, rsf = ref_gpu._quantize(torch, ops, kind, "abs_max", x, R, gs)
        rblk = ref_gpu._to_blocked(torch, rsf)
        torch.cuda.synchronize()
        iters = 20 if M <= 4096 else 6

        def o_gemm():
            return ours_mm(aq, wq, ablk, wblk, alpha)

        def o_quant():
            return ours_q(x)

        def o_both():
            q_, s_ = ours_q(x)
            return ours_mm(q_, wq, Q.to_blocked(s_), wblk, alpha)

        def o_both_static():          # a serving stack: weights quantised once -> static_weights=True (b200q.h)
            q_, s_ = ours_q(x)
            return ours_mm(q_, wq, Q.to_blocked(s_), wblk, alpha, static_weights=True)

        def r_gemm():
            return ref_mm(rq, wq, rblk, wblk, alpha)

        def r_quant():
            return ref_qop(x, R, rq, rsf) if kind == "mx" else ref_qop(x, R, rq, rsf, gs)

        def r_both():
            r_quant()
            return ref_mm(rq, wq, ref_gpu._to_blocked(torch, rsf), wblk, alpha)

        rec = dict(kind=kind, N=N, K=K, M=M)
        for name, fn in (("ours_gemm", o_gemm), ("ref_gemm", r_gemm), ("ours_quant", o_quant), ("ref_quant", r_quant),
                         ("ours_both", o_both), ("ours_both_static", o_both_static), ("ref_both", r_both)):
            try:
                rec[name + "_us"] = round(graph_time(fn, iters), 2)
            except Exception as e:   # noqa: BLE001
                rec[name + "_us"] = None
                rec[name + "_error"] = f"{type(e).__name__}: {e}"[:200]
        fl = 2.0 * M * N * K
        for side in ("ours", "ref"):
            t = rec.get(side + "_gemm_us")
            rec[side + "_gemm_tflops"] = round(fl / t / 1e6, 1) if t else None
        rows_out.append(rec)
        print(json.dumps(rec), flush=True)


def sweep_f8(N, K, Ms):
    """MXFP8 GEMM (e4m3 operands, e8m0 scales per 32): ours vs the reference's matmul_mxf8_bf16_tn kernel (gemm.cu:328-380)."""
    alpha = torch.ones(1, device=dev)
    w = torch.randint(0, 120, (N, K), dtype=torch.uint8, device=dev).view(torch.float8_e4m3fn)
    wsf = torch.randint(126, 129, (((N + 127) // 128) * 128 * (K // 32),), dtype=torch.uint8, device=dev).view(torch.float8_e8m0fnu)
    for M in Ms:
        a = torch.randint(0, 120, (M, K), dtype=torch.uint8, device=dev).view(torch.float8_e4m3fn)
        asf = torch.randint(126, 129, (((M + 127) // 128) * 128 * (K // 32),), dtype=torch.uint8, device=dev).view(torch.float8_e8m0fnu)
        iters = 20 if M <= 4096 else 6
        rec = dict(kind="mxf8", N=N, K=K, M=M)
        for name, fn in (("ours_gemm", lambda: Q.matmul_mxf8_bf16_tn(a, w, asf, wsf, alpha)),
                         ("ref_gemm", lambda: ops.matmul_mxf8_bf16_tn(a, w, asf, wsf, alpha))):
            try:
                rec[name + "_us"] = round(graph_time(fn, iters), 2)
            except Exception as e:   # noqa: BLE001
                rec[name + "_us"] = None
                rec[name + "_error"] = f"{type(e).__name__}: {e}"[:200]
        fl = 2.0 * M * N * K
        for side in ("ours", "ref"):
            t = rec.get(side + "_gemm_us")
            rec[side + "_gemm_tflops"] = round(fl / t / 1e6, 1) if t else None
        rows_out.append(rec)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    if os.environ.get("MSWEEP_70B_SMALL"):
        # the reference benchmark's own layer (Llama-3.1-70B: K = 8192, N = 57344) at the batch sizes where its tables were tightest
        sweep("mx", 57344, 8192, [1, 16, 32, 64, 128, 256, 512])
        sys.exit(0)
    Ms = [1, 16, 128, 1024, 4096, 16384]
    sweep("mx", 14336, 4096, Ms)
    sweep("nv", 14336, 4096, Ms)
    sweep("mx", 28672, 8192, [2048, 16384])
    sweep_f8(14336, 4096, [16, 1024, 4096, 16384])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_msweep.md"), "w") as f:
        f.write("# ours vs the compiled reference, one B200, CUDA-graph replay, best of 3 (us)\n\n")
        f.write("| kind | N | K | M | GEMM ours | GEMM ref | quantise ours | quantise ref | quant+GEMM ours | ours, static weights | quant+GEMM ref (+to_blocked stand-in) |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows_out:
            f.write(f"| {r['kind']} | {r['N']} | {r['K']} | {r['M']} | {r.get('ours_gemm_us')} | {r.get('ref_gemm_us')} | "
                    f"{r.get('ours_quant_us')} | {r.get('ref_quant_us')} | {r.get('ours_both_us')} | {r.get('ours_both_static_us')} | {r.get('ref_both_us')} |\n")
