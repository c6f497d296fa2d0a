#!/usr/bin/env python
"""M sweep on Llama-3-8B / 70B FFN shapes (BASELINE.json configs 1-4), CUDA-graph replayed like the reference's
benchmarks (triton do_bench_cudagraph): GEMM only ("ideal") and quantise(H=128, abs_max)+GEMM ("actual").
Writes gpurun_out/msweep.md (markdown table) and prints JSON lines."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qutlass_b200 import _lib
import oracle as O
lib = _lib.load(); dev = torch.device("cuda")

def graph_time(fn, iters=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn(s.cuda_stream); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters): fn(s.cuda_stream)
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best

rows_out = []
def sweep(kind, N, K, Ms, had=128):
    knd = 0 if kind == "mx" else 1
    group = 32 if kind == "mx" else 16
    H = torch.from_numpy(O.bf16_bits(O.hadamard_matrix(had)).astype(np.int16)).view(torch.bfloat16).to(dev)
    b = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev)
    lo, hi = (126, 129) if kind == "mx" else (0x30, 0x41)
    sfb = torch.randint(lo, hi, (((N + 127) // 128) * 128 * (K // group),), dtype=torch.uint8, device=dev)
    alpha = torch.ones(1, device=dev); gs = torch.ones(1, device=dev)
    for M in Ms:
        x = torch.randn(M, K, dtype=torch.bfloat16, device=dev)
        a = torch.empty(M, K // 2, dtype=torch.uint8, device=dev)
        sfa = torch.zeros(((M + 127) // 128) * 128 * (K // group), dtype=torch.uint8, device=dev)
        d = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        def quant(st):
            if kind == "mx":
                rc = lib.b200q_quantize_mx(x.data_ptr(), H.data_ptr(), a.data_ptr(), None, sfa.data_ptr(), None, M * K, K, had, 1 | 0x100, st)
            else:
                rc = lib.b200q_quantize_nv(x.data_ptr(), H.data_ptr(), a.data_ptr(), None, sfa.data_ptr(), gs.data_ptr(), M * K, K, had, 1 | 0x100, st)
            assert rc == 0, lib.b200q_last_error()
        def gemm(st):
            rc = lib.b200q_gemm_fp4(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(), d.data_ptr(), M, N, K, knd, st)
            assert rc == 0, lib.b200q_last_error()
        def both(st):
            quant(st); gemm(st)
        quant(torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
        iters = 20 if M <= 4096 else 6
        tg = graph_time(gemm, iters); tq = graph_time(quant, iters); tb = graph_time(both, iters)
        fl = 2.0 * M * N * K
        rec = dict(kind=kind, N=N, K=K, M=M, gemm_us=round(tg, 2), quant_us=round(tq, 2), both_us=round(tb, 2),
                   gemm_tflops=round(fl / tg / 1e6, 1), actual_tflops=round(fl / tb / 1e6, 1))
        rows_out.append(rec); print(json.dumps(rec), flush=True)

if __name__ == "__main__":
    Ms = [1, 16, 128, 1024, 4096, 16384]
    sweep("mx", 14336, 4096, Ms)
    sweep("nv", 14336, 4096, Ms)
    sweep("mx", 28672, 8192, [2048, 4096, 8192, 16384])     # Llama-3-70B FFN: per-GPU shards of M=16384 over 8/4/2/1 GPUs
    # MXFP8 GEMM only ("next" row): e4m3 operands, ue8m0 scales per 32
    for M in (16, 1024, 4096):
        N, K = 14336, 4096
        a = torch.randint(0, 120, (M, K), dtype=torch.uint8, device=dev); b = torch.randint(0, 120, (N, K), dtype=torch.uint8, device=dev)
        sfa = torch.randint(126, 129, (((M + 127) // 128) * 128 * (K // 32),), dtype=torch.uint8, device=dev)
        sfb = torch.randint(126, 129, (N * (K // 32),), dtype=torch.uint8, device=dev)
        d = torch.empty(M, N, dtype=torch.bfloat16, device=dev); alpha = torch.ones(1, device=dev)
        def gemm8(st):
            rc = lib.b200q_gemm_fp4(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(), d.data_ptr(), M, N, K, 2, st)
            assert rc == 0, lib.b200q_last_error()
        t = graph_time(gemm8, 20)
        rec = dict(kind="mxf8", N=N, K=K, M=M, gemm_us=round(t, 2), quant_us=0, both_us=0, gemm_tflops=round(2.0 * M * N * K / t / 1e6, 1), actual_tflops=0)
        rows_out.append(rec); print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)     # gpurun merges this directory back; copy to profiles/ after review
    with open(os.path.join(ROOT, "gpurun_out", "msweep.md"), "w") as f:
        f.write("# M sweep (one B200, CUDA-graph replay, best of 3; quantise = Hadamard-128 abs_max)\n\n")
        f.write("| kind | N | K | M | GEMM us | GEMM TFLOP/s | quantise us | quant+GEMM us | quant+GEMM TFLOP/s |\n|---|---|---|---|---|---|---|---|---|\n")
        for r in rows_out:
            f.write(f"| {r['kind']} | {r['N']} | {r['K']} | {r['M']} | {r['gemm_us']} | {r['gemm_tflops']} | {r['quant_us']} | {r['both_us']} | {r['actual_tflops']} |\n")
