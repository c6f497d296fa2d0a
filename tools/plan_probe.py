#!/usr/bin/env python
"""Planner check: for each M time the automatic configuration against every explicit one (CUDA-graph replay, MXFP4/NVFP4,
N=14336 K=4096 and the 70B shape) -> which (cta_group, block_n) the heuristic should pick."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
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

def run(kind, N, K, Ms):
    knd = 0 if kind == "mx" else 1
    group = 32 if kind == "mx" else 16
    b = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev)
    lo, hi = (126, 129) if kind == "mx" else (0x30, 0x41)
    sfb = torch.randint(lo, hi, (((N + 127) // 128) * 128 * (K // group),), dtype=torch.uint8, device=dev)
    alpha = torch.ones(1, device=dev)
    for M in Ms:
        a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
        sfa = torch.randint(lo, hi, (((M + 127) // 128) * 128 * (K // group),), dtype=torch.uint8, device=dev)
        d = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        rec = dict(kind=kind, N=N, K=K, M=M)
        for (cg, bn) in ((0, 0), (1, 128), (1, 256), (2, 128), (2, 192), (2, 256)):
            def gemm(st):
                rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(), d.data_ptr(), M, N, K, knd, cg, bn, st)
                assert rc == 0, lib.b200q_last_error()
            rec[f"{cg}x{bn}"] = round(graph_time(gemm, 20 if M <= 4096 else 6), 2)
        best = min((v, k) for k, v in rec.items() if "x" in k and k != "0x0")
        rec["best"] = best[1]; rec["auto_vs_best"] = round(rec["0x0"] / best[0], 3)
        print(json.dumps(rec), flush=True)

if __name__ == "__main__":
    run("mx", 14336, 4096, [192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 8192, 16384])
    run("nv", 14336, 4096, [256, 1024, 4096])
    run("mx", 28672, 8192, [512, 2048])
    run("mx", 4096, 14336, [1024, 4096])
    run("mx", 6144, 4096, [1024, 4096])
