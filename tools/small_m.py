#!/usr/bin/env python
"""Run the small-M (decode) GEMM a few times (for ncu) and time a few configurations."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
N, K = 14336, 4096
def run(M, cg, bn, iters=30, kind=0):
    a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
    b = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev)
    sfa = torch.randint(126, 129, (((M + 127) // 128) * 128 * (K // 32),), dtype=torch.uint8, device=dev)
    sfb = torch.randint(126, 129, (N * (K // 32),), dtype=torch.uint8, device=dev)
    d = torch.empty(M, N, dtype=torch.bfloat16, device=dev); alpha = torch.ones(1, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def go(stream=st):
        rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(), d.data_ptr(), M, N, K, kind, cg, bn, stream)
        assert rc == 0, lib.b200q_last_error()
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): go()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    # graph-replayed timing (what the reference benchmarks use)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        go(s.cuda_stream)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters): go(s.cuda_stream)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us_graph = e0.elapsed_time(e1) / iters * 1e3
    print(json.dumps(dict(M=M, cg=cg, bn=bn, us_eager=round(us, 2), us_graph=round(us_graph, 2),
                          gbs_graph=round((N * K / 2 + N * K / 32) / us_graph / 1e3, 0))), flush=True)
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        M, cg, bn = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        a = None
        run(M, cg, bn, iters=5)
    else:
        for M in (1, 16, 128):
            for (cg, bn) in ((1, 64), (1, 128), (1, 192), (1, 256)):
                run(M, cg, bn)
