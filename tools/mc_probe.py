#!/usr/bin/env python
"""Multicast-cluster GEMM (cta_group=4: CTA pairs in clusters of four, A tiles multicast) vs the plain CTA-pair kernel:
bit equality of D and time.  python tools/mc_probe.py [parity] [time]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")

def mk(M, N, K, kind):
    group = 32 if kind == 0 else 16
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev, generator=g)
    b = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev, generator=g)
    lo, hi = (126, 129) if kind == 0 else (0x30, 0x41)
    nsa = ((M + 127) // 128) * 128 * (((K // group) + 3) // 4) * 4
    nsb = ((N + 127) // 128) * 128 * (((K // group) + 3) // 4) * 4
    sfa = torch.randint(lo, hi, (nsa,), dtype=torch.uint8, device=dev, generator=g)
    sfb = torch.randint(lo, hi, (nsb,), dtype=torch.uint8, device=dev, generator=g)
    return a, b, sfa, sfb

ALPHA = torch.ones(1, device=dev)

def gemm(t, M, N, K, kind, cg, bn, d, st=None):
    a, b, sfa, sfb = t
    rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), ALPHA.data_ptr(), d.data_ptr(),
                                M, N, K, kind, cg, bn, st or torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.b200q_last_error()

def parity():
    for kind in (0, 1):
        for (M, N, K) in ((256, 512, 512), (512, 1024, 1024), (1000, 1544, 2048), (4096, 14336, 4096), (300, 384, 256), (2048, 768, 4096)):
            t = mk(M, N, K, kind)
            for bn in (256, 192):
                d0 = torch.zeros(M, N, dtype=torch.bfloat16, device=dev); d1 = torch.zeros_like(d0)
                gemm(t, M, N, K, kind, 2, bn, d0); gemm(t, M, N, K, kind, 4, bn, d1)
                torch.cuda.synchronize()
                bad = int((d0.view(torch.int16) != d1.view(torch.int16)).sum())
                print(json.dumps(dict(check="mc_parity", kind=kind, M=M, N=N, K=K, bn=bn, mismatches=bad)), flush=True)

def timing():
    M, N, K = 4096, 14336, 4096
    for kind in (0, 1):
        ts = [mk(M, N, K, kind) for _ in range(3)]
        ds = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(3)]
        for rnd in range(2):
            for (cg, bn) in ((2, 256), (4, 256), (2, 192), (4, 192)):
                for flags in (0, 1):
                    os.environ["B200Q_GEMM_DEBUG_FLAGS"] = str(flags)
                    for i in range(3): gemm(ts[i], M, N, K, kind, cg, bn, ds[i])
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(30): gemm(ts[i % 3], M, N, K, kind, cg, bn, ds[i % 3])
                    e1.record(); torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) / 30 * 1e3
                    print(json.dumps(dict(check="mc_time", kind=kind, cg=cg, bn=bn, flags=flags, rnd=rnd, us=round(us, 1), tflops=round(2.0 * M * N * K / us / 1e6))), flush=True)
        os.environ["B200Q_GEMM_DEBUG_FLAGS"] = "0"

if __name__ == "__main__":
    which = sys.argv[1:] or ["parity", "time"]
    if "parity" in which: parity()
    if "time" in which: timing()
