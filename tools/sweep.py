#!/usr/bin/env python
"""One-process timing sweep (GEMM configs x debug flags x K, quantise variants) -> gpurun_out/sweep2.jsonl"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qutlass_b200 as Q
from qutlass_b200 import _lib
import oracle as O

lib = _lib.load()
dev = torch.device("cuda")
out_path = os.path.join(ROOT, "gpurun_out", "sweep2.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
fout = open(out_path, "a")

def emit(**kw):
    fout.write(json.dumps(kw) + "\n"); fout.flush()
    print(json.dumps(kw), flush=True)

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

def gemm_sweep(kind, shapes, cfgs, variants):
    knd = 0 if kind == "mx" else 1
    group = 32 if kind == "mx" else 16
    for (M, N, K) in shapes:
        a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
        b = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev)
        lo, hi = (126, 129) if kind == "mx" else (0x30, 0x41)
        nsa = ((M + 127) // 128) * 128 * (((K // group) + 3) // 4) * 4
        nsb = ((N + 127) // 128) * 128 * (((K // group) + 3) // 4) * 4
        sfa = torch.randint(lo, hi, (nsa,), dtype=torch.uint8, device=dev)
        sfb = torch.randint(lo, hi, (nsb,), dtype=torch.uint8, device=dev)
        d = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        alpha = torch.ones(1, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for (cg, bn) in cfgs:
            for (flags, sfb_n) in variants:
                os.environ["B200Q_GEMM_DEBUG_FLAGS"] = str(flags)
                os.environ["B200Q_SF_BUFS"] = str(sfb_n)
                def run():
                    rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(),
                                                d.data_ptr(), M, N, K, knd, cg, bn, st)
                    if rc: raise RuntimeError(lib.b200q_last_error().decode())
                try:
                    us = timeit(run)
                    emit(check="gemm_time", kind=kind, M=M, N=N, K=K, cg=cg, bn=bn, flags=flags, sf_bufs=sfb_n, us=round(us, 1),
                         tflops=round(2.0 * M * N * K / us / 1e6, 0))
                except Exception as e:
                    emit(check="gemm_time", kind=kind, M=M, N=N, K=K, cg=cg, bn=bn, flags=flags, sf_bufs=sfb_n, error=str(e)[:200])
        os.environ["B200Q_GEMM_DEBUG_FLAGS"] = "0"; os.environ["B200Q_SF_BUFS"] = "2"

def quant_sweep():
    for (M, K) in ((4096, 4096), (16384, 4096), (128, 4096), (1, 4096)):
        x = torch.randn(M, K, dtype=torch.bfloat16, device=dev)
        for kind in ("mx", "nv"):
            for had in ((32, 64, 128) if kind == "mx" else (16, 128)):
                H = torch.from_numpy(O.bf16_bits(O.hadamard_matrix(had)).astype(np.int16)).view(torch.bfloat16).to(dev)
                gs = torch.ones(1, device=dev)
                for method in ("abs_max", "quest"):
                    if kind == "mx":
                        fn = lambda: Q.fusedQuantizeMx(x, H, method=method)
                    else:
                        fn = lambda: Q.fusedQuantizeNv(x, H, gs, method=method)
                    us = timeit(fn)
                    byts = M * K * (2.5 + (2.0 / 32 if kind == "mx" else 2.0 / 16))
                    emit(check="quant_time", kind=kind, M=M, K=K, had=had, method=method, us=round(us, 1), gbs=round(byts / us / 1e3, 0))

if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "quant"]
    if "quant" in which:
        quant_sweep()
    if "gemmst" in which:
        F = lambda base, n: base | (n << 12)
        gemm_sweep("mx", [(4096, 14336, 4096)], [(2, 256)], [(0, 2), (F(256, 4), 2), (F(256, 10), 2), (F(512, 6), 2), (F(512, 14), 2), (F(768, 6), 2)])
        gemm_sweep("mx", [(4096, 14336, 4096)], [(2, 192)], [(0, 2), (F(512, 6), 2), (F(768, 6), 2)])
    if "gemml2" in which:
        # is the mainloop bound by operand traffic (L2 -> SM / shared-memory writes)?  1 = no D stores, 1<<20 = no B loads, 1<<21 = no A loads
        B_, A_ = 1 << 20, 1 << 21
        gemm_sweep("mx", [(4096, 14336, 4096)], [(2, 256), (2, 192), (1, 256)], [(0, 2), (1, 2), (B_, 2), (B_ | 1, 2), (A_, 2), (A_ | 1, 2), (A_ | B_ | 1, 2)])
    if "gemmk" in which:
        # per-tile overhead vs per-k-tile cost: time = rounds x (k_tiles x t_ktile + t_tile); fit over K
        B_, A_ = 1 << 20, 1 << 21
        gemm_sweep("mx", [(4096, 14336, 2048), (4096, 14336, 4096), (4096, 14336, 8192), (4096, 14336, 16384)], [(2, 256), (2, 192)],
                   [(0, 2), (1, 2), (A_ | B_ | 1, 2), (A_ | B_ | 3, 2)])
    if "gemmissue" in which:
        # is the MMA-issuing warp the limiter at BN = 192?  32 = skip the scale copies (tcgen05.cp), timing only
        B_, A_ = 1 << 20, 1 << 21
        gemm_sweep("mx", [(4096, 14336, 4096)], [(2, 192), (2, 256), (2, 128)], [(A_ | B_ | 1, 2), (A_ | B_ | 1 | 32, 2), (1, 2), (1 | 32, 2), (0, 2), (32, 2)])
    if "gemmq" in which:
        gemm_sweep("mx", [(4096, 14336, 4096)], [(2, 192), (2, 256)], [(0, 2), (1, 2), (4, 2)])
        gemm_sweep("nv", [(4096, 14336, 4096)], [(2, 192), (2, 256)], [(0, 2), (4, 2)])
    if "quantk" in which:
        # kernel-only quantise timing through the C-ABI (no torch allocations in the loop)
        for (M, K) in ((4096, 4096), (16384, 4096)):
            x = torch.randn(M, K, dtype=torch.bfloat16, device=dev)
            qo = torch.empty(M, K // 2, dtype=torch.uint8, device=dev)
            for kind, group in (("mx", 32), ("nv", 16)):
                sfo = torch.empty(M * K // group + 65536, dtype=torch.uint8, device=dev)
                sfb = torch.empty(M * K // group + 65536, dtype=torch.uint8, device=dev)
                gs = torch.ones(1, device=dev)
                for had in (32, 64, 128):
                    H = torch.from_numpy(O.bf16_bits(O.hadamard_matrix(had)).astype(np.int16)).view(torch.bfloat16).to(dev)
                    for bf in ("1", "T"):
                        os.environ["B200Q_QUANT_MMA"] = "0"
                        TRUST = 0x100 if bf == "T" else 0
                        st = torch.cuda.current_stream().cuda_stream
                        if kind == "mx":
                            fn = lambda: lib.b200q_quantize_mx(x.data_ptr(), H.data_ptr(), qo.data_ptr(), sfo.data_ptr(), sfb.data_ptr(), None, M * K, K, had, 1 | TRUST, st)
                        else:
                            fn = lambda: lib.b200q_quantize_nv(x.data_ptr(), H.data_ptr(), qo.data_ptr(), sfo.data_ptr(), sfb.data_ptr(), gs.data_ptr(), M * K, K, had, 1 | TRUST, st)
                        us = timeit(fn, iters=50)
                        byts = M * K * (2.5 + 2.0 / group)
                        emit(check="quantk", kind=kind, M=M, K=K, had=had, butterfly=bf, us=round(us, 1), gbs=round(byts / us / 1e3, 0))
        os.environ["B200Q_QUANT_MMA"] = "0"
    if "gemm" in which:
        cfgs = [(1, 128), (1, 256), (2, 128), (2, 192), (2, 256)]
        gemm_sweep("mx", [(4096, 14336, 4096)], cfgs, [(0, 2), (1, 2), (3, 2), (4, 2)])
        gemm_sweep("mx", [(4096, 14336, 16384)], [(1, 256), (2, 192), (2, 256)], [(0, 2), (3, 2)])
        gemm_sweep("nv", [(4096, 14336, 4096)], [(2, 192), (2, 256)], [(0, 2)])
        gemm_sweep("mx", [(16, 14336, 4096), (128, 14336, 4096), (1024, 14336, 4096), (16384, 14336, 4096)], [(1, 64), (1, 128), (2, 256)], [(0, 2)])
