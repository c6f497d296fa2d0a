#!/usr/bin/env python
"""MXFP8 GEMM at the config-1 shape: every configuration incl. the cluster-of-4 multicast one (bit equality + time)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref_gpu
HAVE_REF = ref_gpu.available() and os.environ.get("F8_PROBE_REF", "1") == "1"
if HAVE_REF:
    torch.ops.load_library(ref_gpu.LIB)      # the compiled reference, same process, same buffers (tools/ref_msweep.py does the same)
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
N, K = 14336, 4096
for M in tuple(int(v) for v in os.environ.get('F8_PROBE_M', '1024,4096,16384').split(',')):
    sets = []
    for i in range(3):
        a = torch.randint(0, 120, (M, K), dtype=torch.uint8, device=dev); b = torch.randint(0, 120, (N, K), dtype=torch.uint8, device=dev)
        sfa = torch.randint(126, 129, (((M + 127) // 128) * 128 * (K // 32),), dtype=torch.uint8, device=dev)
        sfb = torch.randint(126, 129, (N * (K // 32),), dtype=torch.uint8, device=dev)
        d = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        sets.append((a, b, sfa, sfb, d))
    alpha = torch.ones(1, device=dev); st = torch.cuda.current_stream().cuda_stream
    ref = None
    for (cg, bn) in ((0, 0), (2, 256), (2, 192), (2, 128), (1, 256), (1, 128)):
        def go(i):
            a, b, sfa, sfb, d = sets[i % 3]
            rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(), d.data_ptr(), M, N, K, 2, cg, bn, st)
            assert rc == 0, lib.b200q_last_error()
        for i in range(3): go(i)
        torch.cuda.synchronize()
        out = sets[0][4].clone()
        if ref is None: ref = out
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(12): go(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 12 * 1e3
        print(json.dumps(dict(M=M, cg=cg, bn=bn, us=round(us, 1), tflops=round(2.0 * M * N * K / us / 1e6), equal=bool(torch.equal(out, ref)))), flush=True)
    if HAVE_REF:
        op = torch.ops._qutlass_C.matmul_mxf8_bf16_tn
        def go_ref(i):
            a, b, sfa, sfb, d = sets[i % 3]
            return op(a.view(torch.float8_e4m3fn), b.view(torch.float8_e4m3fn), sfa.view(torch.float8_e8m0fnu), sfb.view(torch.float8_e8m0fnu), alpha)
        for i in range(3): out = go_ref(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(12): go_ref(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 12 * 1e3
        print(json.dumps(dict(M=M, impl="reference matmul_mxf8_bf16_tn (allocates its output)", us=round(us, 1), tflops=round(2.0 * M * N * K / us / 1e6),
                              equal=bool(torch.equal(go_ref(0), ref)))), flush=True)
