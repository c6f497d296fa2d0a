#!/usr/bin/env python
"""A handful of decode-size GEMM launches (M = 16, N = 14336, K = 4096, rotating weights) for an ncu capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import qutlass_b200 as Q
dev = torch.device("cuda")
M, N, K = 16, 14336, 4096
al = torch.ones(1, device=dev)
a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
sfa = torch.randint(126, 129, (128 * K // 32,), dtype=torch.uint8, device=dev).view(torch.float8_e8m0fnu)
ws = [(torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev), torch.randint(126, 129, (N * K // 32,), dtype=torch.uint8, device=dev).view(torch.float8_e8m0fnu)) for _ in range(6)]
for i in range(8):
    w, sfw = ws[i % 6]
    Q.matmul_mxf4_bf16_tn(a, w, sfa, sfw, al, static_weights=True)
torch.cuda.synchronize()
