#!/usr/bin/env python
"""One quantise launch per variant (for ncu): python tools/quant_one.py M K had kind tc"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
M, K, h = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]); kind = sys.argv[4]; os.environ["B200Q_QUANT_TC"] = sys.argv[5]
import torch
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
hm = torch.tensor([[1.0]])
while hm.size(0) < h: hm = torch.cat([torch.cat([hm, hm], 1), torch.cat([hm, -hm], 1)], 0)
H = (hm * h ** -0.5).to(torch.bfloat16).to(dev)
group = 32 if kind == "mx" else 16
x = torch.randn(M, K, dtype=torch.bfloat16, device=dev) * 25
q = torch.empty(M, K // 2, dtype=torch.uint8, device=dev)
nsf = ((M + 127) // 128) * 128 * (K // group)
sf = torch.empty(nsf, dtype=torch.uint8, device=dev); sfb = torch.empty(nsf, dtype=torch.uint8, device=dev); gs = torch.ones(1, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    if kind == "mx": rc = lib.b200q_quantize_mx(x.data_ptr(), H.data_ptr(), q.data_ptr(), sf.data_ptr(), sfb.data_ptr(), None, M * K, K, h, 1 | 0x100, st)
    else: rc = lib.b200q_quantize_nv(x.data_ptr(), H.data_ptr(), q.data_ptr(), sf.data_ptr(), sfb.data_ptr(), gs.data_ptr(), M * K, K, h, 1 | 0x100, st)
    assert rc == 0, lib.b200q_last_error()
torch.cuda.synchronize()
