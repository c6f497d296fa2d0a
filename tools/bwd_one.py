#!/usr/bin/env python
"""A few launches of the tensor-core backward kernels at 16384 x 4096 (for `ncu -k regex:bwd_t_tc_kernel|bwd_qt_tc_kernel`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
import qutlass_b200 as Q
lib = _lib.load(); dev = torch.device("cuda")
n, m = 16384, 4096
h = torch.tensor([[1.0]])
while h.size(0) < 32:
    h = torch.cat([torch.cat([h, h], 1), torch.cat([h, -h], 1)], 0)
H = (h * 32 ** -0.5).to(torch.bfloat16).to(dev)
xs = [torch.randn(n, m, dtype=torch.bfloat16, device=dev) * 25 for _ in range(3)]
q = torch.empty(m, n // 2, dtype=torch.uint8, device=dev); sf = torch.empty(m, n // 32, dtype=torch.uint8, device=dev)
xq = [torch.randint(0, 256, (n, m // 2), dtype=torch.uint8, device=dev) for _ in range(3)]
xsf = [torch.randint(120, 134, (n, m // 32), dtype=torch.uint8, device=dev) for _ in range(3)]
al = torch.tensor([3.0], device=dev)
st = torch.cuda.current_stream().cuda_stream
for i in range(6):
    assert lib.b200q_backward_t_bf16(xs[i % 3].data_ptr(), H.data_ptr(), q.data_ptr(), sf.data_ptr(), m, n, 1, Q.ROT_TRUSTED_HADAMARD, st) == 0
    assert lib.b200q_backward_qt_bf16(xq[i % 3].data_ptr(), xsf[i % 3].data_ptr(), H.data_ptr(), al.data_ptr(), q.data_ptr(), sf.data_ptr(), m, n, 1, Q.ROT_TRUSTED_HADAMARD, st) == 0
torch.cuda.synchronize()
print("ok")
