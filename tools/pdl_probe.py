#!/usr/bin/env python
"""A/B inside ONE process: does launching the quantise kernels as programmatic dependents (griddepcontrol.wait after
their prologue) shorten the chains they sit in?  B200Q_NO_PDL is read at every launch, so the variants interleave:
  ''  = PDL on the quantisers and the GEMM (default)      '2' = GEMM only (the behaviour before)      '1' = off everywhere
Chains: the config-1 step (quantise + GEMM, M=4096), quantise alone, and the decode step (M=16).  One JSON line each."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import qutlass_b200 as Q
from qutlass_b200 import _lib

N, K, HAD = 14336, 4096, 128
dev = torch.device("cuda")
lib = _lib.load()
idx = torch.arange(HAD)
bits = idx[:, None] & idx[None, :]
par = torch.zeros_like(bits)
while bits.any():
    par ^= bits & 1
    bits = bits >> 1
H = ((1.0 - 2.0 * par.double()) * HAD ** -0.5).to(torch.bfloat16).to(dev)
alpha = torch.ones(1, device=dev)
NS = 4
w = torch.randn(N, K, dtype=torch.bfloat16, device=dev)
q_, s_ = Q.fusedQuantizeMx(w, H, method="abs_max")
wqs = [q_.clone() for _ in range(NS)]
wsfs = [Q.to_blocked(s_).clone() for _ in range(NS)]
del w
st = torch.cuda.current_stream().cuda_stream
meth = Q.METHOD_ABSMAX | Q.ROT_TRUSTED_HADAMARD


def make(M):
    acts = [torch.randn(M, K, dtype=torch.bfloat16, device=dev) for _ in range(NS)]
    aqs = [torch.empty(M, K // 2, dtype=torch.uint8, device=dev) for _ in range(NS)]
    asfs = [torch.empty(((M + 127) // 128) * 128 * (K // 32), dtype=torch.uint8, device=dev) for _ in range(NS)]
    outs = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(NS)]

    def quant(i):
        s = i % NS
        _lib.check(lib.b200q_quantize_mx(acts[s].data_ptr(), H.data_ptr(), aqs[s].data_ptr(), None, asfs[s].data_ptr(), None,
                                         M * K, K, HAD, meth, st))

    def gemm(i):
        s = i % NS
        _lib.check(lib.b200q_gemm_fp4(aqs[s].data_ptr(), wqs[s].data_ptr(), asfs[s].data_ptr(), wsfs[s].data_ptr(),
                                      alpha.data_ptr(), outs[s].data_ptr(), M, N, K, 0, st))

    def step(i):
        quant(i)
        gemm(i)
    return quant, step


def timed(fn, n, warm=10):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


q4096, s4096 = make(4096)
q16, s16 = make(16)
CHAINS = [("step M=4096 (quantise + GEMM)", s4096, 200), ("quantise alone M=4096", q4096, 200), ("step M=16", s16, 400),
          ("quantise alone M=16", q16, 400)]
res = {}
for rnd in range(int(os.environ.get("PROBE_ROUNDS", 4))):
    for mode in ("", "2", "1"):
        if mode:
            os.environ["B200Q_NO_PDL"] = mode
        else:
            os.environ.pop("B200Q_NO_PDL", None)
        for name, fn, n in CHAINS:
            res.setdefault((name, mode), []).append(timed(fn, n))
os.environ.pop("B200Q_NO_PDL", None)
for (name, mode), ts in res.items():
    ts = sorted(ts)
    print(json.dumps({"chain": name, "B200Q_NO_PDL": mode or "unset", "min_us": round(ts[0], 2),
                      "median_us": round(ts[len(ts) // 2], 2), "max_us": round(ts[-1], 2)}), flush=True)
