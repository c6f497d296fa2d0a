#!/usr/bin/env python
"""Where does the fused quantise+GEMM kernel spend its time?  A/B timings inside ONE process (config 1, MXFP4):
two launches, GEMM alone (both tile orders), the fused kernel with the quantisers switched off, and the full fused kernel
with 2 / 4 quantiser warps.  Prints one JSON line per variant."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200Q_FUSE"] = "1"
import torch
import qutlass_b200 as Q
from qutlass_b200 import _lib

M, N, K, HAD = int(os.environ.get("PROBE_M", 4096)), 14336, 4096, 128
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
acts = [torch.randn(M, K, dtype=torch.bfloat16, device=dev) for _ in range(NS)]
aqs = [torch.empty(M, K // 2, dtype=torch.uint8, device=dev) for _ in range(NS)]
asfs = [torch.empty(((M + 127) // 128) * 128 * (K // 32), dtype=torch.uint8, device=dev) for _ in range(NS)]
outs = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(NS)]
ws = torch.zeros(4096, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
meth = Q.METHOD_ABSMAX | Q.ROT_TRUSTED_HADAMARD

def quant(i):
    s = i % NS
    _lib.check(lib.b200q_quantize_mx(acts[s].data_ptr(), H.data_ptr(), aqs[s].data_ptr(), None, asfs[s].data_ptr(), None, M * K, K, HAD, meth, st))
def gemm(i):
    s = i % NS
    _lib.check(lib.b200q_gemm_fp4(aqs[s].data_ptr(), wqs[s].data_ptr(), asfs[s].data_ptr(), wsfs[s].data_ptr(), alpha.data_ptr(), outs[s].data_ptr(), M, N, K, 0, st))
def two(i):
    quant(i); gemm(i)
def fused(i):
    s = i % NS
    _lib.check(lib.b200q_linear_fp4(acts[s].data_ptr(), H.data_ptr(), aqs[s].data_ptr(), None, asfs[s].data_ptr(), wqs[s].data_ptr(), wsfs[s].data_ptr(),
                                    alpha.data_ptr(), None, outs[s].data_ptr(), ws.data_ptr(), M, N, K, HAD, meth, 0, st))

def timed(fn, n=100, warm=10):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

VARIANTS = [
    ("two launches (quantise, GEMM)", two, 0, None),
    ("GEMM alone, M-fastest tiles", gemm, 0, None),
    ("GEMM alone, N-fastest tiles", gemm, 2048, None),
    ("fused, quantisers off, 4 warps (chunked drain, 448 thr)", fused, 1024, 4),
    ("fused, quantisers off, 2 warps (plain epilogue, 384 thr)", fused, 1024, 2),
    ("fused, 4 quantiser warps + epilogue helpers", fused, 0, 4),
    ("fused, 4 quantiser warps + helpers, producer does not wait (timing only)", fused, 8192, 4),
    ("fused, 4 quantiser warps, no helpers", fused, 4096, 4),
    ("fused, 4 quantiser warps, no helpers, no waits (timing only)", fused, 4096 | 8192, 4),
    ("fused, 4 quantiser warps, no helpers, N-fastest", fused, 4096 | 2048, 4),
    ("fused, 4 quantiser warps, no helpers, N-fastest, no waits (timing only)", fused, 4096 | 2048 | 8192, 4),
    ("fused, 2 quantiser warps, no helpers, N-fastest", fused, 4096 | 2048, 2),
    ("fused, 4 quantiser warps + helpers, N-fastest", fused, 2048, 4),
]
for i in range(NS): quant(i)
res = {v[0]: [] for v in VARIANTS}
for rnd in range(int(os.environ.get("PROBE_ROUNDS", 6))):       # interleaved: clock / thermal drift hits every variant alike
    for name, fn, flags, warps in VARIANTS:
        os.environ["B200Q_GEMM_DEBUG_FLAGS"] = str(flags)
        if warps: os.environ["B200Q_FUSE_WARPS"] = str(warps)
        res[name].append(timed(fn, n=60, warm=5))
        ws.zero_(); torch.cuda.synchronize()
for name, ts in res.items():
    ts = sorted(ts)
    print(json.dumps({"variant": name, "min_us": round(ts[0], 2), "median_us": round(ts[len(ts) // 2], 2), "max_us": round(ts[-1], 2)}), flush=True)
