#!/usr/bin/env python
"""BASELINE configs[3]: fusedQuantizeMx Quartet ("quest") + abs_max with Hadamard-32/64/128, M sweep {1,16,128,1024,4096,16384},
K = 4096 -- ours vs the compiled reference's kernels in ONE process (reference library loaded first, like tools/ref_msweep.py),
CUDA-graph replay over rotating input/output sets (footprint > L2 for the large M), median of 5.
Columns: us, achieved GB/s on the algorithmic bytes (2 B in + 0.5 B codes + 1/32 B scale per element), fraction of the measured
HBM copy bandwidth (MEASURED_PEAKS.json).  -> gpurun_out/quant_sweep.{jsonl,md}"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref_gpu
HAVE_REF = ref_gpu.available()
if HAVE_REF:
    torch.ops.load_library(ref_gpu.LIB)
    ops = torch.ops._qutlass_C
import qutlass_b200 as Q   # noqa: E402
dev = torch.device("cuda")
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
K = 4096


def hadamard(h):
    idx = torch.arange(h)
    bits = idx[:, None] & idx[None, :]
    par = torch.zeros_like(bits)
    while bits.any():
        par ^= bits & 1
        bits = bits >> 1
    return ((1.0 - 2.0 * par.double()) * h ** -0.5).to(torch.bfloat16).to(dev)


def graph_time(fn, iters, reps=5):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn(0); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for i in range(iters): fn(i)
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / iters * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


rows = []
MS = tuple(int(v) for v in os.environ.get("QUANT_SWEEP_M", "1,16,128,1024,4096,16384").split(","))
OUT = os.environ.get("QUANT_SWEEP_OUT", "quant_sweep")
for M in MS:
    nsets = 2 if M >= 4096 else 4
    nsets = max(nsets, min(8, int(200e6 / (M * K * 2.6)) + 1)) if M >= 1024 else 4      # rotate through > L2 when it matters
    xs = [torch.randn(M, K, dtype=torch.bfloat16, device=dev) for _ in range(nsets)]
    for had in (32, 64, 128):
        R = hadamard(had)
        for method in ("quest", "abs_max"):
            Q.fusedQuantizeMx(xs[0], R, method=method)
            iters = 24 if M <= 1024 else 12
            ours = graph_time(lambda i: Q.fusedQuantizeMx(xs[i % nsets], R, method=method), iters)
            rec = dict(M=M, K=K, had=had, method=method, ours_us=round(ours, 2))
            if HAVE_REF:
                pr, pc = (M + 127) // 128 * 128, K // 32
                outs = [(torch.empty(M, K // 2, dtype=torch.uint8, device=dev), torch.empty(pr, pc, dtype=torch.float8_e8m0fnu, device=dev)) for _ in range(nsets)]
                op = ops.fusedQuantizeMxQuest if method == "quest" else ops.fusedQuantizeMxAbsMax
                rec["ref_us"] = round(graph_time(lambda i: op(xs[i % nsets], R, outs[i % nsets][0], outs[i % nsets][1]), iters), 2)
            byts = M * K * (2 + 0.5 + 1.0 / 32)
            rec["ours_gbs"] = round(byts / ours / 1e3, 1)
            rec["ours_frac_hbm"] = round(byts / ours / 1e3 / HBM, 3)
            rows.append(rec)
            print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", OUT + ".md"), "w") as f:
    f.write(f"# configs[3]: fusedQuantizeMx, K = 4096, graph replay over rotating sets, median of 5; HBM peak {HBM} GB/s (measured copy)\n\n")
    f.write("| M | Hadamard | method | ours us | reference us | ours GB/s | of measured HBM |\n|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write(f"| {r['M']} | {r['had']} | {r['method']} | {r['ours_us']} | {r.get('ref_us', '-')} | {r['ours_gbs']} | {r['ours_frac_hbm']} |\n")
