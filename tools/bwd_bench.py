#!/usr/bin/env python
"""Kernel-only timing of the four backward re-quantisers through the C-ABI, CUDA-graph
replayed, against their algorithmic bytes.  Rotating buffer sets larger than L2 unless
--warm.  Writes gpurun_out/bwd_bench.jsonl and prints the lines."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
import qutlass_b200 as Q

lib = _lib.load(); dev = torch.device("cuda")
HBM_PEAK = 6545.9
try:
    HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def graph_time(fns, iters):
    """fns: list of closures (one per rotating buffer set) taking a raw stream"""
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        for f in fns: f(s.cuda_stream)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for i in range(iters): fns[i % len(fns)](s.cuda_stream)
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


def had32():
    h = torch.tensor([[1.0]])
    while h.size(0) < 32:
        h = torch.cat([torch.cat([h, h], 1), torch.cat([h, -h], 1)], 0)
    return (h * 32 ** -0.5).to(torch.bfloat16).to(dev)


def ck(rc):
    assert rc == 0, lib.b200q_last_error()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4096x4096,16384x4096,4096x14336")
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    H = had32(); al = torch.tensor([3.0], device=dev)
    out = []
    for shp in args.shapes.split(","):
        n, m = (int(v) for v in shp.split("x"))          # input [n, m]
        sets = max(1, min(args.sets, int(1.2e9 // (n * m * 3.1)) or 1))
        E = n * m
        # backward_t_bf16
        xs = [torch.randn(n, m, dtype=torch.bfloat16, device=dev) * 25 for _ in range(sets)]
        q = [torch.empty(m, n // 2, dtype=torch.uint8, device=dev) for _ in range(sets)]
        sf = [torch.empty(m, n // 32, dtype=torch.uint8, device=dev) for _ in range(sets)]
        flags = Q.ROT_TRUSTED_HADAMARD
        fns = [(lambda st, i=i: ck(lib.b200q_backward_t_bf16(xs[i].data_ptr(), H.data_ptr(), q[i].data_ptr(), sf[i].data_ptr(), m, n, 1, flags, st))) for i in range(sets)]
        t = graph_time(fns, args.iters); by = E * (2 + 0.5 + 1 / 32)
        out.append(dict(kernel="backward_t_bf16", n=n, m=m, us=round(t, 2), gbs=round(by / t / 1e3, 1), frac=round(by / t / 1e3 / HBM_PEAK, 3)))
        fns = [(lambda st, i=i: ck(lib.b200q_backward_t_bf16(xs[i].data_ptr(), H.data_ptr(), q[i].data_ptr(), sf[i].data_ptr(), m, n, 1, 0, st))) for i in range(sets)]
        t = graph_time(fns, args.iters)
        out.append(dict(kernel="backward_t_bf16(generic rot)", n=n, m=m, us=round(t, 2), gbs=round(by / t / 1e3, 1), frac=round(by / t / 1e3 / HBM_PEAK, 3)))
        # forward quantiser on a same-sized tensor for comparison (H = 32, abs_max, trusted)
        a = [torch.empty(n, m // 2, dtype=torch.uint8, device=dev) for _ in range(sets)]
        sfa = [torch.zeros(((n + 127) // 128) * 128 * (m // 32), dtype=torch.uint8, device=dev) for _ in range(sets)]
        fns = [(lambda st, i=i: ck(lib.b200q_quantize_mx(xs[i].data_ptr(), H.data_ptr(), a[i].data_ptr(), None, sfa[i].data_ptr(), None, n * m, m, 32, 1 | 0x100, st))) for i in range(sets)]
        t = graph_time(fns, args.iters)
        out.append(dict(kernel="fusedQuantizeMx(H=32) [comparison]", n=n, m=m, us=round(t, 2), gbs=round(by / t / 1e3, 1), frac=round(by / t / 1e3 / HBM_PEAK, 3)))
        # square double
        y = [torch.empty(n, m, dtype=torch.uint8, device=dev) for _ in range(sets)]
        rs = [torch.empty(n, m // 32, dtype=torch.uint8, device=dev) for _ in range(sets)]
        cs = [torch.empty(m, n // 32, dtype=torch.uint8, device=dev) for _ in range(sets)]
        fns = [(lambda st, i=i: ck(lib.b200q_backward_bf16_square_double_mxfp8(xs[i].data_ptr(), n, m, y[i].data_ptr(), rs[i].data_ptr(), cs[i].data_ptr(), st))) for i in range(sets)]
        t = graph_time(fns, args.iters); by = E * (2 + 1 + 2 / 32)
        out.append(dict(kernel="backward_bf16_square_double_mxfp8", n=n, m=m, us=round(t, 2), gbs=round(by / t / 1e3, 1), frac=round(by / t / 1e3 / HBM_PEAK, 3)))
        del xs, y
        # backward_qt / mxfp4_transpose_mxfp8
        xq = [torch.randint(0, 256, (n, m // 2), dtype=torch.uint8, device=dev) for _ in range(sets)]
        xsf = [torch.randint(120, 134, (n, m // 32), dtype=torch.uint8, device=dev) for _ in range(sets)]
        fns = [(lambda st, i=i: ck(lib.b200q_backward_qt_bf16(xq[i].data_ptr(), xsf[i].data_ptr(), H.data_ptr(), al.data_ptr(), q[i].data_ptr(), sf[i].data_ptr(), m, n, 1, flags, st))) for i in range(sets)]
        t = graph_time(fns, args.iters); by = E * (0.5 + 1 / 32) * 2
        out.append(dict(kernel="backward_qt_bf16", n=n, m=m, us=round(t, 2), gbs=round(by / t / 1e3, 1), frac=round(by / t / 1e3 / HBM_PEAK, 3)))
        npad = (n + 255) // 256 * 256
        y8 = [torch.empty(m, npad, dtype=torch.uint8, device=dev) for _ in range(sets)]
        e8 = [torch.empty(m, npad // 32, dtype=torch.uint8, device=dev) for _ in range(sets)]
        fns = [(lambda st, i=i: ck(lib.b200q_mxfp4_transpose_mxfp8(xq[i].data_ptr(), xsf[i].data_ptr(), n, m, y8[i].data_ptr(), e8[i].data_ptr(), st))) for i in range(sets)]
        t = graph_time(fns, args.iters); by = E * (0.5 + 1 / 32 + 1 + 1 / 32)
        out.append(dict(kernel="mxfp4_transpose_mxfp8", n=n, m=m, us=round(t, 2), gbs=round(by / t / 1e3, 1), frac=round(by / t / 1e3 / HBM_PEAK, 3)))
        del xq, y8
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    form = "pipelined" if os.environ.get("B200Q_BWD_PIPE") == "1" else ("oneshot" if os.environ.get("B200Q_BWD_PIPE") == "0" else "default")
    with open(os.path.join(ROOT, "gpurun_out", "bwd_bench.jsonl"), "w") as f:
        for r in out:
            r["form"] = form
            f.write(json.dumps(r) + "\n"); print(json.dumps(r))


if __name__ == "__main__":
    main()
