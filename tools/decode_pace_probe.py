#!/usr/bin/env python
"""Decode kernel (gemm_decode.cu): paced weight streaming (B200Q_DECODE_PACE = SM cycles between the stage requests of a CTA)
vs everything-at-once, graph replay of 24 launches, weights L2-resident ("same") or streamed from HBM (8 rotating sets of
29 MB, "rot"), safe / static-weights launches; every paced result is compared bit-for-bit with the unpaced one.
-> JSON lines (product library)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
N, K = int(os.environ.get("PROBE_N", 14336)), int(os.environ.get("PROBE_K", 4096))
KIND = int(os.environ.get("PROBE_KIND", 0))
NSETS = 8
PACES = [int(v) for v in os.environ.get("PROBE_PACES", "0,300,380,460,540,620,720").split(",")]
alpha = torch.ones(1, device=dev)
g = 16 if KIND == 1 else 32
wsets = [(torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev),
          torch.randint(118 if KIND == 1 else 126, 122 if KIND == 1 else 129, (N * K // g,), dtype=torch.uint8, device=dev)) for _ in range(NSETS)]


def set_pace(v):
    os.environ["B200Q_DECODE_PACE"] = str(v)
    lib.b200q_reload_env()


def graph_time(fn, iters, reps=7):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn(0); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=s):
            for i in range(iters): fn(i)
    torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / iters * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 2)


for M in (1, 16, 32):
    a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
    sfa = torch.randint(118 if KIND == 1 else 126, 122 if KIND == 1 else 129, (128 * K // g,), dtype=torch.uint8, device=dev)
    d = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(NSETS)]

    def go(i, rotate, kindflag):
        w, sfw = wsets[i % NSETS if rotate else 0]
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), w.data_ptr(), sfa.data_ptr(), sfw.data_ptr(), alpha.data_ptr(),
                                    d[i % NSETS].data_ptr(), M, N, K, KIND | kindflag, 1, 16, st)
        assert rc == 0, lib.b200q_last_error()

    ref = None
    for pace in PACES:
        set_pace(pace)
        rec = dict(M=M, N=N, K=K, kind=KIND, pace=pace)
        for i in range(NSETS): go(i, True, 0)
        torch.cuda.synchronize()
        outs = torch.stack([t.clone() for t in d]).view(torch.int16)
        if ref is None: ref = outs
        rec["equal_to_unpaced"] = bool(torch.equal(outs, ref))
        for rotate in (True, False):
            for kf in (0, 0x100):
                rec[f"{'rot' if rotate else 'same'}_{'static' if kf else 'safe'}_us"] = graph_time(lambda i: go(i, rotate, kf), 24)
        print(json.dumps(rec), flush=True)
set_pace(0)
