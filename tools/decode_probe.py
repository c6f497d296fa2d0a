#!/usr/bin/env python
"""Where do the 7 us of a decode-size GEMM (M <= 128, N = 14336, K = 4096) go?   B200Q_LIB=prof python tools/decode_probe.py
(a) graph-replay time with the SAME buffers every iteration (weights L2-resident after the first) vs 8 rotating weight sets
    (250 MB > L2: every iteration streams from HBM) -- tells HBM-bound from latency-bound;
(b) timeline of CTA 0 of one launch (profiling flag 1 << 24): cycles from kernel entry to set-up done, grid dependency
    resolved, first k-tile landed, last MMA issued, accumulator complete, drained, stores issued, exit; ns entry -> exit.
Needs the profiling build (python -m qutlass_b200.build --profiling)."""
import ctypes, json, os, sys
os.environ.setdefault("B200Q_LIB", "prof")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
raw = ctypes.CDLL(_lib.LIB_PATH)
raw.b200q_debug_read_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
raw.b200q_debug_read_ktrace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
assert lib.b200q_profiling_build() == 1, "needs lib/libb200q_prof.so"
N, K = int(os.environ.get("PROBE_N", 14336)), int(os.environ.get("PROBE_K", 4096))
NSETS = 8
alpha = torch.ones(1, device=dev)
wsets = [(torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev),
          torch.randint(126, 129, (N * K // 32,), dtype=torch.uint8, device=dev)) for _ in range(NSETS)]


def setenv(flags):
    os.environ["B200Q_GEMM_DEBUG_FLAGS"] = str(flags)
    lib.b200q_reload_env()


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
    return ts[0], ts[len(ts) // 2]


for M in (1, 16, 128):
    a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
    sfa = torch.randint(126, 129, (128 * K // 32,), dtype=torch.uint8, device=dev)
    d = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(NSETS)]

    def go(i, rotate, kindflag, cg=0, bn=0):
        w, sfw = wsets[i % NSETS if rotate else 0]
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), w.data_ptr(), sfa.data_ptr(), sfw.data_ptr(), alpha.data_ptr(),
                                    d[i % NSETS if rotate else 0].data_ptr(), M, N, K, kindflag, cg, bn, st)
        assert rc == 0, lib.b200q_last_error()

    setenv(0)
    rec = dict(M=M, N=N, K=K)
    for name, rotate, kf in (("same_buffers", False, 0), ("rotating_weights", True, 0), ("same_buffers_static", False, 0x100),
                             ("rotating_static", True, 0x100)):
        best, med = graph_time(lambda i: go(i, rotate, kf), 24)
        rec[name + "_us"] = [round(best, 2), round(med, 2)]
    for cg, bn in ((1, 64), (1, 128), (1, 192), (1, 256)):
        try:
            best, med = graph_time(lambda i: go(i, True, 0x100, cg, bn), 24)
            rec[f"rotating_static_cfg{cg}x{bn}_us"] = [round(best, 2), round(med, 2)]
        except Exception as e:
            rec[f"cfg{cg}x{bn}_error"] = str(e)[:100]
    print(json.dumps(rec), flush=True)
    # timeline of one launch (eager, rotating weights so it streams from HBM), static and not
    for kf in (0, 0x100):
        setenv(1 << 24)
        for i in range(4): go(i, True, kf)
        torch.cuda.synchronize()
        go(5, True, kf)
        buf = (ctypes.c_ulonglong * 64)(); kb = (ctypes.c_ulonglong * 8)()
        assert raw.b200q_debug_read_trace(buf, 64) == 0 and raw.b200q_debug_read_ktrace(kb, 8) == 0
        t0 = kb[0]
        ev = [buf[e] for e in range(6)]
        print(json.dumps(dict(M=M, static=bool(kf), timeline_cycles_from_entry=dict(
            setup_done=kb[2] - t0, past_grid_dependency=kb[3] - t0, acc_owned=ev[0] - t0, first_ktile_landed=ev[1] - t0,
            last_mma_issued=ev[2] - t0, acc_complete=ev[3] - t0, drained=ev[4] - t0, stores_issued=ev[5] - t0, exit=kb[5] - t0),
            ns_entry_to_exit=kb[6] - kb[1], mhz=round((kb[5] - t0) / max(kb[6] - kb[1], 1) * 1e3, 1))), flush=True)
    setenv(0)
