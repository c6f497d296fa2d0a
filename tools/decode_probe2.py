#!/usr/bin/env python
"""Decode kernel (gemm_decode.cu, configuration (1,16)) vs the general kernel's (1,128) small-M tile.
B200Q_LIB=prof python tools/decode_probe2.py   -> JSON lines: graph-replay us (same / rotating weight buffers, static weights or
not), and the timeline of CTA 0 of one eager launch (cycles from kernel entry), also with the weights loaded for the first
ring only (1 << 22: what the kernel costs when no weight bytes move -- issue-bound floor)."""
import ctypes, json, os, sys
os.environ.setdefault("B200Q_LIB", "prof")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
raw = ctypes.CDLL(_lib.LIB_PATH)
raw.b200q_debug_read_decode_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
assert lib.b200q_profiling_build() == 1
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
    return round(min(ts), 2)


NAMES = ["entry", "gt_entry", "setup_done", "past_grid_dependency", "x_landed", "x_scales_copied", "first_weights_landed",
         "last_mma_issued", "acc_complete", "stores_issued", "exit", "gt_exit"]
for M in (1, 16, 32, 48, 64):
    a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
    sfa = torch.randint(126, 129, (128 * K // 32,), dtype=torch.uint8, device=dev)
    d = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(NSETS)]

    def go(i, rotate, kindflag, cg, bn):
        w, sfw = wsets[i % NSETS if rotate else 0]
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), w.data_ptr(), sfa.data_ptr(), sfw.data_ptr(), alpha.data_ptr(),
                                    d[i % NSETS].data_ptr(), M, N, K, kindflag, cg, bn, st)
        assert rc == 0, lib.b200q_last_error()

    setenv(0)
    rec = dict(M=M, N=N, K=K)
    for cg, bn in ((1, 16), (1, 128)):
        for rotate in (False, True):
            for kf in (0, 0x100):
                rec[f"cfg{cg}x{bn}_{'rot' if rotate else 'same'}_{'static' if kf else 'safe'}_us"] = graph_time(lambda i: go(i, rotate, kf, cg, bn), 24)
    print(json.dumps(rec), flush=True)
    for kf in (0, 0x100):
        for extra in (0, 1 << 22):
            setenv((1 << 24) | extra)
            for i in range(4): go(i, True, kf, 1, 16)
            torch.cuda.synchronize()
            go(5, True, kf, 1, 16)
            buf = (ctypes.c_ulonglong * 16)()
            assert raw.b200q_debug_read_decode_trace(buf, 16) == 0
            t0 = buf[0]
            tl = {NAMES[i]: buf[i] - t0 for i in range(2, 11) if i != 1}
            print(json.dumps(dict(M=M, static=bool(kf), weights_first_ring_only=bool(extra), timeline_cycles_from_entry=tl,
                                  ns_entry_to_exit=buf[11] - buf[1], mhz=round((buf[10] - t0) / max(buf[11] - buf[1], 1) * 1e3, 1))), flush=True)
    setenv(0)


# ---- the whole decode step through the Python surface: two calls (quantise, GEMM with static weights) vs ONE launch
sys.path.insert(0, ROOT)
import qutlass_b200 as Q
idx = torch.arange(128)
bits = idx[:, None] & idx[None, :]
par = torch.zeros_like(bits)
while bits.any():
    par ^= bits & 1
    bits = bits >> 1
R = ((1.0 - 2.0 * par.double()) * 128 ** -0.5).to(torch.bfloat16).to(dev)
w = torch.randn(N, K, dtype=torch.bfloat16, device=dev)
wq, wsf = Q.fusedQuantizeMx(w, R, method="abs_max")
wblk = Q.to_blocked(wsf)
for M in (1, 16, 32):
    x = torch.randn(M, K, dtype=torch.bfloat16, device=dev)
    Q.fusedQuantizeMx(x, R, method="abs_max")

    def two(i):
        q, s_ = Q.fusedQuantizeMx(x, R, method="abs_max")
        return Q.matmul_mxf4_bf16_tn(q, wq, Q.to_blocked(s_), wblk, alpha, static_weights=True)

    def two_safe(i):
        q, s_ = Q.fusedQuantizeMx(x, R, method="abs_max")
        return Q.matmul_mxf4_bf16_tn(q, wq, Q.to_blocked(s_), wblk, alpha)

    def one(i):
        return Q.fused_linear_fp4(x, R, wq, wblk, alpha)[0]

    os.environ["B200Q_FUSE_DECODE"] = "1"
    lib.b200q_reload_env()

    # timeline of CTA 0 of the ONE-launch step (eager, after a sync)
    setenv(1 << 24)
    os.environ["B200Q_FUSE_DECODE"] = "1"; lib.b200q_reload_env()
    for i in range(3): one(i)
    torch.cuda.synchronize()
    one(0)
    buf = (ctypes.c_ulonglong * 16)()
    assert raw.b200q_debug_read_decode_trace(buf, 16) == 0
    t0 = buf[0]
    print(json.dumps(dict(M=M, one_launch_timeline_cycles_from_entry={NAMES[i]: buf[i] - t0 for i in (2, 3, 4, 6, 7, 8, 9, 10)},
                          note="x_landed = MMA warp sees group 0 of the quantised activations", ns_entry_to_exit=buf[11] - buf[1])), flush=True)
    setenv(0)
    os.environ["B200Q_FUSE_DECODE"] = "1"; lib.b200q_reload_env()
    print(json.dumps(dict(M=M, step="quantise+GEMM, graph replay of 24, us per step",
                          two_calls_static=graph_time(two, 24), two_calls_safe=graph_time(two_safe, 24), one_launch=graph_time(one, 24))), flush=True)
