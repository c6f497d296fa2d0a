#!/usr/bin/env python
"""What does this B200 sustain in FP4 tensor work, and what limits it?   B200Q_LIB=prof python tools/fp4_peak_probe.py
Config 1 (M=4096, N=14336, K=4096) through the profiling build, one variant after the other on the same box:
  * schedules: (2,256) default, (2,448) hybrid (no accumulator hand-off bubble, no ragged last round), (2,192), cluster-of-4 multicast;
  * ablations (timing only, WRONG results): no D stores (flag 1), no A/B tile loads (3 << 20), no scale copies (32);
  * operand data: random codes vs ALL-ZERO codes and scales (same instruction stream, minimal datapath toggling) -- a
    power-limited part runs the zero case faster at the same clock-independent cycle count.
For every variant: us per launch over 300 back-to-back launches, and over a 1.5 s sustained run; CTA 0's cycle count and
wall time inside one launch (clock64 / globaltimer) = the SM clock the kernel actually ran at."""
import ctypes, json, os, sys, time
os.environ.setdefault("B200Q_LIB", "prof")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
raw = ctypes.CDLL(_lib.LIB_PATH)
raw.b200q_debug_read_ktrace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
assert lib.b200q_profiling_build() == 1, "needs lib/libb200q_prof.so"
M, N, K = 4096, 14336, 4096
FL = 2.0 * M * N * K
alpha = torch.ones(1, device=dev)
st = torch.cuda.current_stream().cuda_stream


def mk(zero):
    sets = []
    for i in range(3):
        if zero:
            a = torch.zeros(M, K // 2, dtype=torch.uint8, device=dev); b = torch.zeros(N, K // 2, dtype=torch.uint8, device=dev)
            sfa = torch.full((M * K // 32,), 127, dtype=torch.uint8, device=dev); sfb = torch.full((N * K // 32,), 127, dtype=torch.uint8, device=dev)
        else:
            a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev); b = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev)
            sfa = torch.randint(126, 129, (M * K // 32,), dtype=torch.uint8, device=dev); sfb = torch.randint(126, 129, (N * K // 32,), dtype=torch.uint8, device=dev)
        sets.append((a, b, sfa, sfb, torch.empty(M, N, dtype=torch.bfloat16, device=dev)))
    return sets


def setenv(flags, hybrid=False):
    os.environ["B200Q_GEMM_DEBUG_FLAGS"] = str(flags)
    os.environ["B200Q_GEMM_HYBRID"] = "1" if hybrid else "0"
    lib.b200q_reload_env()


def run(sets, cg, bn, n):
    for i in range(n):
        a, b, sfa, sfb, d = sets[i % 3]
        rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(), d.data_ptr(),
                                    M, N, K, 0x100, cg, bn, st)
        assert rc == 0, lib.b200q_last_error()


def timed(sets, cg, bn, n):
    run(sets, cg, bn, 5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(sets, cg, bn, n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


data = {"random": mk(False), "zero": mk(True)}
variants = [("2x256", 2, 256, 0, False), ("hybrid 2x448", 2, 448, 0, False), ("2x192", 2, 192, 0, False), ("mc4 2x256", 4, 256, 0, False),
            ("2x256 no stores", 2, 256, 1, False), ("2x256 no stores, no A/B loads", 2, 256, 1 | (3 << 20), False),
            ("2x256 no stores, no A/B loads, no scale copies", 2, 256, 1 | (3 << 20) | 32, False),
            ("2x192 no stores, no A/B loads", 2, 192, 1 | (3 << 20), False),
            # operands loaded ONCE (first ring), then reused from shared memory: realistic data, no operand traffic
            ("2x256 resident operands, no stores", 2, 256, 1 | (1 << 22), False),
            ("2x256 resident operands, stores on", 2, 256, (1 << 22), False),
            ("2x192 resident operands, no stores", 2, 192, 1 | (1 << 22), False),
            ("2x256 resident operands, no stores, no scale copies", 2, 256, 1 | 32 | (1 << 22), False)]
for dname, sets in data.items():
    for name, cg, bn, flags, hyb in variants:
        if dname == "zero" and ("no A/B" in name or "resident" in name):
            continue
        try:
            setenv(flags)
            us = timed(sets, cg, bn, 300)
            n_sus = int(1.5e6 / us)
            us_sus = timed(sets, cg, bn, n_sus)
            setenv(flags | (1 << 24))
            run(sets, cg, bn, 3); torch.cuda.synchronize()
            kb = (ctypes.c_ulonglong * 8)()
            raw.b200q_debug_read_ktrace(kb, 8)
            cyc, ns = kb[5] - kb[0], kb[6] - kb[1]
            print(json.dumps(dict(data=dname, variant=name, us=round(us, 2), tflops=round(FL / us / 1e6, 0), us_sustained=round(us_sus, 2),
                                  tflops_sustained=round(FL / us_sus / 1e6, 0), cta0_cycles=cyc, cta0_ns=ns,
                                  sm_mhz_in_kernel=round(cyc / max(ns, 1) * 1e3, 0))), flush=True)
        except Exception as e:
            print(json.dumps(dict(data=dname, variant=name, error=str(e)[:200])), flush=True)
        time.sleep(0.5)
setenv(0)
