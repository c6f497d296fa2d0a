#!/usr/bin/env python
"""Sustained-load probe: run one GEMM config back to back for ~1.5 s while nvidia-smi samples clocks/power."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
M, N = 4096, 14336

def sample_start(path):
    q = "clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"
    return subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                            stdout=open(path, "w"), stderr=subprocess.DEVNULL)

def run(kind, K, cg, bn, secs=1.5):
    knd = 0 if kind == "mx" else 1
    group = 32 if kind == "mx" else 16
    a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
    b = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev)
    lo, hi = (126, 129) if kind == "mx" else (0x30, 0x41)
    sfa = torch.randint(lo, hi, (M * (K // group),), dtype=torch.uint8, device=dev)
    sfb = torch.randint(lo, hi, (N * (K // group),), dtype=torch.uint8, device=dev)
    d = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    alpha = torch.ones(1, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def go():
        rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(), d.data_ptr(),
                                    M, N, K, knd, cg, bn, st)
        assert rc == 0, lib.b200q_last_error()
    for _ in range(5): go()
    torch.cuda.synchronize()
    # calibrate iteration count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [go() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 10 * 1e-3
    iters = max(20, int(secs / per))
    path = os.path.join(ROOT, "gpurun_out", f"clk_{kind}_{K}_{cg}_{bn}.csv")
    p = sample_start(path)
    time.sleep(0.15)
    e0.record()
    for _ in range(iters): go()
    e1.record(); torch.cuda.synchronize()
    time.sleep(0.05)
    p.terminate(); p.wait()
    ms = e0.elapsed_time(e1) / iters
    rows = [l.strip().split(", ") for l in open(path) if l.strip()]
    rows = [r for r in rows if len(r) >= 5]
    # keep samples under load (power above 60% of max seen)
    pw = [float(r[2]) for r in rows]
    thr = 0.6 * max(pw) if pw else 0
    load = [r for r in rows if float(r[2]) >= thr]
    sm = sorted(float(r[0]) for r in load)
    out = dict(kind=kind, K=K, cg=cg, bn=bn, us=round(ms * 1e3, 1), tflops=round(2.0 * M * N * K / ms / 1e9, 0), iters=iters,
               samples=len(load), sm_mhz_median=sm[len(sm) // 2] if sm else None, sm_mhz_min=sm[0] if sm else None,
               power_w_median=sorted(float(r[2]) for r in load)[len(load) // 2] if load else None,
               power_w_max=max(pw) if pw else None, temp_max=max(float(r[3]) for r in rows) if rows else None,
               sw_power_cap=any(r[4].startswith("Active") for r in load))
    print(json.dumps(out), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "power_probe.jsonl"), "a") as f:
        f.write(json.dumps(out) + "\n")

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for flags in ("0", "3"):
        os.environ["B200Q_GEMM_DEBUG_FLAGS"] = flags
        print("# flags", flags)
        for (cg, bn) in ((1, 256), (2, 256), (1, 128), (2, 128)):
            run("mx", 4096, cg, bn)
    os.environ["B200Q_GEMM_DEBUG_FLAGS"] = "0"
    run("mx", 16384, 1, 256)
    run("mx", 16384, 2, 256)
    run("nv", 4096, 2, 256)
    # bf16 cuBLAS reference point for the same sampling method
    x = torch.randn(8192, 8192, dtype=torch.bfloat16, device=dev); y = torch.randn(8192, 8192, dtype=torch.bfloat16, device=dev)
    path = os.path.join(ROOT, "gpurun_out", "clk_cublas.csv")
    for _ in range(3): torch.matmul(x, y)
    torch.cuda.synchronize()
    p = sample_start(path); time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(1500): torch.matmul(x, y)
    e1.record(); torch.cuda.synchronize(); p.terminate(); p.wait()
    ms = e0.elapsed_time(e1) / 1500
    rows = [l.strip().split(", ") for l in open(path) if l.strip()]
    pw = [float(r[2]) for r in rows]; load = [r for r in rows if float(r[2]) >= 0.6 * max(pw)]
    sm = sorted(float(r[0]) for r in load)
    print(json.dumps(dict(kind="cublas_bf16_8192", tflops=round(2 * 8192**3 / ms / 1e9, 0), sm_mhz_median=sm[len(sm)//2], power_w_max=max(pw))))
