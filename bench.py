#!/usr/bin/env python
"""bench.py -- headline benchmark of the microscaled-FP4 hot path (driver contract).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--kind mx|nv] [--had H]

Workload (BASELINE.json configs[1]): Llama-3-8B FFN, M=4096, N=14336, K=4096, MXFP4 W4A4 abs_max.
One "step" is the reference's "actual" benchmark iteration (benchmarks/bench_mxfp4_sm100.py:93-106):
    fusedQuantizeMx(a, H, "abs_max")  ->  to_blocked (a no-op here)  ->  matmul_mxf4_bf16_tn
with the weights quantised once outside the loop.  metric = effective TFLOP/s = 2*M*N*K / t.

N > 1 (torchrun): weak scaling -- every rank owns its own 4096 activation rows (global M = 4096*N),
weights are quantised on rank 0 and broadcast ONCE over NCCL at setup; no collective in the timed loop.

--impl reference: the reference's CPU-side oracle path (unpack e2m1 * scale -> fp32 torch.matmul -> bf16)
timed on the host cores (oracle/cpu_baseline.py); rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, N, K = 4096, 14336, 4096
METRIC = "MXFP4 GEMM effective TFLOPS on Llama-3-8B FFN shapes; % of B200 FP4 peak"
NOMINAL_FP4_TFLOPS = 9000.0


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (one `nvidia-smi -lms 20` process)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, interval_ms: int = 20):
        self.index = index
        self.interval_ms = interval_ms
        self.proc = None
        self.path = os.path.join(ROOT, "gpurun_out", f"bench_clocks_gpu{index}.csv") if os.path.isdir(
            os.path.join(ROOT, "gpurun_out")) else f"/tmp/bench_clocks_gpu{index}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.interval_ms)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.25)   # let the first samples land before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [[x.strip() for x in l.split(",")] for l in open(self.path) if l.strip()]
        rows = [r for r in rows if len(r) >= 7]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        pw = [float(r[2]) for r in rows]
        load = [r for r in rows if float(r[2]) >= 0.6 * max(pw)] or rows   # samples taken under load
        sm = sorted(float(r[0]) for r in load)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in load)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(rows[0][1]), reasons=reasons, samples=len(load),
                    power_w_max=max(pw))


def _ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed
    `ncu --set full` capture of this same command (profiles/rNN_ncu_gemm_summary.txt); (None, None) if absent.  ncu cannot run
    inside a timed bench, so this is a recorded figure and the line says which file it came from."""
    cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_ncu_gemm_summary.txt")) \
        if os.path.isdir(os.path.join(ROOT, "profiles")) else []
    if not cands:
        return None, None
    p = os.path.join(ROOT, "profiles", cands[-1])
    try:
        import re
        line = open(p).readline()
        rd = re.search(r"dram__bytes_read.sum=([0-9.]+) (\w+)", line)
        wr = re.search(r"dram__bytes_write.sum=([0-9.]+) (\w+)", line)
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)], "profiles/" + cands[-1]
    except Exception:
        return None, None


def _bind_to_gpu_numa_node(local: int):
    """Best effort: run this rank (and therefore first-touch its pinned host buffers) on the CPUs of the NUMA node the
    GPU hangs off, so that N ranks do not all stage their PCIe traffic through node 0.  Returns the node or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        devn = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devn:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def _reference_gpu(kind, had, steps):
    """GEMM / quantise / step times of the reference's CUTLASS kernels, measured by oracle/ref_gpu.py in a child process
    AFTER our own timed regions (same box, same shape, same rotating-buffer policy)."""
    try:
        from oracle import ref_gpu
        if not ref_gpu.available():
            return {"unavailable": "oracle/_ref/qutlass_ref_C.so not built (python oracle/build_ref.py)"}
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_gpu.py"), "bench", kind, str(M), str(N), str(K),
                            str(had), str(steps)], capture_output=True, text=True, timeout=180)
        for l in reversed(r.stdout.strip().splitlines()):
            if l.startswith("{"):
                return json.loads(l)
        return {"unavailable": ("child failed: " + (r.stderr or r.stdout)[-300:]).replace("\n", " ")}
    except Exception as e:   # noqa: BLE001 -- a comparison leg must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def run_reference(args):
    """CPU arm: the reference's test-oracle path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import cpu_baseline as C
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would otherwise pin the
    # reference arm to one core when it is launched for N > 1)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncpu))
    kind = args.kind
    # bounded sample: the full config costs a few seconds per step on a many-core host; cap steps
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 1))
    m_s = 64 if os.environ.get("B200Q_BENCH_TINY") else M   # tiny mode: contract self-test only
    r = C.time_cpu_path(m_s, N, K, kind, steps=steps, warmup=warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["tflops"], "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp4 e2m1 x e2m1 -> fp32 accumulate -> bf16 (CPU: fp32 matmul of dequantised operands)",
        "data": "synthetic",
        "config": {"workload": f"Llama-3-8B FFN M={M} N={N} K={K} {'MXFP4' if kind == 'mx' else 'NVFP4'} W4A4, "
                               "CPU dequantise(A,B)+torch.matmul fp32 -> bf16"},
        "cpu_baseline": {"value": r["tflops"], "unit": "TFLOP/s", "cores": r["threads"], "kind": "port",
                         "sample": f"{steps} step(s) of M={m_s} rows x full N,K (dequant LUT + fp32 matmul)"},
        "e2e": {"value": r["tflops"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--kind", default="mx", choices=["mx", "nv"])
    ap.add_argument("--had", type=int, default=128)
    ap.add_argument("--cta-group", type=int, default=0)
    ap.add_argument("--block-n", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-CUTLASS-kernel comparison leg (child process)")
    ap.add_argument("--no-c4", action="store_true", help="skip the configs[4] (70B FFN, M=16384 sharded) leg")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained leg in seconds (0 = skip)")
    ap.add_argument("--fuse", action="store_true", help="step = the single fused quantise+GEMM kernel (B200Q_FUSE=1; measured slower)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import qutlass_b200 as Q
    from qutlass_b200 import _lib

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: qutlass_b200 has no CPU path"}))
        return 2
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:
        orig_affinity = os.sched_getaffinity(0)
    except Exception:
        orig_affinity = None
    numa_node = _bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    kind = args.kind
    knd = Q.KIND_MXF4 if kind == "mx" else Q.KIND_NVF4
    sf_dtype = torch.float8_e8m0fnu if kind == "mx" else torch.float8_e4m3fn
    group = 32 if kind == "mx" else 16

    # ---------------- setup (untimed): rotation, weights quantised once on rank 0, ONE broadcast
    torch.manual_seed(1234 + rank)
    # Sylvester Hadamard * H^-1/2 in bf16 (what scipy.linalg.hadamard gives the reference benchmark, bench_mxfp4_sm100.py:51-54)
    idx = torch.arange(args.had)
    bits = (idx[:, None] & idx[None, :])
    par = torch.zeros_like(bits)
    while bits.any():
        par ^= bits & 1
        bits = bits >> 1
    H = ((1.0 - 2.0 * par.double()) * args.had ** -0.5).to(torch.bfloat16).to(dev)
    gs = torch.tensor([1.0], dtype=torch.float32, device=dev)
    alpha = torch.tensor([1.0], dtype=torch.float32, device=dev)
    NSETS = 4  # rotate buffer sets so every timed iteration reads/writes data that is not L2 resident
    wq = torch.empty(N, K // 2, dtype=torch.uint8, device=dev)
    wsf = torch.empty(((N + 127) // 128) * 128 * (((K // group) + 3) // 4) * 4, dtype=sf_dtype, device=dev)
    if rank == 0:
        w = torch.randn(N, K, dtype=torch.bfloat16, device=dev)
        if kind == "mx":
            q_, s_ = Q.fusedQuantizeMx(w, H, method="abs_max")
        else:
            q_, s_ = Q.fusedQuantizeNv(w, H, gs, method="abs_max")
        wq.copy_(q_)
        wsf.copy_(Q.to_blocked(s_))
        del w, q_, s_
    if world > 1:
        dist.broadcast(wq, 0)
        wsf_u8 = wsf.view(torch.uint8)
        dist.broadcast(wsf_u8, 0)
    wqs = [wq] + [wq.clone() for _ in range(NSETS - 1)]
    wsfs = [wsf] + [wsf.clone() for _ in range(NSETS - 1)]
    acts = [torch.randn(M, K, dtype=torch.bfloat16, device=dev) for _ in range(NSETS)]
    pr, pc = ((M + 127) // 128) * 128, (((K // group) + 3) // 4) * 4
    aqs = [torch.empty(M, K // 2, dtype=torch.uint8, device=dev) for _ in range(NSETS)]
    asfs = [torch.empty(pr * pc, dtype=sf_dtype, device=dev) for _ in range(NSETS)]
    outs = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(NSETS)]
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    method = Q.METHOD_ABSMAX | Q.ROT_TRUSTED_HADAMARD   # H is built right above: the caller-side hint of b200q.h

    def quant(i):
        s = i % NSETS
        if kind == "mx":
            rc = lib.b200q_quantize_mx(acts[s].data_ptr(), H.data_ptr(), aqs[s].data_ptr(), None, asfs[s].data_ptr(), None,
                                       M * K, K, args.had, method, stream)
        else:
            rc = lib.b200q_quantize_nv(acts[s].data_ptr(), H.data_ptr(), aqs[s].data_ptr(), None, asfs[s].data_ptr(),
                                       gs.data_ptr(), M * K, K, args.had, method, stream)
        _lib.check(rc)

    def gemm(i):
        s = i % NSETS
        # the weights were quantised once at setup: the caller-side promise of b200q.h (B200Q_GEMM_STATIC_WEIGHTS)
        _lib.check(lib.b200q_gemm_fp4_cfg(aqs[s].data_ptr(), wqs[s].data_ptr(), asfs[s].data_ptr(), wsfs[s].data_ptr(),
                                          alpha.data_ptr(), outs[s].data_ptr(), M, N, K, knd | Q.GEMM_STATIC_WEIGHTS,
                                          args.cta_group, args.block_n, stream))

    def step_two_launches(i):
        quant(i)
        gemm(i)

    # the whole step through ONE C-ABI call: one persistent kernel (quantiser warps inside the GEMM) when eligible
    fuse_ws = torch.zeros(max(int(lib.b200q_linear_fp4_workspace_bytes(M)), 256), dtype=torch.uint8, device=dev)
    fused = args.fuse and args.cta_group == 0
    if fused:
        os.environ["B200Q_FUSE"] = "1"
        lib.b200q_reload_env()
    launches_per_step = lib.b200q_linear_fp4_launches(M, N, K, args.had, method, knd) if fused else \
        1 + (lib.b200q_gemm_fp4_launches(M, N, K, knd) if args.cta_group == 0 else 1)

    def step_fused(i):
        s = i % NSETS
        _lib.check(lib.b200q_linear_fp4(acts[s].data_ptr(), H.data_ptr(), aqs[s].data_ptr(), None, asfs[s].data_ptr(),
                                        wqs[s].data_ptr(), wsfs[s].data_ptr(), alpha.data_ptr(),
                                        gs.data_ptr() if kind == "nv" else None, outs[s].data_ptr(), fuse_ws.data_ptr(),
                                        M, N, K, args.had, method, knd, stream))

    step = step_fused if fused else step_two_launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, per_rank=False):
        """W untimed warm-up steps, then EXACTLY `steps` steps between two CUDA events on the launching stream, bracketed by
        barrier + synchronize on both sides; returns the MAX over ranks (ms per step) [and every rank's own figure]."""
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        mine = ms / steps
        ranks = [mine]
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            allt = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            ranks = [float(x.item()) / steps for x in allt]
            ms = max(float(x.item()) for x in allt)
        return (ms / steps, ranks) if per_rank else ms / steps

    def gather_obj(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    flops = 2.0 * M * N * K
    # every rank samples ITS OWN GPU (the 2.6 % weak-scaling loss needs per-GPU clocks)
    # (rank 0 at 20 ms like round 1; the other ranks at 100 ms so that N samplers do not compete with N launch loops for the
    # host's cores -- their burst leg may see few samples, the 2 s sustained leg sees ~20)
    smp_ms = 20 if rank == 0 else 100
    sampler = ClockSampler(local, smp_ms)
    sampler.start()
    ms_step, ms_step_ranks = timed(step, args.steps, args.warmup, per_rank=True)
    clocks = sampler.stop()
    clocks_all = gather_obj(clocks)
    # dominant kernel alone (CUDA events on the launching stream), and the quantise kernel alone
    ms_gemm, ms_gemm_ranks = timed(gemm, args.steps, 3, per_rank=True)
    ms_quant, ms_quant_ranks = timed(quant, args.steps, 3, per_rank=True)
    ms_two = timed(step_two_launches, args.steps, 3) if fused else ms_step

    # ---------------- sustained leg: the same step back to back for >= args.sustain_s seconds (power / thermal steady state)
    sus = None
    if args.sustain_s > 0:
        n_sus = int(min(200000, max(args.steps, args.sustain_s * 1e3 / ms_step)))
        sampler = ClockSampler(local, smp_ms)
        sampler.start()
        ms_sus, ms_sus_ranks = timed(step, n_sus, 0, per_rank=True)
        clocks_sus = sampler.stop()
        n_g = int(min(200000, max(args.steps, 0.5 * args.sustain_s * 1e3 / ms_gemm)))
        ms_gemm_sus = timed(gemm, n_g, 0)
        sus = dict(steps=n_sus, ms_per_step=ms_sus, ms_per_rank=ms_sus_ranks, clocks=clocks_sus, gemm_steps=n_g,
                   ms_gemm=ms_gemm_sus, clocks_all=gather_obj(clocks_sus))

    value = flops * world / (ms_step * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp4 (e2m1 x e2m1, block-scaled, fp32 accumulate, bf16 out)", "data": "synthetic",
        "config": {
            "workload": f"Llama-3-8B FFN M={M} (per GPU) N={N} K={K} {'MXFP4' if kind == 'mx' else 'NVFP4'} W4A4 abs_max, "
                        f"step = fused Hadamard-{args.had} rotate+quantise of activations + block-scaled FP4 GEMM "
                        "(weights pre-quantised -> B200Q_GEMM_STATIC_WEIGHTS)" + (", one persistent kernel (b200q_linear_fp4)" if launches_per_step == 1 else ""),
            "global_batch_rows": M * world, "parallelism": f"dp{world} (M-sharded, weights broadcast once at setup)",
            "l2_policy": f"rotating {NSETS} buffer sets (activations/outputs/weights), {NSETS * 190} MB footprint > 126 MB L2",
        },
        "gpu_launches": launches_per_step * args.steps,
        "two_launch_step_us": ms_two * 1e3,
        "gemm_only_tflops_per_gpu": flops / (ms_gemm * 1e-3) / 1e12,
        "quantize_us": ms_quant * 1e3,
        "per_rank": {"step_ms": ms_step_ranks, "gemm_ms": ms_gemm_ranks, "quantize_ms": ms_quant_ranks,
                     "clocks": clocks_all},
    }
    pk = _peaks()
    if sus is not None:
        line["value_sustained"] = flops * world / (sus["ms_per_step"] * 1e-3) / 1e12
        line["sustained"] = {
            "seconds": sus["steps"] * sus["ms_per_step"] * 1e-3, "steps": sus["steps"], "ms_per_step": sus["ms_per_step"],
            "ms_per_rank": sus["ms_per_rank"], "clocks": sus["clocks"], "clocks_per_rank": sus["clocks_all"],
            "gemm_only_tflops_per_gpu": flops / (sus["ms_gemm"] * 1e-3) / 1e12, "gemm_steps": sus["gemm_steps"],
            "vs_burst": (ms_step / sus["ms_per_step"]),
        }
    if rank == 0:
        # MEASURED_PEAKS.json has no FP4 figure.  kind::mxf4 retires 4x the MACs of kind::f16 per tcgen05.mma issue slot, so the
        # denominator is 4 x the measured cuBLAS bf16 rate -- the BURST figure for the kernel timed alone in a short loop, the
        # SUSTAINED one for the >= 2 s leg (B200_PROFILING.md).  profiles/r02_fp4_peak.md holds the directly measured FP4
        # tensor ceiling of this pool's parts under their power limit (MMA-only probe) next to it.
        fp4_peak = 4.0 * pk["bf16"]
        ach = flops / (ms_gemm * 1e-3) / 1e12
        traffic, traffic_src = _ncu_traffic_bytes() if (kind == "mx") else (None, None)
        line["roofline"] = {
            "bound": "tensor", "achieved": ach, "peak": fp4_peak, "unit": "TFLOP/s", "frac": ach / fp4_peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "traffic_note": "DRAM bytes per launch from the committed ncu --set full capture named in traffic_source (ncu cannot "
                            f"run inside a timed bench); algorithmic bytes = {M * K // 2 + N * K // 2 + (M + N) * K // group + 2 * M * N}",
            "peak_basis": f"4 x {pk['source']} burst cuBLAS bf16 ({pk['bf16']} TF/s) -- of measured; "
                          "no FP4 figure in MEASURED_PEAKS.json",
            "frac_of_nominal_9PF": ach / NOMINAL_FP4_TFLOPS,
            "frac_of_4x_sustained_bf16": ach / (4.0 * pk["bf16_sustained"]),
            "kernel": "gemm_fp4_kernel", "algorithmic_flops_per_launch": flops,
            "quantize_hbm": {
                "bound": "hbm", "achieved": (M * K * (2 + 0.5 + 1.0 / group)) / (ms_quant * 1e-3) / 1e9,
                "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": (M * K * (2 + 0.5 + 1.0 / group)) / (ms_quant * 1e-3) / 1e9 / pk["hbm_gbs"],
            },
        }
        # measured with tools/fp4_peak_probe.py on a B200 of this pool (profiles/r02_fp4_peak.md): FP4 MMAs on operands resident
        # in shared memory, double-buffered accumulators, nothing loaded or stored -- the ceiling of ANY FP4 kernel under the
        # part's 1 kW power limit with realistic (random) operand data.  A recorded figure, like `traffic`.
        line["roofline"]["fp4_mma_only_ceiling"] = {"burst": 7262.0, "sustained": 6374.0, "unit": "TFLOP/s",
                                                    "source": "profiles/r02_fp4_peak.md", "frac_burst": ach / 7262.0}
        if sus is not None:
            line["roofline"]["fp4_mma_only_ceiling"]["frac_sustained"] = line["sustained"]["gemm_only_tflops_per_gpu"] / 6374.0
        if sus is not None:
            ach_s = line["sustained"]["gemm_only_tflops_per_gpu"]
            line["roofline"]["achieved_sustained"] = ach_s
            line["roofline"]["peak_sustained"] = 4.0 * pk["bf16_sustained"]
            line["roofline"]["frac_sustained"] = ach_s / (4.0 * pk["bf16_sustained"])
        line["clocks"] = clocks

    # ---------------- configs[4]: Llama-3-70B FFN (N=28672, K=8192), global M=16384 SHARDED over the ranks (strong scaling)
    if not args.no_c4:
        from qutlass_b200.sharding import shard_rows
        N4, K4, M4 = 28672, 8192, 16384
        r0, rows4 = shard_rows(M4, world, rank)
        w4q = torch.empty(N4, K4 // 2, dtype=torch.uint8, device=dev)
        w4sf = torch.empty(((N4 + 127) // 128) * 128 * (((K4 // group) + 3) // 4) * 4, dtype=sf_dtype, device=dev)
        if rank == 0:
            w4 = torch.randn(N4, K4, dtype=torch.bfloat16, device=dev)
            q_, s_ = Q.fusedQuantizeMx(w4, H, method="abs_max") if kind == "mx" else Q.fusedQuantizeNv(w4, H, gs, method="abs_max")
            w4q.copy_(q_)
            w4sf.copy_(Q.to_blocked(s_))
            del w4, q_, s_
        from qutlass_b200.sharding import broadcast_weights
        broadcast_weights(w4q, w4sf, src=0)
        S4 = 2   # two weight / activation / output sets: 2 x (125 MB weights + rows4 x 73 KB) > L2
        w4qs, w4sfs = [w4q, w4q.clone()], [w4sf, w4sf.clone()]
        x4 = [torch.randn(max(rows4, 1), K4, dtype=torch.bfloat16, device=dev) for _ in range(S4)]
        pr4 = ((max(rows4, 1) + 127) // 128) * 128
        pc4 = (((K4 // group) + 3) // 4) * 4
        x4q = [torch.empty(max(rows4, 1), K4 // 2, dtype=torch.uint8, device=dev) for _ in range(S4)]
        x4sf = [torch.empty(pr4 * pc4, dtype=sf_dtype, device=dev) for _ in range(S4)]
        d4 = [torch.empty(max(rows4, 1), N4, dtype=torch.bfloat16, device=dev) for _ in range(S4)]

        def quant4(i):
            s_ = i % S4
            if kind == "mx":
                rc = lib.b200q_quantize_mx(x4[s_].data_ptr(), H.data_ptr(), x4q[s_].data_ptr(), None, x4sf[s_].data_ptr(), None,
                                           rows4 * K4, K4, args.had, method, stream)
            else:
                rc = lib.b200q_quantize_nv(x4[s_].data_ptr(), H.data_ptr(), x4q[s_].data_ptr(), None, x4sf[s_].data_ptr(),
                                           gs.data_ptr(), rows4 * K4, K4, args.had, method, stream)
            _lib.check(rc)

        def gemm4(i):
            s_ = i % S4
            _lib.check(lib.b200q_gemm_fp4(x4q[s_].data_ptr(), w4qs[s_].data_ptr(), x4sf[s_].data_ptr(), w4sfs[s_].data_ptr(),
                                          alpha.data_ptr(), d4[s_].data_ptr(), rows4, N4, K4, knd | Q.GEMM_STATIC_WEIGHTS, stream))

        def step4(i):
            if rows4 > 0:
                quant4(i)
                gemm4(i)

        st4 = max(10, min(args.steps, 100))
        ms4, ms4_ranks = timed(step4, st4, 3, per_rank=True)
        ms4_gemm = timed(lambda i: gemm4(i) if rows4 > 0 else None, st4, 3)
        flops4 = 2.0 * M4 * N4 * K4
        line["c4"] = {
            "workload": f"Llama-3-70B FFN N={N4} K={K4} {'MXFP4' if kind == 'mx' else 'NVFP4'}, global M={M4} sharded by "
                        f"qutlass_b200.sharding.shard_rows over {world} rank(s) ({rows4} rows on rank 0), quantise + GEMM",
            "scaling": "strong", "value": flops4 / (ms4 * 1e-3) / 1e12, "unit": "TFLOP/s (aggregate, max-over-ranks time)",
            "ms_per_step": ms4, "ms_per_rank": ms4_ranks, "steps": st4, "rows_per_rank": [shard_rows(M4, world, r)[1] for r in range(world)],
            "gemm_only_ms": ms4_gemm, "gemm_only_tflops_aggregate": flops4 / (ms4_gemm * 1e-3) / 1e12,
        }
        del w4qs, w4sfs, x4, x4q, x4sf, d4, w4q, w4sf
        torch.cuda.empty_cache()

    # ---------------- e2e: host buffers through the C-ABI (H2D + quantise + GEMM + D2H inside the timed region)
    if not args.no_e2e:
        ws_bytes = lib.b200q_linear_workspace_bytes(M, N, K, knd)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        x_host = torch.randn(M, K, dtype=torch.bfloat16).pin_memory()
        d_host = torch.empty(M, N, dtype=torch.bfloat16).pin_memory()
        d_host.zero_()                       # first touch on this rank's (NUMA-bound) CPUs

        def e2e_step(i):
            _lib.check(lib.b200q_linear_fp4_host(x_host.data_ptr(), H.data_ptr(), wqs[i % NSETS].data_ptr(),
                                                 wsfs[i % NSETS].data_ptr(), alpha.data_ptr(), gs.data_ptr(),
                                                 d_host.data_ptr(), ws.data_ptr(), M, N, K, args.had, knd, stream))

        e2e_steps = max(3, min(args.steps, 20))
        ms_e2e, ms_e2e_ranks = timed(e2e_step, e2e_steps, 3, per_rank=True)
        # the host ceiling: the SAME bytes (H2D of x, D2H of D) with no kernels at all, both directions concurrently, all
        # ranks at once -- what the PCIe / host-memory path of this box sustains for this traffic pattern
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        x_dev = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
        cur = torch.cuda.current_stream()

        def copy_only(i):
            ev = torch.cuda.Event()
            ev.record(cur)
            s_up.wait_event(ev)
            s_dn.wait_event(ev)
            with torch.cuda.stream(s_up):
                x_dev.copy_(x_host, non_blocking=True)
            with torch.cuda.stream(s_dn):
                d_host.copy_(outs[i % NSETS], non_blocking=True)
            cur.wait_stream(s_up)
            cur.wait_stream(s_dn)

        ms_copy = timed(copy_only, e2e_steps, 2)
        line["e2e"] = {"value": flops * world / (ms_e2e * 1e-3) / 1e12, "unit": "TFLOP/s",
                       "h2d_bytes_per_step": M * K * 2, "d2h_bytes_per_step": M * N * 2, "ms_per_step": ms_e2e,
                       "ms_per_rank": ms_e2e_ranks, "numa_node": numa_node,
                       "host_copy_ceiling_ms": ms_copy,
                       "host_copy_ceiling_note": "same H2D + D2H bytes per rank, no kernels, both directions concurrent, all ranks "
                                                 "at once: the PCIe / host-memory bound of this box for the e2e traffic",
                       "frac_of_host_ceiling": ms_copy / ms_e2e,
                       "api": "b200q_linear_fp4_host (C-ABI, pinned host buffers, abs_max)"}

    # ---------------- cpu_baseline (rank 0, N=1 only): bounded sample on the host cores
    # (the baseline leg is the one place of this arm that executes oracle/: the CPU port, and -- reported beside it, after
    # every timed region of ours -- the reference's own CUDA kernels: oracle/_ref/qutlass_ref_C.so = the unmodified reference
    # sources compiled for sm_100a by oracle/build_ref.py, run in a child process because it registers the same torch op
    # namespace as our drop-in.  A reported comparison on the same box and shape; it can never fail the bench line.)
    if rank == 0 and world == 1 and not args.no_cpu:
        if orig_affinity is not None:
            try:
                os.sched_setaffinity(0, orig_affinity)      # the CPU baseline gets every core, not just the GPU's NUMA node
            except Exception:
                pass
        from oracle import cpu_baseline as C
        r = C.time_cpu_path(M, N, K, kind, steps=1, warmup=1)
        line["cpu_baseline"] = {"value": r["tflops"], "unit": "TFLOP/s", "cores": r["threads"], "kind": "port",
                                "sample": f"1 step of the full config (M={M}): LUT dequantise A,B + fp32 torch.matmul + bf16"}
    if rank == 0 and not args.no_ref_gpu and not args.no_cpu:
        # every N: rank 0's GPU runs the compiled reference kernels while the other ranks wait at the barrier below
        line["reference_gpu"] = _reference_gpu(kind, args.had, min(args.steps, 200))
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
