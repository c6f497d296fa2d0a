#!/bin/bash
# Round 2, GPU call 2: the whole GPU suite on the new build (PDL-safe GEMM, NV-128 default, reference suite unmodified),
# smoke, bench lines (MX / NV, with c4 + sustained legs), reference M sweep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -25 gpurun_out/r02_pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
echo "== bench mx"; timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.err
echo "== bench nv"; timeout 600 python bench.py --kind nv > gpurun_out/r02_bench_n1_nv.json 2> gpurun_out/r02_bench_n1_nv.err; tail -c 600 gpurun_out/r02_bench_n1_nv.err
python - <<'PY'
import json
for n in ("r02_bench_n1", "r02_bench_n1_nv"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, "step %.1f us" % (d["ms_per_step"] * 1e3), "value %.0f" % d["value"], "sustained %.0f" % d.get("value_sustained", 0),
              "gemm %.0f TF" % d["gemm_only_tflops_per_gpu"], "quant %.2f us" % d["quantize_us"], "c4", d.get("c4", {}).get("value"),
              "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("host_copy_ceiling_ms"), d.get("clocks"))
        print("   ref_gpu", d.get("reference_gpu"))
    except Exception as e:
        print(n, "no line:", e)
PY
echo "== ref msweep"; timeout 500 python tools/ref_msweep.py > gpurun_out/r02_ref_msweep.jsonl 2> gpurun_out/r02_ref_msweep.err; tail -3 gpurun_out/r02_ref_msweep.err; cat gpurun_out/ref_msweep.md
