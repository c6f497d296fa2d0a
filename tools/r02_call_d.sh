#!/bin/bash
# Session-3 call D: validation of the final library -- full GPU suite, smoke, bench lines (MX, NV, reference arm), ncu launch
# list of the bench command, decode + backward numbers of the final build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_final_gpu.txt
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_pytest_gpu.log; tail -4 gpurun_out/r02_final_pytest_gpu.log
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_smoke.log; tail -3 gpurun_out/r02_final_smoke.log
echo "== bench mx"; timeout 600 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; tail -c 300 gpurun_out/r02_final_bench_n1.err
echo "== bench nv"; timeout 600 python bench.py --kind nv > gpurun_out/r02_final_bench_n1_nv.json 2> gpurun_out/r02_final_bench_n1_nv.err; tail -c 300 gpurun_out/r02_final_bench_n1_nv.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference_arm.json 2> gpurun_out/r02_final_bench_reference_arm.err; cut -c1-600 gpurun_out/r02_final_bench_reference_arm.json
python - <<'PY'
import json
for n in ("r02_final_bench_n1", "r02_final_bench_n1_nv"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, "step %.1f us" % (d["ms_per_step"] * 1e3), "value %.0f" % d["value"], "sustained %.0f" % d.get("value_sustained", 0),
              "gemm %.0f TF" % d["gemm_only_tflops_per_gpu"], "quant %.2f us" % d["quantize_us"], "c4", d.get("c4", {}).get("value"),
              "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("host_copy_ceiling_ms"), "frac", d["roofline"]["frac"], d.get("clocks"))
        r = d.get("reference_gpu", {})
        print("   ref_gpu gemm %s quant %s step %s" % (r.get("gemm_us"), r.get("quantize_us"), r.get("step_us_without_to_blocked")))
    except Exception as e:
        print(n, "no line:", e)
PY
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-c4 --sustain-s 0 > gpurun_out/r02_final_ncu_bench.log 2>&1; tail -2 gpurun_out/r02_final_ncu_bench.log | cut -c1-200
echo "== decode (final build)"; PROBE_PACES=0,-1 timeout 300 python tools/decode_pace_probe.py > gpurun_out/r02_final_decode_mx.jsonl 2> gpurun_out/r02_final_decode_mx.err; cat gpurun_out/r02_final_decode_mx.jsonl; tail -2 gpurun_out/r02_final_decode_mx.err
echo "== backward (final build, library rule)"; timeout 200 python tools/bwd_bench.py > gpurun_out/r02_final_bwd_bench.jsonl 2> gpurun_out/r02_final_bwd.err; grep -v generic gpurun_out/r02_final_bwd_bench.jsonl | cut -c1-170; tail -2 gpurun_out/r02_final_bwd.err
