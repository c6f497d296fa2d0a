#!/bin/bash
# bench line + ncu launch list of the final library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; tail -c 200 gpurun_out/r02_final_bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_final_bench_n1.json"))
print("step %.1f us" % (d["ms_per_step"] * 1e3), "value %.0f" % d["value"], "sustained %.0f" % d.get("value_sustained", 0), "gemm %.0f TF" % d["gemm_only_tflops_per_gpu"], "quant %.2f us" % d["quantize_us"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d.get("clocks"))
r = d.get("reference_gpu", {}); print("ref_gpu gemm %s quant %s step %s" % (r.get("gemm_us"), r.get("quantize_us"), r.get("step_us_without_to_blocked")))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-c4 --sustain-s 0 > gpurun_out/r02_final_ncu_bench.log 2>&1; tail -1 gpurun_out/r02_final_ncu_bench.log | cut -c1-100
