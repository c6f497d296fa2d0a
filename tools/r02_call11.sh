#!/bin/bash
# Validation + profiles of the round-2 build: decode tests first, full suite, bench lines, ncu launch list + full captures of the
# three dominant kernels (GEMM (2,256), tcgen05 quantiser, decode GEMM), reference M sweep (incl. MXFP8), decode probe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== decode tests"; timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode" > gpurun_out/r02_decode_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_decode_tests.log; tail -4 gpurun_out/r02_decode_tests.log
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -6 gpurun_out/r02_pytest_gpu.log
echo "== decode probe 2"; B200Q_LIB=prof timeout 400 python tools/decode_probe2.py > gpurun_out/r02_decode_probe2.jsonl 2> gpurun_out/r02_decode_probe2.err; grep -v "timeline_cycles_from_entry\": {\"setup" gpurun_out/r02_decode_probe2.jsonl | cut -c1-700; tail -3 gpurun_out/r02_decode_probe2.err
echo "== bench mx"; timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.err
echo "== bench nv"; timeout 600 python bench.py --kind nv > gpurun_out/r02_bench_n1_nv.json 2> gpurun_out/r02_bench_n1_nv.err; tail -c 400 gpurun_out/r02_bench_n1_nv.err
python - <<'PY'
import json
for n in ("r02_bench_n1", "r02_bench_n1_nv"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, "step %.1f us" % (d["ms_per_step"] * 1e3), "value %.0f" % d["value"], "sustained %.0f" % d.get("value_sustained", 0),
              "gemm %.0f TF" % d["gemm_only_tflops_per_gpu"], "quant %.2f us" % d["quantize_us"], "c4", d.get("c4", {}).get("value"),
              "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("host_copy_ceiling_ms"))
        r = d.get("reference_gpu", {})
        print("   ref_gpu gemm %s quant %s step %s" % (r.get("gemm_us"), r.get("quantize_us"), r.get("step_us_without_to_blocked")))
    except Exception as e:
        print(n, "no line:", e)
PY
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-c4 --sustain-s 0 > gpurun_out/r02_ncu_bench.log 2>&1; tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-200
echo "== ncu full: GEMM"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_fp4_kernel -s 6 -c 1 -f -o gpurun_out/r02_prof_gemm python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-c4 --sustain-s 0 > /dev/null 2>&1; ls -la gpurun_out/r02_prof_gemm.ncu-rep
echo "== ncu full: quantiser"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:quantize_tc_kernel -s 6 -c 1 -f -o gpurun_out/r02_prof_quant python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-c4 --sustain-s 0 > /dev/null 2>&1; ls -la gpurun_out/r02_prof_quant.ncu-rep
echo "== ncu full: decode GEMM"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_fp4_decode_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_decode python tools/decode_one.py > /dev/null 2>&1; ls -la gpurun_out/r02_prof_decode.ncu-rep
echo "== ref msweep"; timeout 600 python tools/ref_msweep.py > gpurun_out/r02_ref_msweep.jsonl 2> gpurun_out/r02_ref_msweep.err; tail -3 gpurun_out/r02_ref_msweep.err; cat gpurun_out/ref_msweep.md; grep mxf8 gpurun_out/r02_ref_msweep.jsonl
