#!/bin/bash
# First GPU session of round 2: prove (or kill) the two opt-ins written blind at the end of round 1, then measure.
# Every step is under its own `timeout` (an unproven tcgen05 kernel may hang; the hybrid kernel's waits trap after ~2 s).
# usage (from the repo root on the GPU box):  bash tools/r02_first_call.sh      -> gpurun_out/r02_*
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 1. hybrid (2,448) tile pairs: parity (small shapes first: -x stops at the first failure)"
B200Q_TEST_HYBRID=1 timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k hybrid > gpurun_out/r02_hybrid_parity.log 2>&1
echo "rc=$?" >> gpurun_out/r02_hybrid_parity.log; tail -6 gpurun_out/r02_hybrid_parity.log
if grep -q "rc=0" gpurun_out/r02_hybrid_parity.log; then
  echo "== 2. hybrid vs default at config 1 (same box, back to back)"
  timeout 90 python bench.py --no-cpu --no-e2e --steps 200 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench.err
  B200Q_GEMM_HYBRID=1 timeout 90 python bench.py --no-cpu --no-e2e --steps 200 > gpurun_out/r02_bench_hybrid.json 2>> gpurun_out/r02_bench.err
  timeout 90 python bench.py --no-cpu --no-e2e --steps 200 --kind nv > gpurun_out/r02_bench_default_nv.json 2>> gpurun_out/r02_bench.err
  B200Q_GEMM_HYBRID=1 timeout 90 python bench.py --no-cpu --no-e2e --steps 200 --kind nv > gpurun_out/r02_bench_hybrid_nv.json 2>> gpurun_out/r02_bench.err
  python - <<'PY'
import json
for n in ("default", "hybrid", "default_nv", "hybrid_nv"):
    try:
        d = json.load(open(f"gpurun_out/r02_bench_{n}.json"))
        print(n, "step %.1f us" % (d["ms_per_step"] * 1e3), "gemm %.0f TF" % d["gemm_only_tflops_per_gpu"], d.get("clocks"))
    except Exception as e:
        print(n, "no line:", e)
PY
fi
echo "== 3. NVFP4 abs_max Hadamard-128 with the reference's sm_100 arithmetic"
B200Q_TEST_NV128_QUIRK=1 timeout 120 python -m pytest tests/test_gpu_reference_lib.py -m gpu -q -rxXs > gpurun_out/r02_reference_lib.log 2>&1
tail -12 gpurun_out/r02_reference_lib.log
echo "== 4. ours vs the compiled reference, M sweep"
timeout 400 python tools/ref_msweep.py > gpurun_out/r02_ref_msweep.jsonl 2> gpurun_out/r02_ref_msweep.err; tail -3 gpurun_out/r02_ref_msweep.err; cat gpurun_out/ref_msweep.md 2>/dev/null
