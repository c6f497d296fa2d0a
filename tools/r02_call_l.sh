#!/bin/bash
# Session-3 call L: the reference benchmark's own layer (K = 8192, N = 57344) at small / medium batch, ours vs the compiled reference
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MSWEEP_70B_SMALL=1 timeout 600 python tools/ref_msweep.py > gpurun_out/r02_s3_ref_msweep_70b_small.jsonl 2> gpurun_out/r02_s3_ref_msweep_70b_small.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_s3_ref_msweep_70b_small.jsonl"):
    d = json.loads(l)
    print("M=%5d GEMM %7.2f / %7.2f (%.2fx)  quant %5.2f / %5.2f  both %7.2f / %7.2f (%.2fx)" % (d["M"], d["ours_gemm_us"], d["ref_gemm_us"], d["ref_gemm_us"] / d["ours_gemm_us"], d["ours_quant_us"], d["ref_quant_us"], d["ours_both_us"], d["ref_both_us"], d["ref_both_us"] / d["ours_both_us"]))
PY
tail -3 gpurun_out/r02_s3_ref_msweep_70b_small.err
