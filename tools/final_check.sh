#!/bin/bash
# One short GPU-box session (fits ~3 GPU-minutes): full parity suite, the headline bench line, the quantiser-PDL A/B probe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log
tail -4 gpurun_out/final_pytest.log
timeout 45 python bench.py --no-cpu > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cat gpurun_out/final_bench.json
timeout 35 python tools/pdl_probe.py > gpurun_out/final_pdl_probe.jsonl 2> gpurun_out/final_pdl_probe.err; echo "probe rc=$?"; cat gpurun_out/final_pdl_probe.jsonl
