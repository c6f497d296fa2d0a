#!/bin/bash
# Final check of the shipped library: full GPU suite + smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_pytest_gpu.log; tail -4 gpurun_out/r02_final_pytest_gpu.log
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_smoke.log; tail -3 gpurun_out/r02_final_smoke.log
